// Host thread pool of the batched prover / verifier (no CUDA in this header: tests/host/host_pool_test.cpp compiles it
// with g++ and stresses it, tests/test_host_pool.py).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mp {

// Persistent host threads for the per-proof phases of the batched prover / verifier (transcripts and scalar algebra,
// one proof per item).  A sub-batch runs six such phases; starting and joining 16 threads for each cost about as
// much as the arithmetic of a 128-proof phase, so the threads are kept and woken per phase.  One pool per context:
// worker contexts run their phases concurrently with each other.
class HostPool {
 public:
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    start_.notify_all();
    for (auto& t : threads_) t.join();
  }
  // fn(i) for i in [0, count) on `threads` threads (the caller is one of them); returns when all items are done
  template <typename F>
  void run(size_t count, int threads, F&& fn) {
    if (threads <= 1 || count < 2) {
      for (size_t i = 0; i < count; i++) fn(i);
      return;
    }
    const int helpers = (int)std::min<size_t>((size_t)threads - 1, count - 1);
    const std::function<void(size_t)> job = [&fn](size_t i) { fn(i); };
    {
      std::lock_guard<std::mutex> lk(mu_);
      while ((int)threads_.size() < helpers) {
        int idx = (int)threads_.size();
        threads_.emplace_back([this, idx] { loop(idx); });
      }
      job_ = &job;
      count_ = count;
      next_.store(0);
      want_ = helpers;
      active_ = helpers;
      gen_++;
    }
    start_.notify_all();
    for (size_t i = next_.fetch_add(1); i < count; i = next_.fetch_add(1)) fn(i);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return active_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop(int idx) {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      start_.wait(lk, [&] { return quit_ || (job_ && gen_ != seen && idx < want_); });
      if (quit_) return;
      seen = gen_;
      const std::function<void(size_t)>* job = job_;
      const size_t count = count_;
      lk.unlock();
      for (size_t i = next_.fetch_add(1); i < count; i = next_.fetch_add(1)) (*job)(i);
      lk.lock();
      if (--active_ == 0) done_.notify_one();
    }
  }
  std::mutex mu_;
  std::condition_variable start_, done_;
  std::vector<std::thread> threads_;
  const std::function<void(size_t)>* job_ = nullptr;
  size_t count_ = 0;
  std::atomic<size_t> next_{0};
  uint64_t gen_ = 0;
  int want_ = 0, active_ = 0;
  bool quit_ = false;
};

}  // namespace mp
