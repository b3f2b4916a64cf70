// TEST INFRASTRUCTURE ONLY: stress of csrc/host_pool.hpp under g++ (no CUDA).  Returns 0 when every item of every phase ran
// exactly once and the calls returned only after all their items were done.
#include "../../mental-poker_b200/csrc/host_pool.hpp"

#include <stdio.h>
using namespace mp;

extern "C" int h_pool_stress(int pools, int phases, int max_threads) {
  std::atomic<int> failures{0};
  std::vector<std::thread> owners;
  for (int p = 0; p < pools; p++)          // one pool per owner thread, as one pool per worker context
    owners.emplace_back([&, p] {
      HostPool pool;
      uint64_t seed = 88172645463325252ull + (uint64_t)p;
      for (int ph = 0; ph < phases; ph++) {
        seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
        const size_t count = (size_t)(seed % 300);                 // includes 0 and 1
        const int threads = 1 + (int)((seed >> 20) % (uint64_t)max_threads);
        std::vector<std::atomic<int>> hits(count);
        for (auto& h : hits) h.store(0);
        std::atomic<size_t> done{0};
        pool.run(count, threads, [&](size_t i) {
          hits[i].fetch_add(1);
          volatile uint64_t x = i;                                 // uneven item cost
          for (uint64_t k = 0; k < (i % 7) * 200; k++) x = x * 6364136223846793005ull + 1;
          done.fetch_add(1);
        });
        if (done.load() != count) failures.fetch_add(1);           // run() returned before its items finished
        for (size_t i = 0; i < count; i++)
          if (hits[i].load() != 1) failures.fetch_add(1);
      }
    });
  for (auto& t : owners) t.join();
  return failures.load();
}
