// Shared internals of the shuffle_*.cu translation units: device-side state, scratch slots,
// error macros and the few helpers more than one of them needs.
#pragma once
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "ctx.cuh"
#include "diag.cuh"
#include "frvec.cuh"
#include "host_pool.hpp"
#include "msm.cuh"
#include "nvtx_ranges.hpp"
#include "shuffle.cuh"
#include "shuffle_host.hpp"

namespace mp {

// ------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------
struct ShuffleState : ShuffleParamsHost {
  affine* d_ck = nullptr;     // device, Montgomery: h, g_1..g_n, then enc_g, ghat, pk (n + 4 points)
  // fixed-base table of those n + 4 bases for the commitment jobs: tab_ck[w*(n+4) + i] = 2^(c*w) * base_i
  affine* d_tab_ck = nullptr;
  int tab_c = 0;
  uint8_t ck_pk[kPointBytes];  // public key currently in slot n + 3 of d_ck / d_tab_ck
  bool ck_pk_valid = false;
  cudaEvent_t ev = nullptr;   // marks small device->host copies the host waits for mid-stream
  // second stream + MSM workspace: independent launch sequences (the verifier's commitment-space
  // jobs next to its ciphertext MSMs) run concurrently instead of paying their fold latencies in turn
  cudaStream_t aux = nullptr;
  MsmWorkspace* aux_ws = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // own stream + workspace for the prover's diagonal ciphertext products: they only need the first
  // challenge and their result only enters the LAST transcript absorb, so they are queued as soon as x
  // is known and the small launches of rounds B, C and D (main stream) fill in around them
  cudaStream_t bulk = nullptr;
  MsmWorkspace* bulk_ws = nullptr;
  cudaEvent_t ev_bulk_go = nullptr, ev_bulk_done = nullptr;
  // fixed-base tables for remasking: tab[base][j][d-1] = d * 2^(8j) * base, base 0 = g, 1 = pk
  affine* d_tab = nullptr;
  uint8_t tab_pk[kPointBytes];
  bool tab_pk_valid = false;
  // worker contexts of mp_shuffle_prove_batch (own stream / workspace each; same parameters)
  std::vector<mp_ctx*> workers;
  uint64_t params_gen = 0;            // bumped by every set_params
  std::vector<uint64_t> worker_gen;   // generation each worker was configured for
  DiagDevice* diag = nullptr;  // Karatsuba plan of the prover's diagonal products (diag.cu), built on first use
  uint8_t* pinned = nullptr;  // small pinned staging for results
  size_t pinned_cap = 0;
  HostPool pool;              // host threads of the batched prover / verifier phases
  ~ShuffleState() {
    if (d_ck) cudaFree(d_ck);
    if (d_tab_ck) cudaFree(d_tab_ck);
    if (pinned) cudaFreeHost(pinned);
    if (ev) cudaEventDestroy(ev);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (aux) cudaStreamDestroy(aux);
    if (aux_ws) msm_workspace_destroy(aux_ws);
    if (ev_bulk_go) cudaEventDestroy(ev_bulk_go);
    if (ev_bulk_done) cudaEventDestroy(ev_bulk_done);
    if (bulk) cudaStreamDestroy(bulk);
    if (bulk_ws) msm_workspace_destroy(bulk_ws);
    if (d_tab) cudaFree(d_tab);
#ifndef MP_CURVE_BLS12_377
    if (diag) diag_device_destroy(diag);  // (the Karatsuba plan belongs to the large-deck prover: Stark build only)
#endif
    for (mp_ctx* w : workers) ctx_destroy(w);
  }
};

enum Slot {  // ctx->scratch slots owned by this file
  sSmallUp = mp_ctx::kSlotUser,
  sG1Canon, sG1Mont, sG1Scal, sG1Out,
  sCtCanon, sCtMont, sCtScal, sCtOut,
  sResults, sPartials,
  sFrA, sFrB, sFrD, sFrBv, sFrXpow, sFrAme, sFrTmp0, sFrTmp1, sFrTmp2, sFrPairs, sFrSmall,
  sPerm, sRho, sCanonOut, sCtTable,
  sKaraTmp, sKaraPts, sKaraScal, sKaraOut,
  sBadItems,
};

#define CK(x)                                                    \
  do {                                                           \
    cudaError_t _e = (x);                                        \
    if (_e != cudaSuccess) return ctx->cuda_fail(_e, #x);        \
  } while (0)
#define NEED(ptr)                                                                  \
  do {                                                                             \
    if (!(ptr)) return ctx->fail(MP_ERR_CUDA, "device allocation failed (%s)", #ptr); \
  } while (0)

inline FrPow2Table h_pow2_table(const fr& x) {
  FrPow2Table t;
  t.p[0] = x;
  for (int k = 1; k < 32; k++) t.p[k] = fr_sqr(t.p[k - 1]);
  return t;
}

// packs small host arrays into one upload
struct SmallUpload {
  std::vector<uint8_t> bytes;
  size_t add(const void* p, size_t len) {
    size_t off = (bytes.size() + 31) & ~(size_t)31;
    bytes.resize(off + len);
    memcpy(bytes.data() + off, p, len);
    return off;
  }
  size_t add_frs(const std::vector<fr>& v) { return add(v.data(), v.size() * sizeof(fr)); }
};

inline uint8_t* pinned(ShuffleState* S, size_t bytes) {
  if (S->pinned_cap < bytes) {
    if (S->pinned) cudaFreeHost(S->pinned);
    S->pinned = nullptr;
    S->pinned_cap = 0;
    size_t want = std::max<size_t>(bytes * 2, 1 << 16);
    if (cudaMallocHost(&S->pinned, want) != cudaSuccess) return nullptr;
    S->pinned_cap = want;
  }
  return S->pinned;
}

// Decks up to this many cards take the host-scalar lockstep prover / batched verifier (every
// group operation still runs on the GPU); larger decks use the device scalar kernels.  The
// environment override exists so the tests can drive both implementations at the same sizes.
inline size_t small_deck_max() {
  const char* e = getenv("MP_SMALL_DECK_MAX");
  return e ? (size_t)strtoull(e, nullptr, 10) : 8192;
}

// Runs fn(worker, i) for i in [0, B) on P worker contexts (own stream / workspace / tables each),
// one host thread per worker.  Returns the first error; ctx->launches = total kernel launches.
// `sleeping_waits`: the workers' host waits sleep instead of spinning (ctx.cuh stream_wait).
template <typename F>
int32_t run_on_workers(mp_ctx* ctx, int P, uint64_t B, F&& fn, bool sleeping_waits = false) {
  ShuffleState* S = ctx->shuffle;
  while ((int)S->workers.size() < P) {
    mp_ctx* w = nullptr;
    if (ctx_create(&w, ctx->device) != MP_OK) return ctx->fail(MP_ERR_CUDA, "cannot create worker context");
    S->workers.push_back(w);
    S->worker_gen.push_back(0);
  }
  for (int t = 0; t < P; t++) {
    if (S->worker_gen[t] == S->params_gen) continue;
    int32_t st = shuffle_set_params(S->workers[t], S->m, S->n, S->enc_g, S->ck64.data() + kPointBytes, S->ck64.data(), S->ghat);
    if (st != MP_OK) return ctx->fail(st, "worker set_params failed: %s", S->workers[t]->err.c_str());
    S->worker_gen[t] = S->params_gen;
  }
  std::atomic<uint64_t> next{0};
  std::atomic<int32_t> first_err{MP_OK};
  std::atomic<int> launches{0};
  auto run = [&](int t) {
    mp_ctx* w = S->workers[t];
    cudaSetDevice(w->device);
    w->blocking = sleeping_waits;
    for (uint64_t i = next.fetch_add(1); i < B; i = next.fetch_add(1)) {
      if (first_err.load() != MP_OK) break;
      int32_t st = fn(w, i);
      launches.fetch_add(w->launches);
      if (st < 0) {
        int32_t expected = MP_OK;
        if (first_err.compare_exchange_strong(expected, st)) ctx->fail(st, "item %llu: %s", (unsigned long long)i, w->err.c_str());
        break;
      }
    }
  };
  if (P == 1) {
    run(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < P; t++) pool.emplace_back(run, t);
    for (auto& th : pool) th.join();
  }
  ctx->launches = launches.load();
  return first_err.load();
}


// Statement hashes of a large-deck batch, shared.  Every proof's transcript starts with one serial Blake2s pass over the
// statement (parameters, key, input deck, shuffled deck: 17 MB at 2^16 cards, 23.5 ms of one core) -- once in the
// prover, once in the verifier.  With a core per worker that costs nothing extra; when several GPUs' workers share a
// host it is what the GPUs wait for.  Here one background thread hashes the statements of up to
// eight proofs AT ONCE (Blake2sLanes: one vector lane per proof, ~4x the bytes per core-second) and the workers pick up
// their hasher when they need it.  Prover: the head (up to the input deck); the shuffled deck, which only exists after
// the remask kernel, is still hashed by the worker.  Verifier: everything up to the proof.
struct StatementHashes {
  std::vector<Blake2s> started;
  std::vector<char> ready;
  std::mutex mu;
  std::condition_variable cv;
  std::thread th;
  // decks2 == nullptr: head only
  void start(const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* decks, const uint8_t* decks2, size_t N, uint64_t B) {
    started.resize(B);
    ready.assign(B, 0);
    th = std::thread([=] {
      NvtxRange nvtx("statement hashes, 8 lanes");
      const size_t stride = N * kCtBytes;
      for (uint64_t g0 = 0; g0 < B; g0 += Blake2sLanes::kLanes) {
        const int K = (int)std::min<uint64_t>(Blake2sLanes::kLanes, B - g0);
        TranscriptLanes tl(K);
        const uint8_t *d1[Blake2sLanes::kLanes], *d2[Blake2sLanes::kLanes];
        for (int l = 0; l < K; l++) {
          d1[l] = decks + (g0 + l) * stride;
          d2[l] = decks2 ? decks2 + (g0 + l) * stride : nullptr;
        }
        absorb_statement_head_lanes(tl, S, pk, d1, N);
        if (decks2) tl.feed_points64(d2, 2 * N);
        std::lock_guard<std::mutex> lk(mu);
        for (int l = 0; l < K; l++) {
          tl.extract(l, &started[g0 + l]);
          ready[g0 + l] = 1;
        }
        cv.notify_all();
      }
    });
  }
  const Blake2s& wait(uint64_t i) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return ready[i] != 0; });
    return started[i];
  }
  ~StatementHashes() {
    if (th.joinable()) th.join();
  }
};
// shared hashing on?  MP_HASH_LANES=0 / 1 forces it; otherwise on when the caller's thread budget (host_threads) is
// below HALF the number of worker contexts.  Measured on one B200 with the process pinned to 16 / 8 / 6 / 4 vCPUs
// (two concurrent batch calls of 8 workers each, so the budget per call is half of that): own pass per worker 43.3 /
// 41.0 / 38.3 / 33.8 proofs/s, shared 40.6 / 40.3 / 40.3 / 39.1 -- the cross-over is between 4 and 3 threads per call.
inline bool share_statement_hashes(int host_threads, int workers) {
  const char* e = getenv("MP_HASH_LANES");   // read per call: the tests switch it
  if (e) return atoi(e) != 0;
  const int budget = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  return workers > 1 && 2 * budget < workers;
}

// Small-deck batches: fn(worker, p0, count) over [0, B) in chunks of at most `sub` proofs.  A sub-batch alternates host
// phases (transcripts, scalar algebra on host threads) and device phases (the batched MSMs), so one sub-batch at a time
// leaves the GPU idle for the host share of the wall time (measured at 512 x 52 cards, 16 threads: 6.0 ms host beside
// 9.4 ms device).  A few worker contexts working through the chunks keep some sub-batches on the device while others
// are on the host; a batch too small to split runs on the calling context as before.  Chunks stay >= 128 proofs: both
// kinds of phase carry a fixed cost (a chain of dependent launches with their fold tails; waking the host threads).
// Measured, 512 x 52-card proofs proved per second on one B200 with 16 / 4 / 2 host threads:
//   1 context 31.5 k / 22.3 k / 15.7 k    2 contexts 36.3 k / 30.8 k / 23.9 k
//   3 contexts 36.1 k / 34.3 k / 28.0 k   4 contexts 35.2 k / 35.1 k / 32.5 k
// The default is 4: it matters most with few host threads per GPU, which is the 8-GPU case (one host shared by all
// ranks).  MP_SMALL_WORKERS overrides it.
template <typename F>
int32_t run_chunks(mp_ctx* ctx, uint64_t B, size_t sub, F&& fn) {
  static const int workers = [] { const char* e = getenv("MP_SMALL_WORKERS"); return e ? std::max(1, atoi(e)) : 4; }();
  const size_t kMinChunk = 128;
  if (workers < 2 || B < 2 * kMinChunk) {
    int total = 0;
    for (uint64_t p0 = 0; p0 < B; p0 += sub) {
      ctx->launches = 0;
      int32_t st = fn(ctx, p0, (size_t)std::min<uint64_t>(sub, B - p0));
      if (st != MP_OK) return st;
      total += ctx->launches;
    }
    ctx->launches = total;
    return MP_OK;
  }
  const size_t chunk = std::min<size_t>(sub, std::max<size_t>(kMinChunk, (size_t)((B + workers - 1) / workers)));
  const uint64_t chunks = (B + chunk - 1) / chunk;
  return run_on_workers(ctx, (int)std::min<uint64_t>(workers, chunks), chunks, [&](mp_ctx* w, uint64_t c) {
    w->launches = 0;
    return fn(w, c * chunk, (size_t)std::min<uint64_t>(chunk, B - c * chunk));
  });
}

template <typename F>
void parallel_for(size_t count, int threads, F&& fn) {
  if (threads <= 1 || count < 2) {
    for (size_t i = 0; i < count; i++) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&] {
      for (size_t i = next.fetch_add(1); i < count; i = next.fetch_add(1)) fn(i);
    });
  for (auto& th : pool) th.join();
}

// shuffle_setup.cu
// makes table 1 of ShuffleState::d_tab (8-bit fixed-base windows) the table of `pk` (cached across calls)
int32_t shuffle_ensure_pk_table(mp_ctx* ctx, const uint8_t* pk);
// stream / ws default to the context's; pass the auxiliary pair to overlap with other work
int32_t run_g1_jobs(mp_ctx* ctx, const TermList& tl, xyzz** d_out_ret, int* d_bad, cudaStream_t stream = nullptr,
                    MsmWorkspace* ws = nullptr);
int32_t commit_rows_device(mp_ctx* ctx, const fr* d_rows, uint64_t stride, const fr* d_blinds, int count, int len,
                           uint32_t* d_scal, xyzz* d_out);
// out[k*(n+1)] = blinds[k], out[k*(n+1) + 1 + j] = rows[k*stride + j] (canonical; short rows zero padded)
cudaError_t commit_scalars_launch(const fr* d_rows, uint64_t stride, const fr* d_blinds, int count, int n, int len,
                                  uint32_t* d_out, cudaStream_t stream);
// shuffle_prove_batch.cu
int32_t prove_sub_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms, const uint8_t* rhos,
                        const uint8_t* rands, size_t Bs, uint8_t* out_decks, uint8_t* proofs, int threads);

}  // namespace mp
