// Host-side driver of the GPU shuffle argument.  See shuffle.cuh for the design notes.
#include "shuffle.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/mpshuffle.h"
#include "ctx.cuh"
#include "frvec.cuh"
#include "msm.cuh"
#include "transcript.hpp"

namespace mp {

// ------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------
struct ShuffleState {
  int m = 0, n = 0;
  std::vector<uint8_t> ck64;  // (n+1) * 64 canonical: h, g_1 .. g_n   (MSM order of a commitment)
  uint8_t enc_g[64], ghat[64], gsum[64];  // gsum = g_1 + .. + g_n  (com(c,..,c; 0) = c * gsum)
  affine* d_ck = nullptr;     // device, Montgomery: h, g_1..g_n, then enc_g, ghat, pk (n + 4 points)
  // fixed-base table of those n + 4 bases for the commitment jobs: tab_ck[w*(n+4) + i] = 2^(c*w) * base_i
  affine* d_tab_ck = nullptr;
  int tab_c = 0;
  uint8_t ck_pk[64];          // public key currently in slot n + 3 of d_ck / d_tab_ck
  bool ck_pk_valid = false;
  cudaEvent_t ev = nullptr;   // marks small device->host copies the host waits for mid-stream
  // fixed-base tables for remasking: tab[base][j][d-1] = d * 2^(8j) * base, base 0 = g, 1 = pk
  affine* d_tab = nullptr;
  uint8_t tab_pk[64];
  bool tab_pk_valid = false;
  // worker contexts of mp_shuffle_prove_batch (own stream / workspace each; same parameters)
  std::vector<mp_ctx*> workers;
  uint64_t params_gen = 0;            // bumped by every set_params
  std::vector<uint64_t> worker_gen;   // generation each worker was configured for
  uint8_t* pinned = nullptr;  // small pinned staging for results
  size_t pinned_cap = 0;
  ~ShuffleState() {
    if (d_ck) cudaFree(d_ck);
    if (d_tab_ck) cudaFree(d_tab_ck);
    if (pinned) cudaFreeHost(pinned);
    if (ev) cudaEventDestroy(ev);
    if (d_tab) cudaFree(d_tab);
    for (mp_ctx* w : workers) mp_ctx_destroy(w);
  }
};
void shuffle_state_destroy(ShuffleState* s) { delete s; }
int32_t shuffle_m(const mp_ctx* ctx) { return ctx && ctx->shuffle ? ctx->shuffle->m : 0; }
int32_t shuffle_n(const mp_ctx* ctx) { return ctx && ctx->shuffle ? ctx->shuffle->n : 0; }

uint64_t shuffle_proof_len(int32_t m, int32_t n) { return (uint64_t)(11 * m + 8) * 64 + (uint64_t)(5 * n + 9) * 32; }
uint64_t shuffle_randomness_len(int32_t m, int32_t n) { return (uint64_t)11 * m + (uint64_t)5 * n; }

enum Slot {  // ctx->scratch slots owned by this file
  sSmallUp = mp_ctx::kSlotUser,
  sG1Canon, sG1Mont, sG1Scal, sG1Out,
  sCtCanon, sCtMont, sCtScal, sCtOut,
  sResults, sPartials,
  sFrA, sFrB, sFrD, sFrBv, sFrXpow, sFrAme, sFrTmp0, sFrTmp1, sFrTmp2, sFrPairs, sFrSmall,
  sPerm, sRho, sCanonOut, sCtTable,
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

// ------------------------------------------------------------------------------------------
// host-side scalar helpers
// ------------------------------------------------------------------------------------------
static inline fr h_fr(const uint8_t* b) {
  uint32_t w[8];
  memcpy(w, b, 32);
  return fr_from_canonical(w);
}
static inline void h_fr_out(const fr& a, uint8_t* b) {
  uint32_t w[8];
  fr_to_canonical(a, w);
  memcpy(b, w, 32);
}
static std::vector<fr> h_powers(const fr& x, int count) {  // x^0 .. x^(count-1)
  std::vector<fr> p((size_t)std::max(count, 0));
  if (count > 0) p[0] = fr_one();
  for (int k = 1; k < count; k++) p[k] = fr_mul(p[k - 1], x);
  return p;
}
static std::vector<fr> h_frs(const uint8_t* b, int count) {
  std::vector<fr> v((size_t)count);
  for (int i = 0; i < count; i++) v[i] = h_fr(b + 32 * (size_t)i);
  return v;
}
static fr h_dot(const fr* a, const fr* b, int n) {
  fr acc = fr_zero();
  for (int i = 0; i < n; i++) acc = fr_add(acc, fr_mul(a[i], b[i]));
  return acc;
}
static FrPow2Table h_pow2_table(const fr& x) {
  FrPow2Table t;
  t.p[0] = x;
  for (int k = 1; k < 32; k++) t.p[k] = fr_sqr(t.p[k - 1]);
  return t;
}
static bool all_zero(const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (p[i]) return false;
  return true;
}

// A batch of small G1 MSM jobs assembled on the host: job = list of (point, scalar) terms.
struct TermList {
  std::vector<uint8_t> pts;    // 64 B canonical per term
  std::vector<uint32_t> scal;  // 8 words canonical per term
  std::vector<MsmJob> jobs;
  uint32_t start = 0;
  uint32_t count() const { return (uint32_t)(scal.size() / 8); }
  void term(const uint8_t* p64, const fr& s) {
    pts.insert(pts.end(), p64, p64 + 64);
    uint32_t w[8];
    fr_to_canonical(s, w);
    scal.insert(scal.end(), w, w + 8);
  }
  void close_job() {
    jobs.push_back(MsmJob{start, start, count() - start});
    start = count();
  }
};

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

static uint8_t* pinned(ShuffleState* S, size_t bytes) {
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

// ------------------------------------------------------------------------------------------
// small device kernels local to the protocol
// ------------------------------------------------------------------------------------------
// out[k*(n+1)] = blind[k]; out[k*(n+1) + 1 + j] = rows[k*stride + j]  (canonical), j < len; the
// remaining n - len slots of a short row are zero.
__global__ void __launch_bounds__(256) k_commit_scalars(const fr* __restrict__ rows, uint64_t stride,
                                                        const fr* __restrict__ blinds, int count, int n, int len,
                                                        uint32_t* __restrict__ out) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)count * (n + 1)) return;
  uint64_t k = g / (n + 1);
  int j = (int)(g % (n + 1));
  fr v;
  if (j == 0) v = blinds[k];
  else if (j - 1 < len) v = rows[k * stride + (j - 1)];
  else v = fr_zero();
  uint32_t w[8];
  fr_to_canonical(v, w);
#pragma unroll
  for (int i = 0; i < 8; i++) out[g * 8 + i] = w[i];
}

// E_k = diag_k + Enc(b_k*ghat; tau_k):  E[2k] += c1[k] (= tau_k*g),  E[2k+1] += c2[k] (= b_k*ghat + tau_k*pk)
__global__ void __launch_bounds__(64) k_combine_E(xyzz* __restrict__ E, const xyzz* __restrict__ c1,
                                                  const xyzz* __restrict__ c2, int two_m) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 2 * two_m) return;
  xyzz x = E[g], y = (g & 1) ? c2[g >> 1] : c1[g >> 1];
  xyzz_add(x, y);
  E[g] = x;
}

// Fixed-base window tables for remasking (kernel family K4): tab[j * 255 + d - 1] = d * 2^(8j) * P
// for j < 32, d = 1..255, affine Montgomery.  One thread per entry: the scalar d * 2^(8j) has its
// set bits in [8j, 8j + 8), so the double-and-add runs over 8j + 8 bits only.
static constexpr int kTabWin = 32, kTabDigits = 255, kTabSize = kTabWin * kTabDigits;
__global__ void __launch_bounds__(128) k_build_table(const uint32_t* __restrict__ base_canon, affine* __restrict__ tab,
                                                     int* __restrict__ bad) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= kTabSize) return;
  int j = g / kTabDigits;
  uint32_t d = (uint32_t)(g % kTabDigits) + 1;
  affine P = affine_from_canonical(base_canon);
  if (g == 0 && !affine_on_curve(P)) atomicExch(bad, 1);
  xyzz acc = xyzz_identity();
  for (int bit = 7; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((d >> bit) & 1) xyzz_madd(acc, P);
  }
  for (int k = 0; k < 8 * j; k++) acc = xyzz_dbl(acc);
  tab[g] = xyzz_to_affine(acc);  // identity -> (0, 0)
}

// Remask (reference remasking.rs:9-22 -> masking.rs:10-20):  thread (i, comp) computes
// out[i].comp = deck[perm[i]].comp + rho_i * base_comp, base = (g, pk), as 32 table lookups + adds.
__global__ void __launch_bounds__(128) k_remask(const uint32_t* __restrict__ deck_canon, const uint32_t* __restrict__ perm,
                                                const uint32_t* __restrict__ rho_canon, const affine* __restrict__ tab,
                                                uint64_t N, uint32_t* __restrict__ out_canon, int* __restrict__ bad) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 2 * N) return;
  uint64_t i = g >> 1;
  int comp = (int)(g & 1);
  uint64_t src = perm[i];
  if (src >= N) { atomicExch(bad, 2); return; }
  affine card = affine_from_canonical(deck_canon + (src * 2 + comp) * 16);
  if (!affine_on_curve(card)) atomicExch(bad, 1);
  uint32_t k[8];
  {
    // reduce rho below the group order is the caller's contract; a 256-bit value still works
    // because the table covers all 32 bytes
    const uint4* p = reinterpret_cast<const uint4*>(rho_canon + i * 8);
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    k[0] = lo.x; k[1] = lo.y; k[2] = lo.z; k[3] = lo.w; k[4] = hi.x; k[5] = hi.y; k[6] = hi.z; k[7] = hi.w;
  }
  const affine* T = tab + (size_t)comp * kTabSize;
  xyzz acc = xyzz_from_affine(card);
#pragma unroll 1
  for (int j = 0; j < kTabWin; j++) {
    uint32_t d = (k[j >> 2] >> ((j & 3) * 8)) & 0xffu;
    if (d) {
      const uint4* s = reinterpret_cast<const uint4*>(T + j * kTabDigits + (d - 1));
      affine e;
      uint4* dst = reinterpret_cast<uint4*>(&e);
#pragma unroll
      for (int q = 0; q < 4; q++) dst[q] = __ldg(s + q);
      xyzz_madd(acc, e);
    }
  }
  affine r = xyzz_to_affine(acc);
  uint32_t w[16];
  if (affine_is_identity(r)) {
#pragma unroll
    for (int q = 0; q < 16; q++) w[q] = 0;
  } else {
    affine_to_canonical(r, w);
  }
  uint4* o = reinterpret_cast<uint4*>(out_canon + g * 16);
#pragma unroll
  for (int q = 0; q < 4; q++) o[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// (re)builds the table of base `which` (0 = g, 1 = pk) from 64 canonical bytes already on the device
static cudaError_t build_table(ShuffleState* S, int which, const uint8_t* d_base_canon, int* d_bad, cudaStream_t st) {
  if (!S->d_tab) {
    cudaError_t e = cudaMalloc(&S->d_tab, sizeof(affine) * 2 * (size_t)kTabSize);
    if (e != cudaSuccess) return e;
  }
  k_build_table<<<(kTabSize + 127) / 128, 128, 0, st>>>((const uint32_t*)d_base_canon, S->d_tab + (size_t)which * kTabSize, d_bad);
  return cudaGetLastError();
}

// Decks up to this many cards take the host-scalar lockstep prover / batched verifier (every
// group operation still runs on the GPU); larger decks use the device scalar kernels.  The
// environment override exists so the tests can drive both implementations at the same sizes.
static size_t small_deck_max() {
  const char* e = getenv("MP_SMALL_DECK_MAX");
  return e ? (size_t)strtoull(e, nullptr, 10) : 8192;
}

// Runs fn(worker, i) for i in [0, B) on P worker contexts (own stream / workspace / tables each),
// one host thread per worker.  Returns the first error; ctx->launches = total kernel launches.
template <typename F>
static int32_t run_on_workers(mp_ctx* ctx, int P, uint64_t B, F&& fn) {
  ShuffleState* S = ctx->shuffle;
  while ((int)S->workers.size() < P) {
    mp_ctx* w = nullptr;
    if (mp_ctx_create(&w, ctx->device) != MP_OK) return ctx->fail(MP_ERR_CUDA, "cannot create worker context");
    S->workers.push_back(w);
    S->worker_gen.push_back(0);
  }
  for (int t = 0; t < P; t++) {
    if (S->worker_gen[t] == S->params_gen) continue;
    int32_t st = shuffle_set_params(S->workers[t], S->m, S->n, S->enc_g, S->ck64.data() + 64, S->ck64.data(), S->ghat);
    if (st != MP_OK) return ctx->fail(st, "worker set_params failed: %s", mp_last_error_string(S->workers[t]));
    S->worker_gen[t] = S->params_gen;
  }
  std::atomic<uint64_t> next{0};
  std::atomic<int32_t> first_err{MP_OK};
  std::atomic<int> launches{0};
  auto run = [&](int t) {
    mp_ctx* w = S->workers[t];
    cudaSetDevice(w->device);
    for (uint64_t i = next.fetch_add(1); i < B; i = next.fetch_add(1)) {
      if (first_err.load() != MP_OK) break;
      int32_t st = fn(w, i);
      launches.fetch_add(w->launches);
      if (st < 0) {
        int32_t expected = MP_OK;
        if (first_err.compare_exchange_strong(expected, st)) ctx->fail(st, "item %llu: %s", (unsigned long long)i, mp_last_error_string(w));
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

// ------------------------------------------------------------------------------------------
// set-up
// ------------------------------------------------------------------------------------------
static int32_t run_g1_jobs(mp_ctx* ctx, const TermList& tl, xyzz** d_out_ret, int* d_bad) {
  uint32_t T = tl.count();
  int J = (int)tl.jobs.size();
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sG1Canon, (size_t)T * 64 + 64);
  affine* d_mont = (affine*)ctx->scratch(sG1Mont, (size_t)T * sizeof(affine) + 64);
  uint32_t* d_scal = (uint32_t*)ctx->scratch(sG1Scal, (size_t)T * 32 + 64);
  xyzz* d_out = (xyzz*)ctx->scratch(sG1Out, (size_t)J * sizeof(xyzz) + 64);
  NEED(d_canon); NEED(d_mont); NEED(d_scal); NEED(d_out);
  CK(cudaMemcpyAsync(d_canon, tl.pts.data(), (size_t)T * 64, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_scal, tl.scal.data(), (size_t)T * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(points_to_mont((const uint32_t*)d_canon, d_mont, T, d_bad, ctx->stream));
  ctx->launches += 1;
  int c = msm_pick_window(J ? T / J : 1);
  CK(msm_run(ctx->ws, d_scal, T, d_mont, 1, tl.jobs.data(), J, c, d_out, ctx->stream));
  ctx->launches += msm_last_launches(ctx->ws);
  *d_out_ret = d_out;
  return MP_OK;
}

int32_t shuffle_set_params(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g, const uint8_t* ck_g,
                           const uint8_t* ck_h, const uint8_t* ghat) {
  if (!ctx || !enc_g || !ck_g || !ck_h || !ghat) return MP_ERR_INVALID_ARG;
  if (m < 2 || n < 2 || (uint64_t)m * n >= (1ull << 28))
    return ctx->fail(MP_ERR_INVALID_ARG, "shuffle parameters need m >= 2, n >= 2, m*n < 2^28 (got m=%d n=%d)", m, n);
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  if (!ctx->shuffle) ctx->shuffle = new ShuffleState();
  ShuffleState* S = ctx->shuffle;
  S->m = 0;
  S->n = 0;
  S->ck64.resize((size_t)(n + 1) * 64);
  memcpy(S->ck64.data(), ck_h, 64);
  memcpy(S->ck64.data() + 64, ck_g, (size_t)n * 64);
  memcpy(S->enc_g, enc_g, 64);
  memcpy(S->ghat, ghat, 64);
  if (S->d_ck) cudaFree(S->d_ck);
  S->d_ck = nullptr;
  CK(cudaMalloc(&S->d_ck, sizeof(affine) * (size_t)(n + 4)));
  if (!S->ev) CK(cudaEventCreateWithFlags(&S->ev, cudaEventDisableTiming));
  // validate every parameter point and compute gsum = sum g_j with one MSM of unit scalars
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  TermList tl;
  for (int j = 1; j <= n; j++) tl.term(S->ck64.data() + 64 * (size_t)j, fr_one());
  tl.close_job();
  tl.term(ck_h, fr_one()); tl.term(enc_g, fr_one()); tl.term(ghat, fr_one());  // validation only
  tl.close_job();
  xyzz* d_out = nullptr;
  int32_t st = run_g1_jobs(ctx, tl, &d_out, d_bad);
  if (st != MP_OK) return st;
  uint8_t* d_res = (uint8_t*)ctx->scratch(sCanonOut, 64 + 64);
  NEED(d_res);
  CK(xyzz_to_canonical(d_out, (uint32_t*)d_res, 1, ctx->stream));
  ctx->launches += 1;
  // Montgomery copy of the commit key for the prover's commitment jobs
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sG1Canon, (size_t)(n + 3) * 64);
  NEED(d_canon);
  CK(cudaMemcpyAsync(d_canon, S->ck64.data(), (size_t)(n + 1) * 64, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_canon + (size_t)(n + 1) * 64, enc_g, 64, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_canon + (size_t)(n + 2) * 64, ghat, 64, cudaMemcpyHostToDevice, ctx->stream));
  CK(points_to_mont((const uint32_t*)d_canon, S->d_ck, (uint64_t)n + 3, d_bad, ctx->stream));
  CK(build_table(S, 0, d_canon + (size_t)(n + 1) * 64, d_bad, ctx->stream));  // remask table of g
  S->tab_pk_valid = false;
  // fixed-base table for the commitment jobs (the pk column is filled per call)
  S->tab_c = msm_pick_table_window((uint64_t)n + 1);
  if (S->d_tab_ck) cudaFree(S->d_tab_ck);
  S->d_tab_ck = nullptr;
  CK(cudaMalloc(&S->d_tab_ck, sizeof(affine) * (size_t)msm_num_windows(S->tab_c) * (size_t)(n + 4)));
  CK(cudaMemsetAsync(S->d_tab_ck, 0, sizeof(affine) * (size_t)msm_num_windows(S->tab_c) * (size_t)(n + 4), ctx->stream));
  CK(msm_build_table(ctx->ws, S->d_ck, (uint32_t)(n + 4), 0, (uint32_t)(n + 3), S->tab_c, S->d_tab_ck, ctx->stream));
  S->ck_pk_valid = false;
  ctx->launches += 4;
  int bad = 0;
  CK(cudaMemcpyAsync(S->gsum, d_res, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a parameter point is not a canonical point of the Stark curve");
  S->m = m;
  S->n = n;
  S->params_gen++;
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// proof layout (include/mpshuffle.h)
// ------------------------------------------------------------------------------------------
struct Layout {
  size_t cA, cB, cb, hB, zpts, za, zb, zr, zs, zt, svpts, sva, svb, svr, svs, mepts, meE, mea, mer, meb, mes, metau, end;
  Layout(int m, int n) {
    const size_t P = 64, F = 32;
    cA = 0; cB = cA + m * P; cb = cB + m * P; hB = cb + P; zpts = hB + m * P;
    za = zpts + (2 * (size_t)m + 3) * P; zb = za + n * F; zr = zb + n * F; zs = zr + F; zt = zs + F;
    svpts = zt + F; sva = svpts + 3 * P; svb = sva + n * F; svr = svb + n * F; svs = svr + F;
    mepts = svs + F; meE = mepts + (2 * (size_t)m + 1) * P; mea = meE + 4 * (size_t)m * P;
    mer = mea + n * F; meb = mer + F; mes = meb + F; metau = mes + F; end = metau + F;
  }
};

static void absorb_statement(Transcript& fs, const ShuffleState* S, const uint8_t* pk, const uint8_t* deck,
                             const uint8_t* deck2, size_t N, const uint8_t* cA) {
  fs.begin();
  fs.feed_label("shuffle_argument");
  fs.feed_points64(S->enc_g, 1);
  fs.feed_points64(pk, 1);
  fs.feed_points64(S->ck64.data() + 64, (size_t)S->n);
  fs.feed_points64(S->ck64.data(), 1);
  fs.feed_points64(S->ghat, 1);
  fs.feed_points64(deck, 2 * N);
  fs.feed_points64(deck2, 2 * N);
  fs.feed_points64(cA, (size_t)S->m);
  fs.end();
}

// ------------------------------------------------------------------------------------------
// verifier pieces shared by the single-proof and the batched entry points
// ------------------------------------------------------------------------------------------
struct Challenges {
  fr x, y, z, xh, yh, xz, xs, xm;
};

// Every challenge derives from statement + proof bytes (no device round trip).
static Challenges derive_challenges(const ShuffleState* S, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                                    size_t N, const uint8_t* proof, const Layout& L) {
  const int m = S->m;
  Challenges ch;
  Transcript fs;
  absorb_statement(fs, S, pk, deck, deck2, N, proof + L.cA);
  ch.x = fs.challenge();
  fs.begin(); fs.feed_label("shuffle_argument_b"); fs.feed_points64(proof + L.cB, m); fs.end();
  ch.y = fs.challenge();
  ch.z = fs.challenge();
  fs.begin(); fs.feed_label("hadamard_argument"); fs.feed_points64(proof + L.cb, 1); fs.feed_points64(proof + L.hB, m); fs.end();
  ch.xh = fs.challenge();
  ch.yh = fs.challenge();
  fs.begin(); fs.feed_label("zero_argument"); fs.feed_points64(proof + L.zpts, 2 * (size_t)m + 3); fs.end();
  ch.xz = fs.challenge();
  fs.begin(); fs.feed_label("single_value_product_argument"); fs.feed_points64(proof + L.svpts, 3); fs.end();
  ch.xs = fs.challenge();
  fs.begin(); fs.feed_label("multi_exponentiation_argument");
  fs.feed_points64(proof + L.mepts, 2 * (size_t)m + 1); fs.feed_points64(proof + L.meE, 4 * (size_t)m); fs.end();
  ch.xm = fs.challenge();
  return ch;
}

// What the host still has to compare once the device reports which jobs are the identity.
struct HostChecks {
  bool hadamard_bytes_ok;   // c_B'[m-1] == c_b
  bool zero_bytes_ok;       // zero-argument c_D[m+1] == O
  bool svp_first_ok;        // b~_1 == a~_1
  fr svp_last, xs;          // b~_n must equal xs * bstar
  bool multiexp_bytes_ok;   // multi-exp c_B[m] == O
};
static const int kG1Checks = 8;  // H1, Z1, Z2, Z3, S1, S2, M1, M2 -- one MSM job each

// Appends the eight commitment-space equations of the verifier as "sum scalar*point == O" jobs.
static void append_g1_checks(TermList& tl, const ShuffleState* S, const uint8_t* proof, const Layout& L,
                             const Challenges& ch, HostChecks* hc) {
  const int m = S->m, n = S->n;
  const uint8_t* ck_h = S->ck64.data();
  auto ck_g = [&](int j) { return S->ck64.data() + 64 * (size_t)(j + 1); };  // g_{j+1}, j = 0..n-1
  auto P = [&](size_t off, size_t i) { return proof + off + 64 * i; };
  const fr &y = ch.y, &z = ch.z, &xh = ch.xh, &yh = ch.yh, &xs = ch.xs;
  const std::vector<fr> xzp = h_powers(ch.xz, 2 * m + 1);
  const std::vector<fr> xhp = h_powers(xh, m);
  const std::vector<fr> xmp = h_powers(ch.xm, 2 * m);
  const std::vector<fr> z_a = h_frs(proof + L.za, n), z_b = h_frs(proof + L.zb, n);
  const fr z_r = h_fr(proof + L.zr), z_s = h_fr(proof + L.zs), z_t = h_fr(proof + L.zt);
  const std::vector<fr> sv_a = h_frs(proof + L.sva, n), sv_b = h_frs(proof + L.svb, n);
  const fr sv_r = h_fr(proof + L.svr), sv_s = h_fr(proof + L.svs);
  const std::vector<fr> me_a = h_frs(proof + L.mea, n);
  const fr me_r = h_fr(proof + L.mer), me_b = h_fr(proof + L.meb), me_s = h_fr(proof + L.mes);
  const fr one = fr_one();
  // H1: hB[0] == c_D[0] = y*c_A[0] + c_B[0] - z*gsum
  tl.term(P(L.cA, 0), y); tl.term(P(L.cB, 0), one); tl.term(S->gsum, fr_neg(z)); tl.term(P(L.hB, 0), fr_neg(one));
  tl.close_job();
  // Z1: c_A0 + sum_{i=1}^{m-1} xz^i c_D[i] + xz^m (-gsum) - com(a; r)
  {
    fr s1 = fr_zero();
    tl.term(P(L.zpts, 0), one);
    for (int i = 1; i < m; i++) {
      tl.term(P(L.cA, i), fr_mul(xzp[i], y));
      tl.term(P(L.cB, i), xzp[i]);
      s1 = fr_add(s1, xzp[i]);
    }
    tl.term(S->gsum, fr_neg(fr_add(fr_mul(z, s1), xzp[m])));
    tl.term(ck_h, fr_neg(z_r));
    for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(z_a[j]));
    tl.close_job();
  }
  // Z2: sum_{t=0}^{m-2} xz^{m-t} xh^{t+1} hB[t] + xz * sum_{i=1}^{m-1} xh^i hB[i] + c_Bm1 - com(b; s)
  {
    for (int t = 0; t < m; t++) {
      fr c = fr_zero();
      if (t <= m - 2) c = fr_add(c, fr_mul(xzp[m - t], fr_mul(xhp[t], xh)));
      if (t >= 1) c = fr_add(c, fr_mul(xzp[1], xhp[t]));
      tl.term(P(L.hB, t), c);
    }
    tl.term(P(L.zpts, 1), one);
    tl.term(ck_h, fr_neg(z_s));
    for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(z_b[j]));
    tl.close_job();
  }
  // Z3: sum xz^k c_D_k - com(a * b; t)
  {
    fr ab = fr_zero(), yp = one;
    for (int j = 0; j < n; j++) {
      yp = fr_mul(yp, yh);
      ab = fr_add(ab, fr_mul(fr_mul(z_a[j], z_b[j]), yp));
    }
    for (int k = 0; k <= 2 * m; k++) tl.term(P(L.zpts, 2 + k), xzp[k]);
    tl.term(ck_h, fr_neg(z_t));
    tl.term(ck_g(0), fr_neg(ab));
    tl.close_job();
  }
  // S1: xs*c_b + c_d - com(a~; r~)
  tl.term(P(L.cb, 0), xs); tl.term(P(L.svpts, 0), one); tl.term(ck_h, fr_neg(sv_r));
  for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(sv_a[j]));
  tl.close_job();
  // S2: xs*c_Delta + c_delta - com((xs*b~_{i+1} - b~_i*a~_{i+1})_i; s~)
  tl.term(P(L.svpts, 2), xs); tl.term(P(L.svpts, 1), one); tl.term(ck_h, fr_neg(sv_s));
  for (int i = 0; i + 1 < n; i++)
    tl.term(ck_g(i), fr_neg(fr_sub(fr_mul(xs, sv_b[i + 1]), fr_mul(sv_b[i], sv_a[i + 1]))));
  tl.close_job();
  // M1: c_A0 + sum_{j=1}^{m} xm^j c_B[j-1] - com(a; r)
  tl.term(P(L.mepts, 0), one);
  for (int j = 1; j <= m; j++) tl.term(P(L.cB, j - 1), xmp[j]);
  tl.term(ck_h, fr_neg(me_r));
  for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(me_a[j]));
  tl.close_job();
  // M2: sum xm^k c_B_k - com(b; s)
  for (int k = 0; k < 2 * m; k++) tl.term(P(L.mepts, 1 + k), xmp[k]);
  tl.term(ck_h, fr_neg(me_s));
  tl.term(ck_g(0), fr_neg(me_b));
  tl.close_job();
  hc->hadamard_bytes_ok = memcmp(P(L.hB, m - 1), P(L.cb, 0), 64) == 0;
  hc->zero_bytes_ok = all_zero(P(L.zpts, 2 + m + 1), 64);
  hc->svp_first_ok = fr_eq(sv_b[0], sv_a[0]);
  hc->svp_last = sv_b[n - 1];
  hc->xs = xs;
  hc->multiexp_bytes_ok = all_zero(P(L.mepts, 1 + m), 64);
}

// Verdict in the order the reference reaches the checks (product argument first: Hadamard ->
// zero -> single-value product; then multi-exponentiation).  g1_id[0..8): H1 Z1 Z2 Z3 S1 S2 M1 M2;
// ct_ok: both ciphertext equations (Chat == E_m and the multi-exp opening) hold.
static int32_t verdict(const HostChecks& hc, const fr& bstar, const bool* g1_id, bool ct_ok) {
  if (!g1_id[0] || !hc.hadamard_bytes_ok) return MP_VERIFY_HADAMARD;
  if (!hc.zero_bytes_ok || !g1_id[1] || !g1_id[2] || !g1_id[3]) return MP_VERIFY_ZERO;
  if (!g1_id[4] || !g1_id[5] || !hc.svp_first_ok || !fr_eq(hc.svp_last, fr_mul(hc.xs, bstar))) return MP_VERIFY_SVP;
  if (!hc.multiexp_bytes_ok || !ct_ok || !g1_id[6] || !g1_id[7]) return MP_VERIFY_MULTIEXP;
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// verify
// ------------------------------------------------------------------------------------------
int32_t shuffle_verify(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                       const uint8_t* proof, const void* deck_src, const void* deck2_src) {
  if (!ctx || !pk || !deck || !deck2 || !proof) return MP_ERR_INVALID_ARG;
  if (!deck_src) deck_src = deck;     // host copy doubles as the transfer source
  if (!deck2_src) deck2_src = deck2;  // (a device pointer here means the deck is already resident in HBM)
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  const Layout L(m, n);

  // ---- 1. start moving the decks (independent of every challenge)
  const size_t T = 2 * N + 2 * (size_t)m + 3;  // CT arena: deck | E_m | deck2 | E_0..E_{2m-1} | (g,pk) | (O,ghat)
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, T * 128);
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, T * 2 * sizeof(affine));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, T * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, 4 * sizeof(xyzz));
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_ct_scal); NEED(d_ct_out); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  CK(cudaMemcpyAsync(d_ct_canon, deck_src, N * 128, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_canon + N * 128, proof + L.meE + 128 * (size_t)m, 128, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_canon + (N + 1) * 128, deck2_src, N * 128, cudaMemcpyDefault, ctx->stream));
  {
    std::vector<uint8_t> tail((2 * (size_t)m + 2) * 128, 0);
    memcpy(tail.data(), proof + L.meE, 2 * (size_t)m * 128);
    uint8_t* q = tail.data() + 2 * (size_t)m * 128;
    memcpy(q, S->enc_g, 64);
    memcpy(q + 64, pk, 64);
    memcpy(q + 192, S->ghat, 64);  // (identity, ghat)
    CK(cudaMemcpyAsync(d_ct_canon + (2 * N + 1) * 128, tail.data(), tail.size(), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, T * 2, d_bad, ctx->stream));
  ctx->launches += 1;

  // ---- 2. transcript: every challenge derives from statement + proof bytes
  const Challenges ch = derive_challenges(S, pk, deck, deck2, N, proof, L);
  const fr &x = ch.x, &y = ch.y, &z = ch.z, &xm = ch.xm;

  // ---- 3. O(N) scalar vectors on the device
  const std::vector<fr> me_a = h_frs(proof + L.mea, n);
  const fr me_b = h_fr(proof + L.meb), me_tau = h_fr(proof + L.metau);
  const std::vector<fr> xmp = h_powers(xm, 2 * m);
  SmallUpload up;
  fr yz[2] = {y, z};
  size_t o_yz = up.add(yz, sizeof yz);
  std::vector<fr> coef((size_t)m);
  for (int i = 1; i <= m; i++) coef[i - 1] = fr_neg(xmp[m - i]);
  size_t o_coef = up.add_frs(coef);
  size_t o_mea = up.add_frs(me_a);
  std::vector<uint32_t> tailsc((2 * (size_t)m + 2) * 8);
  for (int k = 0; k < 2 * m; k++) fr_to_canonical(xmp[k], &tailsc[8 * (size_t)k]);
  fr_to_canonical(fr_neg(me_tau), &tailsc[8 * (size_t)(2 * m)]);
  fr_to_canonical(fr_neg(me_b), &tailsc[8 * (size_t)(2 * m + 1)]);
  uint32_t minus_one[8];
  fr_to_canonical(fr_neg(fr_one()), minus_one);
  uint8_t* d_small = (uint8_t*)ctx->scratch(sSmallUp, up.bytes.size() + 64);
  fr* d_partials = (fr*)ctx->scratch(sPartials, sizeof(fr) * (fr_powers_blocks(N) + 2));
  NEED(d_small); NEED(d_partials);
  CK(cudaMemcpyAsync(d_small, up.bytes.data(), up.bytes.size(), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_scal + N * 8, minus_one, 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_ct_scal + (2 * N + 1) * 8, tailsc.data(), tailsc.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  fr* d_bstar = d_partials + fr_powers_blocks(N);
  CK(fr_powers(h_pow2_table(x), N, d_ct_scal, nullptr, (const fr*)(d_small + o_yz), d_partials, d_bstar, ctx->stream));
  CK(fr_outer_canonical((const fr*)(d_small + o_coef), (const fr*)(d_small + o_mea), m, n, d_ct_scal + (N + 1) * 8, ctx->stream));
  ctx->launches += 3;

  // ---- 4. the two ciphertext checks (K1): 4 N-term G1 MSMs in one batched launch sequence
  //   job 0:  sum x^i C_i - E_m                                        == O   (Chat == E_m)
  //   job 1:  sum x^k E_k - Enc(b*ghat; tau) - sum (x^{m-i} a_j) C'_ij  == O
  MsmJob ct_jobs[2] = {{0, 0, (uint32_t)(N + 1)}, {(uint32_t)(N + 1), (uint32_t)(N + 1), (uint32_t)(N + 2 * m + 2)}};
  CK(msm_run(ctx->ws, d_ct_scal, T, d_ct_mont, 2, ct_jobs, 2, msm_pick_window(N), d_ct_out, ctx->stream));
  ctx->launches += msm_last_launches(ctx->ws);

  // ---- 5. the commitment-space checks as small G1 jobs (host builds O(m + n) scalars)
  TermList tl;
  HostChecks hc;
  append_g1_checks(tl, S, proof, L, ch, &hc);
  const int J = (int)tl.jobs.size();  // 8
  xyzz* d_g1_out = nullptr;
  int32_t st = run_g1_jobs(ctx, tl, &d_g1_out, d_bad);
  if (st != MP_OK) return st;

  // ---- 6. collect: [bstar | G1 results | CT results | bad flag]
  const size_t res_bytes = sizeof(fr) + (size_t)(J + 4) * sizeof(xyzz) + 16;
  uint8_t* h_res = pinned(S, res_bytes);
  if (!h_res) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  CK(cudaMemcpyAsync(h_res, d_bstar, sizeof(fr), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_res + sizeof(fr), d_g1_out, (size_t)J * sizeof(xyzz), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_res + sizeof(fr) + (size_t)J * sizeof(xyzz), d_ct_out, 4 * sizeof(xyzz), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(h_res + sizeof(fr) + (size_t)(J + 4) * sizeof(xyzz), d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int bad;
  memcpy(&bad, h_res + sizeof(fr) + (size_t)(J + 4) * sizeof(xyzz), sizeof(int));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck or proof point is not a canonical point of the Stark curve");
  fr bstar;
  memcpy(&bstar, h_res, sizeof(fr));
  const xyzz* res = reinterpret_cast<const xyzz*>(h_res + sizeof(fr));
  auto is_id = [&](int j) { return xyzz_is_identity(res[j]); };

  // ---- 7. verdict
  bool g1_id[kG1Checks];
  for (int j = 0; j < kG1Checks; j++) g1_id[j] = is_id(j);
  return verdict(hc, bstar, g1_id, is_id(J) && is_id(J + 1) && is_id(J + 2) && is_id(J + 3));
}

// ------------------------------------------------------------------------------------------
// batched verify (BASELINE config "batch of independent 52-card proofs"): lockstep over B
// proofs -- host threads derive the transcripts and the O(N) scalars of each proof, then ONE
// ciphertext MSM launch sequence (4 jobs per proof) and ONE G1 launch sequence (8 jobs per proof)
// evaluate every group equation of the whole sub-batch.
// ------------------------------------------------------------------------------------------
// flags[g * ncomp + comp] = (sum of the group's job outputs is the identity)
__global__ void __launch_bounds__(64) k_group_identity(const xyzz* __restrict__ outs, int ncomp, int jobs_per_group,
                                                       uint64_t ngroups, uint8_t* __restrict__ flags) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups * ncomp) return;
  uint64_t group = g / ncomp;
  int comp = (int)(g % ncomp);
  xyzz acc = outs[(group * jobs_per_group) * ncomp + comp];
  for (int j = 1; j < jobs_per_group; j++) {
    xyzz v = outs[(group * jobs_per_group + j) * ncomp + comp];
    xyzz_add(acc, v);
  }
  flags[g] = xyzz_is_identity(acc) ? 1 : 0;
}

static int32_t verify_sub_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* decks2,
                                const uint8_t* proofs, size_t Bs, int32_t* statuses, int threads) {
  ShuffleState* S = ctx->shuffle;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  const Layout L(m, n);
  const size_t plen = shuffle_proof_len(m, n);
  const size_t T1 = 8 * (size_t)m + 5 * (size_t)n + 19;  // G1 terms per proof (see append_g1_checks)
  const size_t SM = 2 * (size_t)m + 3;                   // small ciphertext entries per proof
  const size_t ct_total = 2 * Bs * N + Bs * SM;
  // host staging
  std::vector<uint8_t> g1_pts(Bs * T1 * 64), sm_pts(Bs * SM * 128);
  std::vector<uint32_t> g1_scal(Bs * T1 * 8), ct_scal(ct_total * 8);
  std::vector<HostChecks> hcs(Bs);
  std::vector<fr> bstars(Bs);
  std::vector<int> bad_layout(Bs, 0);
  uint32_t minus_one[8];
  fr_to_canonical(fr_neg(fr_one()), minus_one);

  auto work = [&](size_t p) {
    const uint8_t* deck = decks + p * N * 128;
    const uint8_t* deck2 = decks2 + p * N * 128;
    const uint8_t* proof = proofs + p * plen;
    const Challenges ch = derive_challenges(S, pk, deck, deck2, N, proof, L);
    TermList tl;
    append_g1_checks(tl, S, proof, L, ch, &hcs[p]);
    if (tl.count() != T1) { bad_layout[p] = 1; return; }
    memcpy(&g1_pts[p * T1 * 64], tl.pts.data(), T1 * 64);
    memcpy(&g1_scal[p * T1 * 8], tl.scal.data(), T1 * 32);
    // ciphertext scalars: x^i | -(xm^{m-i} a_j) | small: -1, xm^k, -tau, -b
    uint32_t* sx = &ct_scal[(p * N) * 8];
    uint32_t* s2 = &ct_scal[(Bs * N + p * N) * 8];
    uint32_t* ss = &ct_scal[(2 * Bs * N + p * SM) * 8];
    fr xi = fr_one(), yi = fr_zero(), prod = fr_one();
    for (size_t i = 0; i < N; i++) {
      xi = fr_mul(xi, ch.x);
      yi = fr_add(yi, ch.y);
      prod = fr_mul(prod, fr_sub(fr_add(yi, xi), ch.z));
      fr_to_canonical(xi, sx + 8 * i);
    }
    bstars[p] = prod;
    const std::vector<fr> xmp = h_powers(ch.xm, 2 * m);
    const std::vector<fr> me_a = h_frs(proof + L.mea, n);
    for (int i = 1; i <= m; i++) {
      const fr cf = fr_neg(xmp[m - i]);
      for (int j = 0; j < n; j++) fr_to_canonical(fr_mul(cf, me_a[j]), s2 + 8 * ((size_t)(i - 1) * n + j));
    }
    memcpy(ss, minus_one, 32);
    for (int k = 0; k < 2 * m; k++) fr_to_canonical(xmp[k], ss + 8 * (size_t)(1 + k));
    fr_to_canonical(fr_neg(h_fr(proof + L.metau)), ss + 8 * (size_t)(1 + 2 * m));
    fr_to_canonical(fr_neg(h_fr(proof + L.meb)), ss + 8 * (size_t)(2 + 2 * m));
    // small ciphertext points: E_m | E_0..E_{2m-1} | (g, pk) | (O, ghat)
    uint8_t* q = &sm_pts[p * SM * 128];
    memcpy(q, proof + L.meE + 128 * (size_t)m, 128);
    memcpy(q + 128, proof + L.meE, 2 * (size_t)m * 128);
    uint8_t* t = q + 128 * (size_t)(1 + 2 * m);
    memcpy(t, S->enc_g, 64);
    memcpy(t + 64, pk, 64);
    memset(t + 128, 0, 64);
    memcpy(t + 192, S->ghat, 64);
  };
  if (threads <= 1 || Bs < 4) {
    for (size_t p = 0; p < Bs; p++) work(p);
  } else {
    std::vector<std::thread> pool;
    std::atomic<size_t> next{0};
    for (int t = 0; t < threads; t++)
      pool.emplace_back([&] {
        for (size_t p = next.fetch_add(1); p < Bs; p = next.fetch_add(1)) work(p);
      });
    for (auto& th : pool) th.join();
  }
  for (size_t p = 0; p < Bs; p++)
    if (bad_layout[p]) return ctx->fail(MP_ERR_INVALID_ARG, "internal: unexpected verifier term count");

  // device buffers
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, ct_total * 128);
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, ct_total * 2 * sizeof(affine));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, ct_total * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, Bs * 8 * sizeof(xyzz));
  uint8_t* d_g1_canon = (uint8_t*)ctx->scratch(sG1Canon, Bs * T1 * 64);
  affine* d_g1_mont = (affine*)ctx->scratch(sG1Mont, Bs * T1 * sizeof(affine));
  uint32_t* d_g1_scal = (uint32_t*)ctx->scratch(sG1Scal, Bs * T1 * 32);
  xyzz* d_g1_out = (xyzz*)ctx->scratch(sG1Out, Bs * kG1Checks * sizeof(xyzz));
  uint8_t* d_flags = (uint8_t*)ctx->scratch(sResults, Bs * 12 + 64);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_ct_scal); NEED(d_ct_out); NEED(d_g1_canon); NEED(d_g1_mont);
  NEED(d_g1_scal); NEED(d_g1_out); NEED(d_flags); NEED(d_bad);
  cudaStream_t st = ctx->stream;
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  CK(cudaMemcpyAsync(d_ct_canon, decks, Bs * N * 128, cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(d_ct_canon + Bs * N * 128, decks2, Bs * N * 128, cudaMemcpyDefault, st));
  CK(cudaMemcpyAsync(d_ct_canon + 2 * Bs * N * 128, sm_pts.data(), sm_pts.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_ct_scal, ct_scal.data(), ct_scal.size() * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_g1_canon, g1_pts.data(), g1_pts.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_g1_scal, g1_scal.data(), g1_scal.size() * 4, cudaMemcpyHostToDevice, st));
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, ct_total * 2, d_bad, st));
  CK(points_to_mont((const uint32_t*)d_g1_canon, d_g1_mont, Bs * T1, d_bad, st));
  ctx->launches += 2;
  // ciphertext jobs: per proof (deck, E_m) -> group 0, (deck', small tail) -> group 1
  std::vector<MsmJob> jobs(Bs * 4);
  for (size_t p = 0; p < Bs; p++) {
    const uint32_t a = (uint32_t)(p * N), b = (uint32_t)(2 * Bs * N + p * SM), c2 = (uint32_t)(Bs * N + p * N);
    jobs[4 * p + 0] = MsmJob{a, a, (uint32_t)N};
    jobs[4 * p + 1] = MsmJob{b, b, 1};
    jobs[4 * p + 2] = MsmJob{c2, c2, (uint32_t)N};
    jobs[4 * p + 3] = MsmJob{b + 1, b + 1, (uint32_t)(2 * m + 2)};
  }
  CK(msm_run(ctx->ws, d_ct_scal, ct_total, d_ct_mont, 2, jobs.data(), (int)jobs.size(), msm_pick_window(N / 2 + 1), d_ct_out, st));
  ctx->launches += msm_last_launches(ctx->ws);
  k_group_identity<<<(unsigned)((Bs * 4 + 63) / 64), 64, 0, st>>>(d_ct_out, 2, 2, Bs * 2, d_flags);
  // G1 jobs: 8 per proof, contiguous terms
  std::vector<MsmJob> g1jobs(Bs * kG1Checks);
  {
    // per-proof job boundaries are identical: take them from a dry layout
    const uint32_t lens[kG1Checks] = {4u, (uint32_t)(2 * m + n + 1), (uint32_t)(m + n + 2), (uint32_t)(2 * m + 3),
                                      (uint32_t)(n + 3), (uint32_t)(n + 2), (uint32_t)(m + n + 2), (uint32_t)(2 * m + 2)};
    for (size_t p = 0; p < Bs; p++) {
      uint32_t off = (uint32_t)(p * T1);
      for (int j = 0; j < kG1Checks; j++) {
        g1jobs[p * kG1Checks + j] = MsmJob{off, off, lens[j]};
        off += lens[j];
      }
    }
  }
  CK(msm_run(ctx->ws, d_g1_scal, Bs * T1, d_g1_mont, 1, g1jobs.data(), (int)g1jobs.size(), msm_pick_window(T1 / kG1Checks), d_g1_out, st));
  ctx->launches += msm_last_launches(ctx->ws);
  k_group_identity<<<(unsigned)((Bs * kG1Checks + 63) / 64), 64, 0, st>>>(d_g1_out, 1, 1, Bs * kG1Checks, d_flags + Bs * 4);
  CK(cudaGetLastError());
  ctx->launches += 2;
  std::vector<uint8_t> flags(Bs * 12);
  int bad = 0;
  CK(cudaMemcpyAsync(flags.data(), d_flags, Bs * 12, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck or proof point in the batch is not a canonical point of the Stark curve");
  for (size_t p = 0; p < Bs; p++) {
    bool g1_id[kG1Checks];
    for (int j = 0; j < kG1Checks; j++) g1_id[j] = flags[Bs * 4 + p * kG1Checks + j] != 0;
    const uint8_t* cf = &flags[p * 4];
    statuses[p] = verdict(hcs[p], bstars[p], g1_id, cf[0] && cf[1] && cf[2] && cf[3]);
  }
  return MP_OK;
}

int32_t shuffle_verify_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* decks2,
                             const uint8_t* proofs, uint64_t B, int32_t* statuses, int32_t host_threads) {
  if (!ctx || !pk || (B && (!decks || !decks2 || !proofs || !statuses))) return MP_ERR_INVALID_ARG;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n, plen = shuffle_proof_len(m, n);
  if (N > small_deck_max()) {
    // large decks: the single-proof verifier per deck, on a few worker contexts so that one
    // proof's serial statement hash (host) overlaps the other proofs' MSMs (device)
    int P = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
    P = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)P, 8, B}));
    return run_on_workers(ctx, P, B, [&](mp_ctx* w, uint64_t p) {
      int32_t st = shuffle_verify(w, pk, decks + p * N * 128, decks2 + p * N * 128, proofs + p * plen);
      if (st >= 0) statuses[p] = st;
      return st < 0 ? st : MP_OK;
    });
  }
  int threads = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  threads = std::max(1, std::min(threads, 64));
  // sub-batches bounded by the job grid (<= 65535 jobs per launch) and ~2^25 ciphertext terms
  size_t sub = std::min<size_t>(4096, std::max<size_t>(1, ((size_t)1 << 24) / N));
  for (uint64_t p0 = 0; p0 < B; p0 += sub) {
    size_t Bs = (size_t)std::min<uint64_t>(sub, B - p0);
    int launches = ctx->launches;
    int32_t st = verify_sub_batch(ctx, pk, decks + p0 * N * 128, decks2 + p0 * N * 128, proofs + p0 * plen, Bs,
                                  statuses + p0, threads);
    (void)launches;
    if (st != MP_OK) return st;
  }
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// batched prove: B independent shuffle_and_remask calls under the same parameters and key.
// Small-deck proofs are latency-bound (a chain of ~5 dependent MSM launches with a 253-doubling
// fold each), so the batch runs P worker contexts concurrently -- one host thread, CUDA stream
// and workspace each -- and the GPU overlaps their kernels.  Proof i is byte-identical to what
// mp_shuffle_and_remask produces for the same inputs.
// ------------------------------------------------------------------------------------------
bool shuffle_uses_small_deck_path(uint64_t n_cards) { return n_cards <= small_deck_max() && !getenv("MP_BATCH_WORKERS"); }

static int32_t prove_sub_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                               const uint8_t* rhos, const uint8_t* rands, size_t Bs, uint8_t* out_decks, uint8_t* proofs,
                               int threads);

int32_t shuffle_prove_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                            const uint8_t* rhos, const uint8_t* rands, uint64_t B, uint8_t* out_decks,
                            uint8_t* proofs, int32_t host_threads) {
  if (!ctx || !pk || (B && (!decks || !perms || !rhos || !rands || !out_decks || !proofs))) return MP_ERR_INVALID_ARG;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n, plen = shuffle_proof_len(m, n), rlen = shuffle_randomness_len(m, n) * 32;
  int P = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  if (N <= small_deck_max() && !getenv("MP_BATCH_WORKERS")) {
    // small decks: lockstep over sub-batches (job grid <= 65535 per launch; bounded staging memory)
    const int threads = std::max(1, std::min(P, 64));
    const size_t per_proof_jobs = (size_t)(m + 4 + 6 * m);
    size_t sub = std::min<size_t>({(size_t)4096, (size_t)60000 / per_proof_jobs, std::max<size_t>(1, ((size_t)1 << 22) / N)});
    sub = std::max<size_t>(sub, 1);
    int total = 0;
    for (uint64_t p0 = 0; p0 < B; p0 += sub) {
      size_t Bs = (size_t)std::min<uint64_t>(sub, B - p0);
      int32_t st = prove_sub_batch(ctx, pk, decks + p0 * N * 128, perms + p0 * N, rhos + p0 * N * 32, rands + p0 * rlen, Bs,
                                   out_decks + p0 * N * 128, proofs + p0 * plen, threads);
      if (st != MP_OK) return st;
      total += ctx->launches;
    }
    ctx->launches = total;
    return MP_OK;
  }
  // large decks (or MP_BATCH_WORKERS set): concurrent worker contexts running the single-proof path.
  // For 2^16-card decks a few workers are enough to hide each proof's serial Blake2s statement
  // absorb (host) behind the other proofs' kernels (device).
  P = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)P, (uint64_t)(N > small_deck_max() ? 3 : 32), B}));
  return run_on_workers(ctx, P, B, [&](mp_ctx* w, uint64_t i) {
    const void* d_shuffled = nullptr;
    int32_t st = shuffle_remask(w, pk, decks + i * N * 128, perms + i * N, rhos + i * N * 32, N, out_decks + i * N * 128,
                                nullptr, &d_shuffled);
    int l = w->launches;
    if (st == MP_OK) {
      st = shuffle_prove(w, pk, decks + i * N * 128, out_decks + i * N * 128, perms + i * N, rhos + i * N * 32,
                         rands + i * rlen, proofs + i * plen, d_shuffled);
      w->launches += l;
    }
    return st;
  });
}

// ------------------------------------------------------------------------------------------
// remask + commitments (stand-alone entry points; the prover reuses the pieces)
// ------------------------------------------------------------------------------------------
int32_t shuffle_remask(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm, const uint8_t* rho,
                       uint64_t N, uint8_t* out_deck, const void* deck_src, const void** d_out_ret) {
  if (!deck_src) deck_src = deck;
  if (d_out_ret) *d_out_ret = nullptr;
  if (!ctx || !pk || (N && (!deck || !perm || !rho || !out_deck))) return MP_ERR_INVALID_ARG;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  if (N == 0) return MP_OK;
  if (N >= (1ull << 28)) return ctx->fail(MP_ERR_INVALID_ARG, "deck too large");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  // input deck staged in the sCtMont slot, output in sCtCanon: exactly where shuffle_prove wants
  // the shuffled deck, so shuffle_and_remask does not move it twice
  uint8_t* d_deck = (uint8_t*)ctx->scratch(sCtMont, (N + 2) * 2 * sizeof(affine));
  uint8_t* d_out = (uint8_t*)ctx->scratch(sCtCanon, (N + 2) * 128);
  uint32_t* d_perm = (uint32_t*)ctx->scratch(sPerm, N * 4);
  uint8_t* d_rho = (uint8_t*)ctx->scratch(sRho, N * 32 + 64);
  uint8_t* d_pk = (uint8_t*)ctx->scratch(sSmallUp, 256);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_deck); NEED(d_out); NEED(d_perm); NEED(d_rho); NEED(d_pk); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  if (!S->tab_pk_valid || memcmp(S->tab_pk, pk, 64) != 0) {  // the pk table is cached across calls
    CK(cudaMemcpyAsync(d_pk, pk, 64, cudaMemcpyHostToDevice, ctx->stream));
    CK(build_table(S, 1, d_pk, d_bad, ctx->stream));
    memcpy(S->tab_pk, pk, 64);
    S->tab_pk_valid = true;
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(d_deck, deck_src, N * 128, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(d_perm, perm, N * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_rho, rho, N * 32, cudaMemcpyHostToDevice, ctx->stream));
  k_remask<<<(unsigned)((2 * N + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t*)d_deck, d_perm, (const uint32_t*)d_rho,
                                                                     S->d_tab, N, (uint32_t*)d_out, d_bad);
  CK(cudaGetLastError());
  ctx->launches += 1;
  int bad = 0;
  CK(cudaMemcpyAsync(out_deck, d_out, N * 128, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (bad == 2) return ctx->fail(MP_ERR_INVALID_ARG, "permutation entry out of range");
  if (bad) {
    S->tab_pk_valid = false;
    return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck point or the public key is not on the Stark curve");
  }
  if (d_out_ret) *d_out_ret = d_out;
  return MP_OK;
}

// commitments of `count` rows that live on the device (Montgomery), result XYZZ on the device
static int32_t commit_rows_device(mp_ctx* ctx, const fr* d_rows, uint64_t stride, const fr* d_blinds, int count, int len,
                                  uint32_t* d_scal, xyzz* d_out) {
  ShuffleState* S = ctx->shuffle;
  const int n = S->n;
  uint64_t total = (uint64_t)count * (n + 1);
  k_commit_scalars<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_rows, stride, d_blinds, count, n, len, d_scal);
  CK(cudaGetLastError());
  ctx->launches += 1;
  std::vector<MsmJob> jobs((size_t)count);
  for (int k = 0; k < count; k++) jobs[k] = MsmJob{(uint32_t)(k * (n + 1)), 0, (uint32_t)(n + 1)};
  CK(msm_run(ctx->ws, d_scal, total, S->d_tab_ck, 1, jobs.data(), count, S->tab_c, d_out, ctx->stream, 0, -1,
             (uint32_t)(n + 4)));
  ctx->launches += msm_last_launches(ctx->ws);
  return MP_OK;
}

int32_t shuffle_commit_batch(mp_ctx* ctx, const uint8_t* values, const uint8_t* blinds, uint64_t k, uint64_t len,
                             uint8_t* out) {
  if (!ctx || (k && (!blinds || !out)) || (k && len && !values)) return MP_ERR_INVALID_ARG;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  if (len > (uint64_t)S->n) return ctx->fail(MP_ERR_INVALID_ARG, "vector length %llu exceeds the commit key length %d",
                                             (unsigned long long)len, S->n);
  if (k == 0) return MP_OK;
  if (k * (S->n + 1) >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "too many commitments in one batch");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  uint32_t* d_in = (uint32_t*)ctx->scratch(sFrTmp0, (k * len + k) * 32 + 64);
  fr* d_rows = (fr*)ctx->scratch(sFrTmp1, (k * len + k) * 32 + 64);
  uint32_t* d_scal = (uint32_t*)ctx->scratch(sG1Scal, k * (S->n + 1) * 32);
  xyzz* d_res = (xyzz*)ctx->scratch(sG1Out, k * sizeof(xyzz));
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sCanonOut, k * 64);
  NEED(d_in); NEED(d_rows); NEED(d_scal); NEED(d_res); NEED(d_canon);
  if (len) CK(cudaMemcpyAsync(d_in, values, k * len * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_in + k * len * 8, blinds, k * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(fr_from_canonical_vec(d_in, d_rows, k * len + k, ctx->stream));
  ctx->launches += 1;
  int32_t st = commit_rows_device(ctx, d_rows, len, d_rows + k * len, (int)k, (int)len, d_scal, d_res);
  if (st != MP_OK) return st;
  CK(xyzz_to_canonical(d_res, (uint32_t*)d_canon, k, ctx->stream));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out, d_canon, k * 64, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// prove
// ------------------------------------------------------------------------------------------
struct RandCursor {
  const uint8_t* p;
  size_t i = 0;
  fr one() { return h_fr(p + 32 * (i++)); }
  std::vector<fr> vec(int k) {
    std::vector<fr> v((size_t)k);
    for (int j = 0; j < k; j++) v[j] = one();
    return v;
  }
};

int32_t shuffle_prove(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2, const uint32_t* perm,
                      const uint8_t* rho, const uint8_t* rand, uint8_t* proof_out, const void* deck2_src) {
  if (!ctx || !pk || !deck || !deck2 || !perm || !rho || !rand || !proof_out) return MP_ERR_INVALID_ARG;
  if (!deck2_src) deck2_src = deck2;
  ShuffleState* S = ctx->shuffle;
  if (!S || S->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  const Layout L(m, n);
  for (size_t i = 0; i < N; i++)
    if (perm[i] >= N) return ctx->fail(MP_ERR_INVALID_ARG, "permutation entry %zu out of range", i);
  RandCursor rc{rand};
  cudaStream_t st = ctx->stream;
  int32_t rcode;

  // ---- device buffers
  const size_t rows_max = (size_t)std::max(2 * m + 1, m + 4);
  const size_t T2 = N + 2;  // CT arena: deck2 | (g, pk) | (O, ghat)
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, T2 * 128);
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, T2 * 2 * sizeof(affine));
  uint32_t* d_perm = (uint32_t*)ctx->scratch(sPerm, N * 4);
  fr* d_rho = (fr*)ctx->scratch(sRho, N * 32 + 64);
  fr* d_a = (fr*)ctx->scratch(sFrA, N * sizeof(fr));
  fr* d_Ame = (fr*)ctx->scratch(sFrAme, (N + n) * sizeof(fr));        // rows: a0_me | b chunk 1..m
  fr* d_b = d_Ame + n;
  fr* d_Az = (fr*)ctx->scratch(sFrD, (N + n) * sizeof(fr));           // zero-arg rows: a0_z | d rows 1..m-1 | -1
  fr* d_d0 = (fr*)ctx->scratch(sFrTmp2, N * sizeof(fr));              // d (all m rows)
  fr* d_Bv = (fr*)ctx->scratch(sFrBv, N * sizeof(fr));
  fr* d_Bz = (fr*)ctx->scratch(sFrB, (N + n) * sizeof(fr));           // zero-arg rows: x^i Bv[i-1] | dlast | b_{m+1}
  fr* d_xpow = (fr*)ctx->scratch(sFrXpow, N * sizeof(fr));
  fr* d_pairs = (fr*)ctx->scratch(sFrPairs, (size_t)(m + 1) * (m + 1) * sizeof(fr) + (4 * (size_t)m + 8) * sizeof(fr));
  fr* d_partials = (fr*)ctx->scratch(sPartials, sizeof(fr) * (std::max(fr_powers_blocks(N), fr_reduce_blocks(N)) + 4));
  fr* d_rows = (fr*)ctx->scratch(sFrTmp0, (4 * (size_t)n + 64) * sizeof(fr));  // svp rows (3 x n) + response vectors
  uint32_t* d_g1_scal = (uint32_t*)ctx->scratch(sG1Scal, (rows_max * (n + 1) + 12 * (size_t)m + 64) * 32);
  xyzz* d_g1_out = (xyzz*)ctx->scratch(sG1Out, (8 * (size_t)m + 16) * sizeof(xyzz));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, (N + n + 4 * (size_t)m + 8) * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, 8 * (size_t)m * sizeof(xyzz));
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sCanonOut, (8 * (size_t)m + 16) * 64);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_perm); NEED(d_rho); NEED(d_a); NEED(d_Ame); NEED(d_Az); NEED(d_d0);
  NEED(d_Bv); NEED(d_Bz); NEED(d_xpow); NEED(d_pairs); NEED(d_partials); NEED(d_rows); NEED(d_g1_scal); NEED(d_g1_out);
  NEED(d_ct_scal); NEED(d_ct_out); NEED(d_canon); NEED(d_bad);
  const size_t pin_bytes = (8 * (size_t)m + 16) * 64 + (4 * (size_t)n + 4 * (size_t)m + 64) * 32;
  uint8_t* h_pin = pinned(S, pin_bytes);
  if (!h_pin) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");

  // ---- uploads that do not depend on any challenge
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  if (deck2_src != (const void*)d_ct_canon)  // (shuffle_and_remask leaves the remasked deck right here)
    CK(cudaMemcpyAsync(d_ct_canon, deck2_src, N * 128, cudaMemcpyDefault, st));
  {
    uint8_t tail[256];
    memset(tail, 0, sizeof tail);
    memcpy(tail, S->enc_g, 64);
    memcpy(tail + 64, pk, 64);
    memcpy(tail + 192, S->ghat, 64);
    CK(cudaMemcpyAsync(d_ct_canon + N * 128, tail, 256, cudaMemcpyHostToDevice, st));
  }
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, T2 * 2, d_bad, st));
  if (!S->ck_pk_valid || memcmp(S->ck_pk, pk, 64) != 0) {  // pk column of the fixed-base table (cached)
    CK(cudaMemcpyAsync(S->d_ck + (n + 3), d_ct_mont + 2 * N + 1, sizeof(affine), cudaMemcpyDeviceToDevice, st));
    CK(msm_build_table(ctx->ws, S->d_ck, (uint32_t)(n + 4), (uint32_t)(n + 3), 1, S->tab_c, S->d_tab_ck, st));
    memcpy(S->ck_pk, pk, 64);
    S->ck_pk_valid = true;
    ctx->launches += 2;
  }
  CK(cudaMemcpyAsync(d_perm, perm, N * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_xpow, rho, N * 32, cudaMemcpyHostToDevice, st));  // staging: canonical rho
  CK(fr_from_canonical_vec((const uint32_t*)d_xpow, d_rho, N, st));
  ctx->launches += 2;

  // ---- round A: c_A[k] = com(chunk_k(a); r_k),  a_i = perm[i] + 1
  const std::vector<fr> r = rc.vec(m), s = rc.vec(m);
  fr* d_blind = d_pairs;  // small scratch for blinding factors (<= 4m + 8 elements at the end of d_pairs)
  d_blind = d_pairs + (size_t)(m + 1) * (m + 1);
  CK(fr_perm_vectors(d_perm, nullptr, N, d_a, nullptr, st));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(d_blind, r.data(), sizeof(fr) * m, cudaMemcpyHostToDevice, st));
  if ((rcode = commit_rows_device(ctx, d_a, n, d_blind, m, n, d_g1_scal, d_g1_out)) != MP_OK) return rcode;
  CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, m, st));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(proof_out + L.cA, d_canon, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(S->ev, st));
  // Every shuffled-deck point is multiplied by m + 1 scalar rows in the diagonal MSMs, so
  // pre-shifting it once (table[w] = 2^(c w) * point) pays: all windows of a job then share ONE
  // bucket set -- one bucket reduction per job instead of W, no fold doublings, and a wider
  // window (fewer entries).  The table depends on no challenge: it is queued behind c_A and
  // runs on the GPU while the host hashes the statement.
  const int c_diag = msm_pick_table_window(N / 2 + 1);
  affine* d_ct_tab = (affine*)ctx->scratch(sCtTable, (size_t)msm_num_windows(c_diag) * T2 * 2 * sizeof(affine));
  NEED(d_ct_tab);
  CK(msm_build_table(ctx->ws, d_ct_mont, (uint32_t)(T2 * 2), 0, (uint32_t)(N * 2), c_diag, d_ct_tab, st));
  ctx->launches += 2;
  CK(cudaEventSynchronize(S->ev));  // c_A is on the host; the table build continues
  Transcript fs;
  absorb_statement(fs, S, pk, deck, deck2, N, proof_out + L.cA);
  const fr x = fs.challenge();

  // ---- round B: b_i = x^{perm[i]+1}, c_B[k] = com(chunk_k(b); s_k)
  CK(fr_powers(h_pow2_table(x), N, nullptr, d_xpow, nullptr, d_partials, nullptr, st));
  CK(fr_perm_vectors(d_perm, d_xpow, N, nullptr, d_b, st));
  ctx->launches += 2;
  CK(cudaMemcpyAsync(d_blind, s.data(), sizeof(fr) * m, cudaMemcpyHostToDevice, st));
  if ((rcode = commit_rows_device(ctx, d_b, n, d_blind, m, n, d_g1_scal, d_g1_out)) != MP_OK) return rcode;
  CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, m, st));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(proof_out + L.cB, d_canon, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  fs.begin(); fs.feed_label("shuffle_argument_b"); fs.feed_points64(proof_out + L.cB, m); fs.end();
  const fr y = fs.challenge();
  const fr z = fs.challenge();

  // ---- round C: everything whose commitments depend only on (x, y, z)
  // C.1  d = y*a + b - z;  column prefix products Bv;  rho* = -sum rho_i b_i
  std::vector<fr> t((size_t)m);
  for (int k = 0; k < m; k++) t[k] = fr_add(fr_mul(y, r[k]), s[k]);
  {
    fr yz[2] = {y, z};
    CK(cudaMemcpyAsync(d_blind, yz, sizeof yz, cudaMemcpyHostToDevice, st));
    CK(fr_affine_comb(d_a, d_b, d_blind, N, d_d0, st));
    CK(fr_column_prefix_products(d_d0, m, n, d_Bv, st));
    CK(fr_dot(d_rho, d_b, N, d_partials, d_partials + fr_reduce_blocks(N), st));
    ctx->launches += 4;
  }
  // bring the last product row (the SVP witness) and rho* to the host
  fr* h_col = reinterpret_cast<fr*>(h_pin);
  CK(cudaMemcpyAsync(h_col, d_Bv + (size_t)(m - 1) * n, sizeof(fr) * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_col + n, d_partials + fr_reduce_blocks(N), sizeof(fr), cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(S->ev, st));

  // C.2  multi-exponentiation first message (B.5'): a0, r0, (b_k, s_k, tau_k); the 2m diagonal
  //      ciphertext MSMs E_k (K2) are the prover's dominant cost and only need x.
  const fr s_prod = rc.one();                                   // B.2: blinding of c_b
  std::vector<fr> sv((size_t)m);                                // B.3: s_1 = t_1, s_m = s_prod, rest random
  sv[0] = t[0];
  for (int i = 1; i < m - 1; i++) sv[i] = rc.one();
  sv[m - 1] = s_prod;
  // B.4 randomness (drawn now to respect the B.6 order; used in round D)
  const std::vector<fr> z_a0 = rc.vec(n), z_bm1 = rc.vec(n);
  const fr z_r0 = rc.one(), z_sm1 = rc.one();
  std::vector<fr> z_t((size_t)2 * m + 1);
  for (int k = 0; k <= 2 * m; k++) z_t[k] = (k != m + 1) ? rc.one() : fr_zero();
  // B.5 randomness
  const std::vector<fr> sv_d = rc.vec(n);
  const fr sv_rd = rc.one();
  std::vector<fr> sv_delta((size_t)n);
  sv_delta[0] = sv_d[0];
  for (int i = 1; i < n - 1; i++) sv_delta[i] = rc.one();
  sv_delta[n - 1] = fr_zero();
  const fr sv_s1 = rc.one(), sv_sx = rc.one();
  // B.5' randomness
  const std::vector<fr> me_a0 = rc.vec(n);
  const fr me_r0 = rc.one();
  std::vector<fr> me_b((size_t)2 * m), me_s((size_t)2 * m), me_tau((size_t)2 * m);
  for (int k = 0; k < 2 * m; k++) {
    if (k == m) { me_b[k] = fr_zero(); me_s[k] = fr_zero(); me_tau[k] = fr_zero(); /* tau_m = rho*, set below */ }
    else { me_b[k] = rc.one(); me_s[k] = rc.one(); me_tau[k] = rc.one(); }
  }
  if (rc.i != shuffle_randomness_len(m, n)) return ctx->fail(MP_ERR_INVALID_ARG, "internal: randomness count mismatch");

  CK(cudaMemcpyAsync(d_Ame, me_a0.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
  CK(fr_to_canonical_vec(d_Ame, d_ct_scal, N + n, st));  // scalar arena: rows a0 | b_1..b_m
  ctx->launches += 1;
  std::vector<MsmJob> diag((size_t)2 * m);
  for (int k = 0; k < 2 * m; k++) {
    int i0 = std::max(1, m - k), i1 = std::min(m, 2 * m - k);
    diag[k] = MsmJob{(uint32_t)((size_t)(k - m + i0) * n), (uint32_t)((size_t)(i0 - 1) * n), (uint32_t)((size_t)(i1 - i0 + 1) * n)};
  }
  {
    // one launch sequence normally; very large decks are split so that a call stays below the
    // 2^32-entry limit of the sort (entries = terms * windows)
    const uint64_t max_terms = ((1ull << 31) / (uint64_t)msm_num_windows(c_diag));
    for (int k0 = 0; k0 < 2 * m;) {
      int k1 = k0;
      uint64_t terms = 0;
      while (k1 < 2 * m && (k1 == k0 || terms + diag[k1].len <= max_terms)) terms += diag[k1++].len;
      CK(msm_run(ctx->ws, d_ct_scal, N + n, d_ct_tab, 2, diag.data() + k0, k1 - k0, c_diag, d_ct_out + 2 * (size_t)k0, st, 0, -1,
                 (uint32_t)T2));
      ctx->launches += msm_last_launches(ctx->ws);
      k0 = k1;
    }
  }

  // wait for col / rho* only (the event precedes the diagonal MSMs, which keep the GPU busy
  // while the host prepares the next batch)
  CK(cudaEventSynchronize(S->ev));
  std::vector<fr> col(h_col, h_col + n);
  const fr rho_star = fr_neg(h_col[n]);
  me_tau[m] = rho_star;

  // C.4  SVP first message on the host side of the scalars (O(n)), committed on the device
  std::vector<fr> bk((size_t)n);
  bk[0] = col[0];
  for (int i = 1; i < n; i++) bk[i] = fr_mul(bk[i - 1], col[i]);
  {
    std::vector<fr> rows3((size_t)3 * n, fr_zero());
    for (int i = 0; i < n; i++) rows3[i] = sv_d[i];
    for (int i = 0; i + 1 < n; i++) {
      rows3[(size_t)n + i] = fr_neg(fr_mul(sv_delta[i], sv_d[i + 1]));
      rows3[(size_t)2 * n + i] = fr_sub(fr_sub(sv_delta[i + 1], fr_mul(col[i + 1], sv_delta[i])), fr_mul(bk[i], sv_d[i + 1]));
    }
    CK(cudaMemcpyAsync(d_rows, rows3.data(), sizeof(fr) * 3 * n, cudaMemcpyHostToDevice, st));
  }
  // C.5  ONE G1 batch (one Pippenger launch sequence = one fold latency):
  //      rows      Hadamard c_B[0..m) = com(Bv[i]; sv[i]) (c_B[0] = c_D[0], c_B[m-1] = c_b), SVP c_d,
  //                c_delta, c_Delta, multi-exp c_A0                              (n+1 terms each)
  //      pairs     multi-exp c_B_k = s_k*h + b_k*g_1                               (2 terms)
  //      enc c1/c2 Enc(b_k*ghat; tau_k) = (tau_k*g, b_k*ghat + tau_k*pk)           (1 / 2 terms)
  //      then E_k = diag_k + enc_k.
  {
    const int R = m + 4;
    std::vector<fr> blinds((size_t)R);
    for (int i = 0; i < m; i++) blinds[i] = sv[i];
    blinds[m] = sv_rd; blinds[m + 1] = sv_s1; blinds[m + 2] = sv_sx; blinds[m + 3] = me_r0;
    CK(cudaMemcpyAsync(d_blind, blinds.data(), sizeof(fr) * R, cudaMemcpyHostToDevice, st));
    const uint64_t tot = (uint64_t)(n + 1);
    k_commit_scalars<<<(unsigned)((tot * m + 255) / 256), 256, 0, st>>>(d_Bv, n, d_blind, m, n, n, d_g1_scal);
    k_commit_scalars<<<(unsigned)((tot * 3 + 255) / 256), 256, 0, st>>>(d_rows, n, d_blind + m, 3, n, n, d_g1_scal + tot * m * 8);
    k_commit_scalars<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_Ame, n, d_blind + m + 3, 1, n, n, d_g1_scal + tot * (m + 3) * 8);
    CK(cudaGetLastError());
    ctx->launches += 3;
    const size_t nsmall = 10 * (size_t)m;  // 2m pairs + 2m singles + 2m pairs
    std::vector<uint32_t> sc(nsmall * 8);
    for (int k = 0; k < 2 * m; k++) {
      fr_to_canonical(me_s[k], &sc[(size_t)(2 * k) * 8]);
      fr_to_canonical(me_b[k], &sc[(size_t)(2 * k + 1) * 8]);
      fr_to_canonical(me_tau[k], &sc[(size_t)(4 * m + k) * 8]);
      fr_to_canonical(me_b[k], &sc[(size_t)(6 * m + 2 * k) * 8]);
      fr_to_canonical(me_tau[k], &sc[(size_t)(6 * m + 2 * k + 1) * 8]);
    }
    const uint32_t base = (uint32_t)(tot * R);
    CK(cudaMemcpyAsync(d_g1_scal + (size_t)base * 8, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice, st));
    std::vector<MsmJob> jobs;
    for (int k = 0; k < R; k++) jobs.push_back(MsmJob{(uint32_t)(k * tot), 0, (uint32_t)tot});
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{base + 2 * k, 0, 2});                               // (h, g_1)
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{base + 4 * m + k, (uint32_t)(n + 1), 1});           // enc_g
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{base + 6 * m + 2 * k, (uint32_t)(n + 2), 2});       // (ghat, pk)
    CK(msm_run(ctx->ws, d_g1_scal, base + nsmall, S->d_tab_ck, 1, jobs.data(), (int)jobs.size(), S->tab_c, d_g1_out, st, 0, -1,
               (uint32_t)(n + 4)));
    ctx->launches += msm_last_launches(ctx->ws);
    k_combine_E<<<(4 * m + 63) / 64, 64, 0, st>>>(d_ct_out, d_g1_out + R + 2 * m, d_g1_out + R + 4 * m, 2 * m);
    CK(cudaGetLastError());
    CK(xyzz_to_canonical(d_ct_out, (uint32_t*)d_canon, 4 * (size_t)m, st));
    uint8_t* d_canon2 = d_canon + 4 * (size_t)m * 64;
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon2, (size_t)R + 2 * m, st));
    ctx->launches += 3;
    CK(cudaMemcpyAsync(proof_out + L.meE, d_canon, 4 * (size_t)m * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.hB, d_canon2, (size_t)m * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.svpts, d_canon2 + (size_t)m * 64, 3 * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.mepts, d_canon2 + (size_t)(m + 3) * 64, (size_t)(2 * m + 1) * 64, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  memcpy(proof_out + L.cb, proof_out + L.hB + 64 * (size_t)(m - 1), 64);  // c_b = c_B[m-1]
  fs.begin(); fs.feed_label("hadamard_argument"); fs.feed_points64(proof_out + L.cb, 1); fs.feed_points64(proof_out + L.hB, m); fs.end();
  const fr xh = fs.challenge();
  const fr yh = fs.challenge();

  // ---- round D: zero argument (B.4) on A' = (a0 | d_2..d_m | -1), B' = (xh^i Bv_i | dlast | b_{m+1})
  const std::vector<fr> xhp = h_powers(xh, m);
  {
    // A rows
    CK(cudaMemcpyAsync(d_Az, z_a0.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_Az + n, d_d0 + n, sizeof(fr) * (N - n), cudaMemcpyDeviceToDevice, st));
    std::vector<fr> m1((size_t)n, fr_neg(fr_one()));
    CK(cudaMemcpyAsync(d_Az + N, m1.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
    // B rows: row i-1 = xh^i * Bv[i-1] (i = 1..m-1); row m-1 = sum_{i=1}^{m-1} xh^i Bv[i]; row m = b_{m+1}
    SmallUpload up;
    std::vector<fr> coef(xhp.begin() + 1, xhp.end());  // xh^1 .. xh^{m-1}
    size_t o_coef = up.add_frs(coef);
    std::vector<fr> yp((size_t)n);
    fr acc = fr_one();
    for (int j = 0; j < n; j++) { acc = fr_mul(acc, yh); yp[j] = acc; }
    size_t o_yp = up.add_frs(yp);
    uint8_t* d_small = (uint8_t*)ctx->scratch(sSmallUp, up.bytes.size() + 64);
    NEED(d_small);
    CK(cudaMemcpyAsync(d_small, up.bytes.data(), up.bytes.size(), cudaMemcpyHostToDevice, st));
    CK(fr_scale_rows(d_Bv, (const fr*)(d_small + o_coef), m - 1, n, d_Bz, st));
    CK(fr_lincomb_rows(d_Bv + n, n, (const fr*)(d_small + o_coef), m - 1, n, d_Bz + (size_t)(m - 1) * n, st));
    CK(cudaMemcpyAsync(d_Bz + N, z_bm1.data(), sizeof(fr) * n, cudaMemcpyHostToDevice, st));
    fr* d_dk = d_pairs + (size_t)(m + 1) * (m + 1) + 2 * (size_t)m + 4;  // 2m+1 diagonal sums
    CK(fr_bilinear_diagonals(d_Az, d_Bz, (const fr*)(d_small + o_yp), m + 1, n, d_pairs, d_dk, st));
    ctx->launches += 4;
    // commitments: c_A0 = com(a0; r0), c_Bm1 = com(b_{m+1}; s_{m+1}), c_D_k = com(d_k; t_k)
    fr bl[2] = {z_r0, z_sm1};
    CK(cudaMemcpyAsync(d_blind, bl, sizeof bl, cudaMemcpyHostToDevice, st));
    uint64_t tot = (uint64_t)(n + 1);
    k_commit_scalars<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_Az, n, d_blind, 1, n, n, d_g1_scal);
    k_commit_scalars<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_Bz + N, n, d_blind + 1, 1, n, n, d_g1_scal + tot * 8);
    CK(cudaGetLastError());
    uint32_t* d_pairs_scal = d_g1_scal + 2 * tot * 8;  // (t_k, d_k) pairs
    std::vector<uint32_t> tk((size_t)(2 * m + 1) * 8);
    for (int k = 0; k <= 2 * m; k++) fr_to_canonical(z_t[k], &tk[(size_t)k * 8]);
    // t_k at even slots (strided copy), d_k at odd slots
    CK(cudaMemcpy2DAsync(d_pairs_scal, 64, tk.data(), 32, 32, 2 * (size_t)m + 1, cudaMemcpyHostToDevice, st));
    CK(fr_scatter_canonical(d_dk, 2 * (size_t)m + 1, d_pairs_scal, 1, 2, st));
    ctx->launches += 3;
    std::vector<MsmJob> jobs = {MsmJob{0, 0, (uint32_t)tot}, MsmJob{(uint32_t)tot, 0, (uint32_t)tot}};
    for (int k = 0; k <= 2 * m; k++) jobs.push_back(MsmJob{(uint32_t)(2 * tot + 2 * k), 0, 2});
    CK(msm_run(ctx->ws, d_g1_scal, 2 * tot + 2 * (2 * (size_t)m + 1), S->d_tab_ck, 1, jobs.data(), (int)jobs.size(),
               S->tab_c, d_g1_out, st, 0, -1, (uint32_t)(n + 4)));
    ctx->launches += msm_last_launches(ctx->ws);
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, 2 * (size_t)m + 3, st));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(proof_out + L.zpts, d_canon, (2 * (size_t)m + 3) * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  fs.begin(); fs.feed_label("zero_argument"); fs.feed_points64(proof_out + L.zpts, 2 * (size_t)m + 3); fs.end();
  const fr xz = fs.challenge();
  fs.begin(); fs.feed_label("single_value_product_argument"); fs.feed_points64(proof_out + L.svpts, 3); fs.end();
  const fr xs = fs.challenge();
  fs.begin(); fs.feed_label("multi_exponentiation_argument");
  fs.feed_points64(proof_out + L.mepts, 2 * (size_t)m + 1); fs.feed_points64(proof_out + L.meE, 4 * (size_t)m); fs.end();
  const fr xm = fs.challenge();

  // ---- responses.  Device: the three O(N) row combinations; host: the O(m + n) rest.
  const std::vector<fr> xzp = h_powers(xz, 2 * m + 1);
  const std::vector<fr> xmp = h_powers(xm, 2 * m);
  {
    SmallUpload up;
    std::vector<fr> ca(xzp.begin(), xzp.begin() + m + 1);       // a = sum_{i=0}^{m} xz^i A'_i
    std::vector<fr> cb((size_t)m + 1);                          // b = sum_{j=0}^{m} xz^{m-j} B'_j
    for (int j = 0; j <= m; j++) cb[j] = xzp[m - j];
    std::vector<fr> cm(xmp.begin(), xmp.begin() + m + 1);       // a_me = sum_{j=0}^{m} xm^j Ame_j
    size_t o_a = up.add_frs(ca), o_b = up.add_frs(cb), o_m = up.add_frs(cm);
    uint8_t* d_small = (uint8_t*)ctx->scratch(sSmallUp, up.bytes.size() + 64);
    NEED(d_small);
    CK(cudaMemcpyAsync(d_small, up.bytes.data(), up.bytes.size(), cudaMemcpyHostToDevice, st));
    fr* d_resp = d_rows;  // 3 x n
    CK(fr_lincomb_rows(d_Az, n, (const fr*)(d_small + o_a), m + 1, n, d_resp, st));
    CK(fr_lincomb_rows(d_Bz, n, (const fr*)(d_small + o_b), m + 1, n, d_resp + n, st));
    CK(fr_lincomb_rows(d_Ame, n, (const fr*)(d_small + o_m), m + 1, n, d_resp + 2 * (size_t)n, st));
    uint32_t* d_resp_canon = d_g1_scal;
    CK(fr_to_canonical_vec(d_resp, d_resp_canon, 3 * (size_t)n, st));
    ctx->launches += 4;
    CK(cudaMemcpyAsync(proof_out + L.za, d_resp_canon, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.zb, d_resp_canon + (size_t)n * 8, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proof_out + L.mea, d_resp_canon + 2 * (size_t)n * 8, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  }
  // zero-argument blinding responses: r' = (r0, t_2..t_m, 0), s' = (xh^i sv_i.., sum xh^i sv_{i+1}, s_{m+1})
  {
    std::vector<fr> rext((size_t)m + 1), sext((size_t)m + 1), xr((size_t)m + 1);
    rext[0] = z_r0;
    for (int i = 1; i < m; i++) rext[i] = t[i];
    rext[m] = fr_zero();
    for (int i = 1; i < m; i++) sext[i - 1] = fr_mul(xhp[i], sv[i - 1]);
    sext[m - 1] = h_dot(xhp.data() + 1, sv.data() + 1, m - 1);
    sext[m] = z_sm1;
    for (int j = 0; j <= m; j++) xr[j] = xzp[m - j];
    h_fr_out(h_dot(xzp.data(), rext.data(), m + 1), proof_out + L.zr);
    h_fr_out(h_dot(xr.data(), sext.data(), m + 1), proof_out + L.zs);
    h_fr_out(h_dot(xzp.data(), z_t.data(), 2 * m + 1), proof_out + L.zt);
  }
  // SVP responses
  for (int i = 0; i < n; i++) {
    h_fr_out(fr_add(fr_mul(xs, col[i]), sv_d[i]), proof_out + L.sva + 32 * (size_t)i);
    h_fr_out(fr_add(fr_mul(xs, bk[i]), sv_delta[i]), proof_out + L.svb + 32 * (size_t)i);
  }
  h_fr_out(fr_add(fr_mul(xs, s_prod), sv_rd), proof_out + L.svr);
  h_fr_out(fr_add(fr_mul(xs, sv_sx), sv_s1), proof_out + L.svs);
  // multi-exp responses
  {
    std::vector<fr> rext((size_t)m + 1);
    rext[0] = me_r0;
    for (int j = 1; j <= m; j++) rext[j] = s[j - 1];
    h_fr_out(h_dot(xmp.data(), rext.data(), m + 1), proof_out + L.mer);
    h_fr_out(h_dot(xmp.data(), me_b.data(), 2 * m), proof_out + L.meb);
    h_fr_out(h_dot(xmp.data(), me_s.data(), 2 * m), proof_out + L.mes);
    h_fr_out(h_dot(xmp.data(), me_tau.data(), 2 * m), proof_out + L.metau);
  }
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck point or the public key is not a canonical point of the Stark curve");
  return MP_OK;
}


// ------------------------------------------------------------------------------------------
// Lockstep batched prover for small decks (BASELINE config: batch of 52-card proofs).
//
// A 52-card proof is a chain of five dependent MSM launch sequences whose cost is latency, not
// work, so B proofs advance TOGETHER: every Fiat-Shamir round is one batched MSM launch over the
// jobs of all proofs (fixed-base table mode for every commitment, variable-base for the diagonal
// ciphertext MSMs).  At these sizes the scalar-field work is O(m^2 n) ~ a few thousand
// multiplications per proof, so it runs on the host threads that also own the transcripts (the
// device kernels of frvec.cu are for the 2^16-card path).  Results are byte-identical to
// mp_shuffle_and_remask (tests/test_gpu_shuffle.py).
// ------------------------------------------------------------------------------------------
namespace {

struct ProverHost {  // one proof's host-side state across the rounds
  Transcript fs;
  std::vector<fr> r, s, a, b, t, d, Bv, col, bk, sv;
  std::vector<fr> z_a0, z_bm1, z_t, sv_d, sv_delta, me_a0, me_b, me_s, me_tau;
  std::vector<fr> Az, Bz;  // zero-argument rows, (m + 1) x n each
  fr s_prod, z_r0, z_sm1, sv_rd, sv_s1, sv_sx, me_r0;
  fr x, y, z, xh, yh;
  std::vector<fr> xhp;
};

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

inline void put_fr(uint32_t* dst, const fr& v) { fr_to_canonical(v, dst); }

// E[(p*2m + k)*2 + comp] += enc part of proof p:  comp 0 -> g1[p*stride + c1_off + k], comp 1 -> c2_off + k
__global__ void __launch_bounds__(64) k_combine_E_batch(xyzz* __restrict__ E, const xyzz* __restrict__ g1, uint32_t stride,
                                                        uint32_t c1_off, uint32_t c2_off, int two_m, uint64_t total) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  uint64_t p = g / (2 * (uint64_t)two_m);
  uint32_t rem = (uint32_t)(g % (2 * (uint64_t)two_m));
  uint32_t k = rem >> 1, comp = rem & 1;
  xyzz x = E[g], y = g1[p * stride + (comp ? c2_off : c1_off) + k];
  xyzz_add(x, y);
  E[g] = x;
}

}  // namespace

static int32_t prove_sub_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                               const uint8_t* rhos, const uint8_t* rands, size_t Bs, uint8_t* out_decks, uint8_t* proofs,
                               int threads) {
  ShuffleState* S = ctx->shuffle;
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n, plen = shuffle_proof_len(m, n), rlen = shuffle_randomness_len(m, n) * 32;
  const Layout L(m, n);
  cudaStream_t st = ctx->stream;
  const uint32_t nb = (uint32_t)(n + 4);       // bases of the fixed-base table: h, g_1..g_n, enc_g, ghat, pk
  const uint32_t tot = (uint32_t)(n + 1);      // scalars of one row commitment: blind + n values
  const int R = m + 4;                         // row commitments of round C
  const uint32_t SC = (uint32_t)(R * tot + 10 * m);   // round-C scalars per proof
  const uint32_t JC = (uint32_t)(R + 6 * m);          // round-C G1 jobs per proof
  const uint32_t SD = (uint32_t)(2 * tot + 2 * (2 * m + 1));  // round-D scalars per proof
  const uint32_t JD = (uint32_t)(2 * m + 3);

  // ---- round 0: remask every deck with one launch (the batch is one deck of Bs*N cards)
  {
    std::vector<uint32_t> gperm(Bs * N);
    for (size_t p = 0; p < Bs; p++)
      for (size_t i = 0; i < N; i++) {
        if (perms[p * N + i] >= N) return ctx->fail(MP_ERR_INVALID_ARG, "proof %zu: permutation entry %zu out of range", p, i);
        gperm[p * N + i] = (uint32_t)(p * N) + perms[p * N + i];
      }
    const void* d_shuffled = nullptr;
    int32_t rc = shuffle_remask(ctx, pk, decks, gperm.data(), rhos, Bs * N, out_decks, nullptr, &d_shuffled);
    if (rc != MP_OK) return rc;
  }
  int launches = ctx->launches;

  // ---- device buffers
  const size_t rows_max = std::max<size_t>(SC, (size_t)m * tot);
  uint8_t* d_ct_canon = (uint8_t*)ctx->scratch(sCtCanon, (Bs * N + 2) * 128);   // holds the remasked decks already
  affine* d_ct_mont = (affine*)ctx->scratch(sCtMont, (Bs * N + 2) * 2 * sizeof(affine));
  uint32_t* d_ct_scal = (uint32_t*)ctx->scratch(sCtScal, Bs * (N + n) * 32);
  xyzz* d_ct_out = (xyzz*)ctx->scratch(sCtOut, Bs * 4 * (size_t)m * sizeof(xyzz));
  uint32_t* d_g1_scal = (uint32_t*)ctx->scratch(sG1Scal, Bs * rows_max * 32 + 64);
  xyzz* d_g1_out = (xyzz*)ctx->scratch(sG1Out, Bs * JC * sizeof(xyzz));
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sCanonOut, Bs * (JC + 4 * (size_t)m) * 64);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_ct_canon); NEED(d_ct_mont); NEED(d_ct_scal); NEED(d_ct_out); NEED(d_g1_scal); NEED(d_g1_out); NEED(d_canon); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  // pk column of the fixed-base table
  if (!S->ck_pk_valid || memcmp(S->ck_pk, pk, 64) != 0) {
    uint8_t* d_pk = (uint8_t*)ctx->scratch(sSmallUp, 256);
    NEED(d_pk);
    CK(cudaMemcpyAsync(d_pk, pk, 64, cudaMemcpyHostToDevice, st));
    CK(points_to_mont((const uint32_t*)d_pk, S->d_ck + (n + 3), 1, d_bad, st));
    CK(msm_build_table(ctx->ws, S->d_ck, nb, (uint32_t)(n + 3), 1, S->tab_c, S->d_tab_ck, st));
    memcpy(S->ck_pk, pk, 64);
    S->ck_pk_valid = true;
    launches += 3;
  }
  // the remasked decks (still in d_ct_canon) in Montgomery form for the diagonal MSMs
  CK(points_to_mont((const uint32_t*)d_ct_canon, d_ct_mont, Bs * N * 2, nullptr, st));
  launches += 1;

  std::vector<ProverHost> H(Bs);
  std::vector<uint32_t> h_scal;
  std::vector<uint8_t> h_pts;
  std::vector<MsmJob> jobs;

  auto run_commit_jobs = [&](size_t n_scalars, size_t njobs_total, size_t npoints_out) -> int32_t {
    CK(cudaMemcpyAsync(d_g1_scal, h_scal.data(), n_scalars * 32, cudaMemcpyHostToDevice, st));
    CK(msm_run(ctx->ws, d_g1_scal, n_scalars, S->d_tab_ck, 1, jobs.data(), (int)njobs_total, S->tab_c, d_g1_out, st, 0, -1, nb));
    launches += msm_last_launches(ctx->ws);
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, npoints_out, st));
    launches += 1;
    h_pts.resize(npoints_out * 64);
    CK(cudaMemcpyAsync(h_pts.data(), d_canon, npoints_out * 64, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MP_OK;
  };
  int32_t rc;

  // ---- round A: c_A[k] = com(chunk_k(a); r_k)
  h_scal.assign(Bs * (size_t)m * tot * 8, 0);
  parallel_for(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    RandCursor rcur{rands + p * rlen};
    h.r = rcur.vec(m);
    h.s = rcur.vec(m);
    h.s_prod = rcur.one();
    h.sv.assign((size_t)m, fr_zero());
    for (int i = 1; i < m - 1; i++) h.sv[i] = rcur.one();
    h.z_a0 = rcur.vec(n); h.z_bm1 = rcur.vec(n);
    h.z_r0 = rcur.one(); h.z_sm1 = rcur.one();
    h.z_t.assign((size_t)2 * m + 1, fr_zero());
    for (int k = 0; k <= 2 * m; k++) if (k != m + 1) h.z_t[k] = rcur.one();
    h.sv_d = rcur.vec(n);
    h.sv_rd = rcur.one();
    h.sv_delta.assign((size_t)n, fr_zero());
    h.sv_delta[0] = h.sv_d[0];
    for (int i = 1; i < n - 1; i++) h.sv_delta[i] = rcur.one();
    h.sv_s1 = rcur.one(); h.sv_sx = rcur.one();
    h.me_a0 = rcur.vec(n);
    h.me_r0 = rcur.one();
    h.me_b.assign((size_t)2 * m, fr_zero()); h.me_s = h.me_b; h.me_tau = h.me_b;
    for (int k = 0; k < 2 * m; k++) if (k != m) { h.me_b[k] = rcur.one(); h.me_s[k] = rcur.one(); h.me_tau[k] = rcur.one(); }
    h.a.resize(N);
    for (size_t i = 0; i < N; i++) h.a[i] = fr_from_u64((uint64_t)perms[p * N + i] + 1);
    for (int k = 0; k < m; k++) {
      uint32_t* dst = &h_scal[((p * m + k) * (size_t)tot) * 8];
      put_fr(dst, h.r[k]);
      for (int j = 0; j < n; j++) put_fr(dst + 8 * (size_t)(1 + j), h.a[(size_t)k * n + j]);
    }
  });
  jobs.assign(Bs * (size_t)m, MsmJob{});
  for (size_t q = 0; q < Bs * (size_t)m; q++) jobs[q] = MsmJob{(uint32_t)(q * tot), 0, tot};
  if ((rc = run_commit_jobs(Bs * (size_t)m * tot, Bs * (size_t)m, Bs * (size_t)m)) != MP_OK) return rc;

  // ---- round B: x; b_i = x^{perm[i]+1}; c_B[k] = com(chunk_k(b); s_k)
  parallel_for(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    memcpy(proof + L.cA, &h_pts[p * (size_t)m * 64], (size_t)m * 64);
    absorb_statement(h.fs, S, pk, decks + p * N * 128, out_decks + p * N * 128, N, proof + L.cA);
    h.x = h.fs.challenge();
    std::vector<fr> xp = h_powers(h.x, (int)N + 1);
    h.b.resize(N);
    for (size_t i = 0; i < N; i++) h.b[i] = xp[perms[p * N + i] + 1];
  });
  // (h_scal is shared: fill it after the transcripts so round A's data is no longer needed)
  parallel_for(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    for (int k = 0; k < m; k++) {
      uint32_t* dst = &h_scal[((p * m + k) * (size_t)tot) * 8];
      put_fr(dst, h.s[k]);
      for (int j = 0; j < n; j++) put_fr(dst + 8 * (size_t)(1 + j), h.b[(size_t)k * n + j]);
    }
  });
  if ((rc = run_commit_jobs(Bs * (size_t)m * tot, Bs * (size_t)m, Bs * (size_t)m)) != MP_OK) return rc;

  // ---- round C: y, z; product-argument rows, SVP and multi-exp first messages, diagonal ciphertexts
  h_scal.assign(Bs * (size_t)SC * 8, 0);
  std::vector<uint32_t> h_ct_scal(Bs * (N + n) * 8);
  parallel_for(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    memcpy(proof + L.cB, &h_pts[p * (size_t)m * 64], (size_t)m * 64);
    h.fs.begin(); h.fs.feed_label("shuffle_argument_b"); h.fs.feed_points64(proof + L.cB, m); h.fs.end();
    h.y = h.fs.challenge();
    h.z = h.fs.challenge();
    h.d.resize(N); h.t.resize(m);
    for (size_t i = 0; i < N; i++) h.d[i] = fr_sub(fr_add(fr_mul(h.y, h.a[i]), h.b[i]), h.z);
    for (int k = 0; k < m; k++) h.t[k] = fr_add(fr_mul(h.y, h.r[k]), h.s[k]);
    h.Bv.resize(N);
    for (int j = 0; j < n; j++) {
      fr acc = h.d[j];
      h.Bv[j] = acc;
      for (int k = 1; k < m; k++) { acc = fr_mul(acc, h.d[(size_t)k * n + j]); h.Bv[(size_t)k * n + j] = acc; }
    }
    h.col.assign(h.Bv.begin() + (size_t)(m - 1) * n, h.Bv.end());
    h.bk.resize(n);
    h.bk[0] = h.col[0];
    for (int i = 1; i < n; i++) h.bk[i] = fr_mul(h.bk[i - 1], h.col[i]);
    h.sv[0] = h.t[0];
    h.sv[m - 1] = h.s_prod;
    fr rho_star = fr_zero();
    for (size_t i = 0; i < N; i++) rho_star = fr_sub(rho_star, fr_mul(h_fr(rhos + (p * N + i) * 32), h.b[i]));
    h.me_tau[m] = rho_star;
    // G1 scalars of this proof: R rows of (n + 1), then 2m pairs (s_k, b_k), 2m singles tau_k, 2m pairs (b_k, tau_k)
    uint32_t* base = &h_scal[p * (size_t)SC * 8];
    auto row = [&](int rix, const fr& blind, const fr* vals, int len) {
      uint32_t* dst = base + (size_t)rix * tot * 8;
      put_fr(dst, blind);
      for (int j = 0; j < len; j++) put_fr(dst + 8 * (size_t)(1 + j), vals[j]);
    };
    for (int i = 0; i < m; i++) row(i, h.sv[i], &h.Bv[(size_t)i * n], n);
    row(m, h.sv_rd, h.sv_d.data(), n);
    std::vector<fr> v1((size_t)n - 1), v2((size_t)n - 1);
    for (int i = 0; i + 1 < n; i++) {
      v1[i] = fr_neg(fr_mul(h.sv_delta[i], h.sv_d[i + 1]));
      v2[i] = fr_sub(fr_sub(h.sv_delta[i + 1], fr_mul(h.col[i + 1], h.sv_delta[i])), fr_mul(h.bk[i], h.sv_d[i + 1]));
    }
    row(m + 1, h.sv_s1, v1.data(), n - 1);
    row(m + 2, h.sv_sx, v2.data(), n - 1);
    row(m + 3, h.me_r0, h.me_a0.data(), n);
    uint32_t* sm = base + (size_t)R * tot * 8;
    for (int k = 0; k < 2 * m; k++) {
      put_fr(sm + 8 * (size_t)(2 * k), h.me_s[k]);
      put_fr(sm + 8 * (size_t)(2 * k + 1), h.me_b[k]);
      put_fr(sm + 8 * (size_t)(4 * m + k), h.me_tau[k]);
      put_fr(sm + 8 * (size_t)(6 * m + 2 * k), h.me_b[k]);
      put_fr(sm + 8 * (size_t)(6 * m + 2 * k + 1), h.me_tau[k]);
    }
    // ciphertext scalars: rows a0 | b_1..b_m
    uint32_t* cs = &h_ct_scal[p * (N + n) * 8];
    for (int j = 0; j < n; j++) put_fr(cs + 8 * (size_t)j, h.me_a0[j]);
    for (size_t i = 0; i < N; i++) put_fr(cs + 8 * ((size_t)n + i), h.b[i]);
  });
  // diagonal ciphertext MSMs of every proof: one variable-base launch sequence
  {
    std::vector<MsmJob> diag(Bs * 2 * (size_t)m);
    for (size_t p = 0; p < Bs; p++)
      for (int k = 0; k < 2 * m; k++) {
        int i0 = std::max(1, m - k), i1 = std::min(m, 2 * m - k);
        diag[p * 2 * m + k] = MsmJob{(uint32_t)(p * (N + n) + (size_t)(k - m + i0) * n), (uint32_t)(p * N + (size_t)(i0 - 1) * n),
                                     (uint32_t)((size_t)(i1 - i0 + 1) * n)};
      }
    CK(cudaMemcpyAsync(d_ct_scal, h_ct_scal.data(), h_ct_scal.size() * 4, cudaMemcpyHostToDevice, st));
    CK(msm_run(ctx->ws, d_ct_scal, Bs * (N + n), d_ct_mont, 2, diag.data(), (int)diag.size(), msm_pick_window(N / 2 + 1), d_ct_out, st));
    launches += msm_last_launches(ctx->ws);
  }
  jobs.clear();
  for (size_t p = 0; p < Bs; p++) {
    const uint32_t b0 = (uint32_t)(p * SC), sm = b0 + (uint32_t)(R * tot);
    for (int k = 0; k < R; k++) jobs.push_back(MsmJob{b0 + (uint32_t)k * tot, 0, tot});
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{sm + 2 * (uint32_t)k, 0, 2});                                  // (h, g_1)
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{sm + (uint32_t)(4 * m + k), (uint32_t)(n + 1), 1});            // enc_g
    for (int k = 0; k < 2 * m; k++) jobs.push_back(MsmJob{sm + (uint32_t)(6 * m + 2 * k), (uint32_t)(n + 2), 2});        // (ghat, pk)
  }
  {
    CK(cudaMemcpyAsync(d_g1_scal, h_scal.data(), Bs * (size_t)SC * 32, cudaMemcpyHostToDevice, st));
    CK(msm_run(ctx->ws, d_g1_scal, Bs * (size_t)SC, S->d_tab_ck, 1, jobs.data(), (int)jobs.size(), S->tab_c, d_g1_out, st, 0, -1, nb));
    launches += msm_last_launches(ctx->ws);
    const uint64_t totalE = Bs * 4 * (uint64_t)m;
    k_combine_E_batch<<<(unsigned)((totalE + 63) / 64), 64, 0, st>>>(d_ct_out, d_g1_out, JC, (uint32_t)(R + 2 * m), (uint32_t)(R + 4 * m),
                                                                     2 * m, totalE);
    CK(cudaGetLastError());
    CK(xyzz_to_canonical(d_g1_out, (uint32_t*)d_canon, Bs * JC, st));
    CK(xyzz_to_canonical(d_ct_out, (uint32_t*)(d_canon + Bs * JC * 64), totalE, st));
    launches += 3;
    h_pts.resize((Bs * JC + totalE) * 64);
    CK(cudaMemcpyAsync(h_pts.data(), d_canon, h_pts.size(), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }

  // ---- round D: Hadamard challenges; zero-argument rows, diagonals and commitments
  h_scal.assign(Bs * (size_t)SD * 8, 0);
  parallel_for(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    const uint8_t* g1 = &h_pts[p * JC * 64];
    memcpy(proof + L.hB, g1, (size_t)m * 64);
    memcpy(proof + L.cb, g1 + 64 * (size_t)(m - 1), 64);
    memcpy(proof + L.svpts, g1 + 64 * (size_t)m, 3 * 64);
    memcpy(proof + L.mepts, g1 + 64 * (size_t)(m + 3), (size_t)(2 * m + 1) * 64);
    memcpy(proof + L.meE, &h_pts[(Bs * JC + p * 4 * (size_t)m) * 64], 4 * (size_t)m * 64);
    h.fs.begin(); h.fs.feed_label("hadamard_argument"); h.fs.feed_points64(proof + L.cb, 1); h.fs.feed_points64(proof + L.hB, m); h.fs.end();
    h.xh = h.fs.challenge();
    h.yh = h.fs.challenge();
    h.xhp = h_powers(h.xh, m);
    // A' = (a0 | d_2..d_m | -1),  B' = (xh^i Bv_i (i = 1..m-1) | sum_{i=1}^{m-1} xh^i Bv_{i+1} | b_{m+1})
    h.Az.assign((size_t)(m + 1) * n, fr_zero());
    h.Bz.assign((size_t)(m + 1) * n, fr_zero());
    const fr minus1 = fr_neg(fr_one());
    for (int j = 0; j < n; j++) {
      h.Az[j] = h.z_a0[j];
      for (int k = 1; k < m; k++) h.Az[(size_t)k * n + j] = h.d[(size_t)k * n + j];
      h.Az[(size_t)m * n + j] = minus1;
      fr last = fr_zero();
      for (int i = 1; i < m; i++) {
        h.Bz[(size_t)(i - 1) * n + j] = fr_mul(h.xhp[i], h.Bv[(size_t)(i - 1) * n + j]);
        last = fr_add(last, fr_mul(h.xhp[i], h.Bv[(size_t)i * n + j]));
      }
      h.Bz[(size_t)(m - 1) * n + j] = last;
      h.Bz[(size_t)m * n + j] = h.z_bm1[j];
    }
    std::vector<fr> yp((size_t)n), dk((size_t)2 * m + 1, fr_zero());
    fr acc = fr_one();
    for (int j = 0; j < n; j++) { acc = fr_mul(acc, h.yh); yp[j] = acc; }
    for (int i = 0; i <= m; i++)
      for (int jj = 0; jj <= m; jj++) {
        fr v = fr_zero();
        for (int tt = 0; tt < n; tt++)
          v = fr_add(v, fr_mul(fr_mul(h.Az[(size_t)i * n + tt], h.Bz[(size_t)jj * n + tt]), yp[tt]));
        int k = i + m - jj;
        dk[k] = fr_add(dk[k], v);
      }
    uint32_t* base = &h_scal[p * (size_t)SD * 8];
    put_fr(base, h.z_r0);
    for (int j = 0; j < n; j++) put_fr(base + 8 * (size_t)(1 + j), h.z_a0[j]);
    put_fr(base + 8 * (size_t)tot, h.z_sm1);
    for (int j = 0; j < n; j++) put_fr(base + 8 * (size_t)(tot + 1 + j), h.z_bm1[j]);
    for (int k = 0; k <= 2 * m; k++) {
      put_fr(base + 8 * (size_t)(2 * tot + 2 * k), h.z_t[k]);
      put_fr(base + 8 * (size_t)(2 * tot + 2 * k + 1), dk[k]);
    }
  });
  jobs.clear();
  for (size_t p = 0; p < Bs; p++) {
    const uint32_t b0 = (uint32_t)(p * SD);
    jobs.push_back(MsmJob{b0, 0, tot});
    jobs.push_back(MsmJob{b0 + tot, 0, tot});
    for (int k = 0; k <= 2 * m; k++) jobs.push_back(MsmJob{b0 + 2 * tot + 2 * (uint32_t)k, 0, 2});
  }
  if ((rc = run_commit_jobs(Bs * (size_t)SD, jobs.size(), Bs * (size_t)JD)) != MP_OK) return rc;

  // ---- remaining challenges and all responses (host)
  parallel_for(Bs, threads, [&](size_t p) {
    ProverHost& h = H[p];
    uint8_t* proof = proofs + p * plen;
    memcpy(proof + L.zpts, &h_pts[p * (size_t)JD * 64], (size_t)JD * 64);
    h.fs.begin(); h.fs.feed_label("zero_argument"); h.fs.feed_points64(proof + L.zpts, 2 * (size_t)m + 3); h.fs.end();
    const fr xz = h.fs.challenge();
    h.fs.begin(); h.fs.feed_label("single_value_product_argument"); h.fs.feed_points64(proof + L.svpts, 3); h.fs.end();
    const fr xs = h.fs.challenge();
    h.fs.begin(); h.fs.feed_label("multi_exponentiation_argument");
    h.fs.feed_points64(proof + L.mepts, 2 * (size_t)m + 1); h.fs.feed_points64(proof + L.meE, 4 * (size_t)m); h.fs.end();
    const fr xm = h.fs.challenge();
    const std::vector<fr> xzp = h_powers(xz, 2 * m + 1), xmp = h_powers(xm, 2 * m);
    for (int j = 0; j < n; j++) {
      fr za = fr_zero(), zb = fr_zero(), ma = h.me_a0[j];
      for (int i = 0; i <= m; i++) {
        za = fr_add(za, fr_mul(xzp[i], h.Az[(size_t)i * n + j]));
        zb = fr_add(zb, fr_mul(xzp[m - i], h.Bz[(size_t)i * n + j]));
      }
      for (int k = 1; k <= m; k++) ma = fr_add(ma, fr_mul(xmp[k], h.b[(size_t)(k - 1) * n + j]));
      h_fr_out(za, proof + L.za + 32 * (size_t)j);
      h_fr_out(zb, proof + L.zb + 32 * (size_t)j);
      h_fr_out(ma, proof + L.mea + 32 * (size_t)j);
      h_fr_out(fr_add(fr_mul(xs, h.col[j]), h.sv_d[j]), proof + L.sva + 32 * (size_t)j);
      h_fr_out(fr_add(fr_mul(xs, h.bk[j]), h.sv_delta[j]), proof + L.svb + 32 * (size_t)j);
    }
    std::vector<fr> rext((size_t)m + 1), sext((size_t)m + 1), xr((size_t)m + 1);
    rext[0] = h.z_r0;
    for (int i = 1; i < m; i++) rext[i] = h.t[i];
    rext[m] = fr_zero();
    for (int i = 1; i < m; i++) sext[i - 1] = fr_mul(h.xhp[i], h.sv[i - 1]);
    sext[m - 1] = h_dot(h.xhp.data() + 1, h.sv.data() + 1, m - 1);
    sext[m] = h.z_sm1;
    for (int j = 0; j <= m; j++) xr[j] = xzp[m - j];
    h_fr_out(h_dot(xzp.data(), rext.data(), m + 1), proof + L.zr);
    h_fr_out(h_dot(xr.data(), sext.data(), m + 1), proof + L.zs);
    h_fr_out(h_dot(xzp.data(), h.z_t.data(), 2 * m + 1), proof + L.zt);
    h_fr_out(fr_add(fr_mul(xs, h.s_prod), h.sv_rd), proof + L.svr);
    h_fr_out(fr_add(fr_mul(xs, h.sv_sx), h.sv_s1), proof + L.svs);
    std::vector<fr> mr((size_t)m + 1);
    mr[0] = h.me_r0;
    for (int j = 1; j <= m; j++) mr[j] = h.s[j - 1];
    h_fr_out(h_dot(xmp.data(), mr.data(), m + 1), proof + L.mer);
    h_fr_out(h_dot(xmp.data(), h.me_b.data(), 2 * m), proof + L.meb);
    h_fr_out(h_dot(xmp.data(), h.me_s.data(), 2 * m), proof + L.mes);
    h_fr_out(h_dot(xmp.data(), h.me_tau.data(), 2 * m), proof + L.metau);
  });
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "the public key is not a canonical point of the Stark curve");
  ctx->launches = launches;
  return MP_OK;
}

}  // namespace mp
