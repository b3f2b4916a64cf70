// Shuffle engine set-up and stand-alone primitives: parameters -> device key and fixed-base tables
// (reference DLCards::setup, mod.rs:105-121), remasking (remasking.rs:9-22 -> masking.rs:10-20) and
// Pedersen commitments.  See shuffle.cuh for the design notes.
#include "shuffle_internal.cuh"

namespace mp {

static constexpr int kPointWords = 2 * kFqLimbs;  // u32 words of a canonical point (x || y)

void shuffle_state_destroy(ShuffleState* s) { delete s; }
MsmWorkspace* shuffle_bulk_workspace(const mp_ctx* ctx) { return ctx && ctx->shuffle ? ctx->shuffle->bulk_ws : nullptr; }
int32_t shuffle_m(const mp_ctx* ctx) { return ctx && ctx->shuffle ? ctx->shuffle->m : 0; }
int32_t shuffle_n(const mp_ctx* ctx) { return ctx && ctx->shuffle ? ctx->shuffle->n : 0; }
bool shuffle_uses_small_deck_path(uint64_t n_cards) { return n_cards <= small_deck_max() && !getenv("MP_BATCH_WORKERS"); }

// ------------------------------------------------------------------------------------------
// kernels
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

cudaError_t commit_scalars_launch(const fr* d_rows, uint64_t stride, const fr* d_blinds, int count, int n, int len,
                                  uint32_t* d_out, cudaStream_t stream) {
  uint64_t total = (uint64_t)count * (n + 1);
  if (total == 0) return cudaSuccess;
  k_commit_scalars<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_rows, stride, d_blinds, count, n, len, d_out);
  return cudaGetLastError();
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
  if (g == 0 && !affine_on_curve(P)) atomicOr(bad, 1);   // bit 0: a point is off the curve
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
  if (src >= N) { atomicOr(bad, 2); return; }           // bit 1: permutation entry out of range
  affine card = affine_from_canonical(deck_canon + (src * 2 + comp) * kPointWords);
  if (!affine_on_curve(card)) atomicOr(bad, 1);
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
      for (int q = 0; q < (int)(sizeof(affine) / 16); q++) dst[q] = __ldg(s + q);
      xyzz_madd(acc, e);
    }
  }
  affine r = xyzz_to_affine(acc);
  uint32_t w[kPointWords];
  if (affine_is_identity(r)) {
#pragma unroll
    for (int q = 0; q < kPointWords; q++) w[q] = 0;
  } else {
    affine_to_canonical(r, w);
  }
  uint4* o = reinterpret_cast<uint4*>(out_canon + g * kPointWords);
#pragma unroll
  for (int q = 0; q < kPointWords / 4; q++) o[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
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

int32_t shuffle_ensure_pk_table(mp_ctx* ctx, const uint8_t* pk) {
  ShuffleState* S = ctx->shuffle;
  if (S->tab_pk_valid && memcmp(S->tab_pk, pk, kPointBytes) == 0) return MP_OK;
  uint8_t* d_pk = (uint8_t*)ctx->scratch(sSmallUp, 256);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_pk); NEED(d_bad);
  S->tab_pk_valid = false;  // the table is being overwritten: valid again only once the device has accepted pk
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  CK(cudaMemcpyAsync(d_pk, pk, kPointBytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(build_table(S, 1, d_pk, d_bad, ctx->stream));
  ctx->launches += 1;
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(stream_wait(ctx, ctx->stream));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "the public key is not on the Stark curve");
  memcpy(S->tab_pk, pk, kPointBytes);
  S->tab_pk_valid = true;
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// set-up
// ------------------------------------------------------------------------------------------
int32_t run_g1_jobs(mp_ctx* ctx, const TermList& tl, xyzz** d_out_ret, int* d_bad, cudaStream_t stream, MsmWorkspace* ws) {
  if (!stream) stream = ctx->stream;
  if (!ws) ws = ctx->ws;
  uint32_t T = tl.count();
  int J = (int)tl.jobs.size();
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sG1Canon, (size_t)T * kPointBytes + 64);
  affine* d_mont = (affine*)ctx->scratch(sG1Mont, (size_t)T * sizeof(affine) + 64);
  uint32_t* d_scal = (uint32_t*)ctx->scratch(sG1Scal, (size_t)T * 32 + 64);
  xyzz* d_out = (xyzz*)ctx->scratch(sG1Out, (size_t)J * sizeof(xyzz) + 64);
  NEED(d_canon); NEED(d_mont); NEED(d_scal); NEED(d_out);
  CK(cudaMemcpyAsync(d_canon, tl.pts.data(), (size_t)T * kPointBytes, cudaMemcpyHostToDevice, stream));
  CK(cudaMemcpyAsync(d_scal, tl.scal.data(), (size_t)T * 32, cudaMemcpyHostToDevice, stream));
  CK(points_to_mont((const uint32_t*)d_canon, d_mont, T, d_bad, stream));
  ctx->launches += 1;
  int c = msm_pick_window(J ? T / J : 1, J);
  CK(msm_run(ws, d_scal, T, d_mont, 1, tl.jobs.data(), J, c, d_out, stream));
  ctx->launches += msm_last_launches(ws);
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
  S->ck64.resize((size_t)(n + 1) * kPointBytes);
  memcpy(S->ck64.data(), ck_h, kPointBytes);
  memcpy(S->ck64.data() + kPointBytes, ck_g, (size_t)n * kPointBytes);
  memcpy(S->enc_g, enc_g, kPointBytes);
  memcpy(S->ghat, ghat, kPointBytes);
  if (S->d_ck) cudaFree(S->d_ck);
  S->d_ck = nullptr;
  CK(cudaMalloc(&S->d_ck, sizeof(affine) * (size_t)(n + 4)));
  if (!S->ev) CK(cudaEventCreateWithFlags(&S->ev, cudaEventDisableTiming));
  if (!S->aux) {
    CK(cudaStreamCreateWithFlags(&S->aux, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&S->ev_join, cudaEventDisableTiming));
    S->aux_ws = msm_workspace_create();
  }
  if (!S->bulk) {
    CK(cudaStreamCreateWithFlags(&S->bulk, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&S->ev_bulk_go, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&S->ev_bulk_done, cudaEventDisableTiming));
    S->bulk_ws = msm_workspace_create();
  }
  // validate every parameter point and compute gsum = sum g_j with one MSM of unit scalars
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  TermList tl;
  for (int j = 1; j <= n; j++) tl.term(S->ck64.data() + kPointBytes * (size_t)j, fr_one());
  tl.close_job();
  tl.term(ck_h, fr_one()); tl.term(enc_g, fr_one()); tl.term(ghat, fr_one());  // validation only
  tl.close_job();
  xyzz* d_out = nullptr;
  int32_t st = run_g1_jobs(ctx, tl, &d_out, d_bad);
  if (st != MP_OK) return st;
  uint8_t* d_res = (uint8_t*)ctx->scratch(sCanonOut, kPointBytes + 64);
  NEED(d_res);
  CK(xyzz_to_canonical(d_out, (uint32_t*)d_res, 1, ctx->stream));
  ctx->launches += 1;
  // Montgomery copy of the commit key for the prover's commitment jobs
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sG1Canon, (size_t)(n + 3) * kPointBytes);
  NEED(d_canon);
  CK(cudaMemcpyAsync(d_canon, S->ck64.data(), (size_t)(n + 1) * kPointBytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_canon + (size_t)(n + 1) * kPointBytes, enc_g, kPointBytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_canon + (size_t)(n + 2) * kPointBytes, ghat, kPointBytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(points_to_mont((const uint32_t*)d_canon, S->d_ck, (uint64_t)n + 3, d_bad, ctx->stream));
  CK(build_table(S, 0, d_canon + (size_t)(n + 1) * kPointBytes, d_bad, ctx->stream));  // remask table of g
  S->tab_pk_valid = false;
  // fixed-base table for the commitment jobs (the pk column is filled per call)
  // One bucket set per job in table mode, so the window is a compromise over the prover's job mix: per proof
  // 3m + 6 row commitments of n + 1 terms and 8m + 1 one- / two-term jobs (c_B_k, Enc(b_k ghat; tau_k), c_D_k), each
  // of which pays the whole 2^(c-1)-bucket reduction.  Sizing for the rows alone (c = 11 at n = 512) made the
  // small jobs 1.2 ms of bucket sweeps per 2^16-card proof; the job-weighted length gives c = 9.
  {
    const uint64_t rows = 3 * (uint64_t)m + 6, small = 8 * (uint64_t)m + 1;
    S->tab_c = msm_pick_table_window((rows * ((uint64_t)n + 1) + small * 2) / (rows + small));
  }
  if (S->d_tab_ck) cudaFree(S->d_tab_ck);
  S->d_tab_ck = nullptr;
  CK(cudaMalloc(&S->d_tab_ck, sizeof(affine) * (size_t)msm_num_windows(S->tab_c) * (size_t)(n + 4)));
  CK(cudaMemsetAsync(S->d_tab_ck, 0, sizeof(affine) * (size_t)msm_num_windows(S->tab_c) * (size_t)(n + 4), ctx->stream));
  CK(msm_build_table(ctx->ws, S->d_ck, (uint32_t)(n + 4), 0, (uint32_t)(n + 3), S->tab_c, S->d_tab_ck, ctx->stream));
  S->ck_pk_valid = false;
  ctx->launches += 4;
  int bad = 0;
  CK(cudaMemcpyAsync(S->gsum, d_res, kPointBytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(stream_wait(ctx, ctx->stream));
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a parameter point is not a canonical point of the curve");
  S->m = m;
  S->n = n;
  S->params_gen++;
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// remask + commitments (stand-alone entry points; the prover reuses the pieces)
// ------------------------------------------------------------------------------------------
int32_t shuffle_remask(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm, const uint8_t* rho,
                       uint64_t N, uint8_t* out_deck, const void* deck_src, const void** d_out_ret, Transcript* fs_head) {
  NvtxRange nvtx("shuffle_remask");
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
  uint8_t* d_out = (uint8_t*)ctx->scratch(sCtCanon, (N + 2) * kCtBytes);
  uint32_t* d_perm = (uint32_t*)ctx->scratch(sPerm, N * 4);
  uint8_t* d_rho = (uint8_t*)ctx->scratch(sRho, N * 32 + 64);
  uint8_t* d_pk = (uint8_t*)ctx->scratch(sSmallUp, 256);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_deck); NEED(d_out); NEED(d_perm); NEED(d_rho); NEED(d_pk); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  // the pk table is cached across calls; it is marked valid only AFTER the device has validated pk (flag clean
  // at the synchronisation below) -- any early return in between leaves the cache invalid
  const bool rebuild_pk = !S->tab_pk_valid || memcmp(S->tab_pk, pk, kPointBytes) != 0;
  if (rebuild_pk) {
    S->tab_pk_valid = false;
    CK(cudaMemcpyAsync(d_pk, pk, kPointBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(build_table(S, 1, d_pk, d_bad, ctx->stream));
    ctx->launches += 1;
  }
  CK(cudaMemcpyAsync(d_deck, deck_src, N * kCtBytes, cudaMemcpyDefault, ctx->stream));
  CK(cudaMemcpyAsync(d_perm, perm, N * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_rho, rho, N * 32, cudaMemcpyHostToDevice, ctx->stream));
  k_remask<<<(unsigned)((2 * N + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t*)d_deck, d_perm, (const uint32_t*)d_rho,
                                                                     S->d_tab, N, (uint32_t*)d_out, d_bad);
  CK(cudaGetLastError());
  ctx->launches += 1;
  int bad = 0;
  CK(cudaMemcpyAsync(out_deck, d_out, N * kCtBytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (fs_head) absorb_statement_head(*fs_head, S, pk, deck, N);  // host hashing overlaps the copies and the kernel
  CK(stream_wait(ctx, ctx->stream));
  // distinct bits (atomicOr): an out-of-range permutation entry can no longer hide an off-curve key or card
  if (bad & 1) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a deck point or the public key is not on the Stark curve");
  if (rebuild_pk) {  // pk validated by k_build_table: the table may be reused by later calls
    memcpy(S->tab_pk, pk, kPointBytes);
    S->tab_pk_valid = true;
  }
  if (bad & 2) return ctx->fail(MP_ERR_INVALID_ARG, "permutation entry out of range");
  if (d_out_ret) *d_out_ret = d_out;
  return MP_OK;
}

// commitments of `count` rows that live on the device (Montgomery), result XYZZ on the device
int32_t commit_rows_device(mp_ctx* ctx, const fr* d_rows, uint64_t stride, const fr* d_blinds, int count, int len,
                                  uint32_t* d_scal, xyzz* d_out) {
  ShuffleState* S = ctx->shuffle;
  const int n = S->n;
  uint64_t total = (uint64_t)count * (n + 1);
  CK(commit_scalars_launch(d_rows, stride, d_blinds, count, n, len, d_scal, ctx->stream));
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
  uint8_t* d_canon = (uint8_t*)ctx->scratch(sCanonOut, k * kPointBytes);
  NEED(d_in); NEED(d_rows); NEED(d_scal); NEED(d_res); NEED(d_canon);
  if (len) CK(cudaMemcpyAsync(d_in, values, k * len * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_in + k * len * 8, blinds, k * 32, cudaMemcpyHostToDevice, ctx->stream));
  CK(fr_from_canonical_vec(d_in, d_rows, k * len + k, ctx->stream));
  ctx->launches += 1;
  int32_t st = commit_rows_device(ctx, d_rows, len, d_rows + k * len, (int)k, (int)len, d_scal, d_res);
  if (st != MP_OK) return st;
  CK(xyzz_to_canonical(d_res, (uint32_t*)d_canon, k, ctx->stream));
  ctx->launches += 1;
  CK(cudaMemcpyAsync(out, d_canon, k * kPointBytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(stream_wait(ctx, ctx->stream));
  return MP_OK;
}

}  // namespace mp
