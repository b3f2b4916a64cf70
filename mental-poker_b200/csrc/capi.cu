// extern "C" surface of libmpshuffle.so (declared in include/mpshuffle.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mpshuffle.h"
#include "comm.cuh"
#include "ctx.cuh"
#include "msm.cuh"
#include "shuffle.cuh"
#include "wire_host.hpp"
#include "transcript.hpp"
#include <chrono>

using namespace mp;

// ------------------------------------------------------------------------------------------
// context plumbing
// ------------------------------------------------------------------------------------------
extern "C" int32_t mp_ctx_create(mp_ctx** out, int32_t device) { return ctx_create(out, device); }
extern "C" void mp_ctx_destroy(mp_ctx* ctx) { ctx_destroy(ctx); }

extern "C" void* mp_ctx_stream(mp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int32_t mp_ctx_sync(mp_ctx* ctx) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "mp_ctx_sync");
  return MP_OK;
}
extern "C" const char* mp_last_error_string(mp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" const char* mp_verify_status_string(int32_t status) {
  switch (status) {
    case MP_OK: return "ok";
    case MP_VERIFY_HADAMARD: return "Hadamard Product (5.1)";
    case MP_VERIFY_ZERO: return "Zero Argument (5.2)";
    case MP_VERIFY_SVP: return "Single Value Product (5.3)";
    case MP_VERIFY_MULTIEXP: return "Multi Exponentiation (4)";
    case MP_VERIFY_CHAUM_PEDERSEN: return "Chaum-Pedersen";
    case MP_VERIFY_SCHNORR: return "Schnorr Identification";
    case MP_VERIFY_MALFORMED: return "malformed input (point off the curve or non-canonical bytes)";
    default: return "unknown";
  }
}
extern "C" int32_t mp_last_kernel_launches(mp_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t mp_last_msm_ec_adds(mp_ctx* ctx) { return ctx ? ctx->last_ec_adds : 0; }
extern "C" int32_t mp_last_msm_window(mp_ctx* ctx) { return ctx ? ctx->last_window : 0; }

// ------------------------------------------------------------------------------------------
// MSM entry points
// ------------------------------------------------------------------------------------------
static uint64_t scheduled_ec_adds(uint64_t n, int c, int ncomp) {
  uint64_t W = (253 + c - 1) / c, B = 1ull << (c - 1);
  return (uint64_t)ncomp * (W * (n + 2 * B) + W * (uint64_t)c);
}

// d_points canonical (n*ncomp points), d_scalars canonical, d_out canonical (ncomp points)
static int32_t msm_device_common(mp_ctx* ctx, const void* d_points, const void* d_scalars,
                                 uint64_t n, int ncomp, int32_t window_bits, void* d_out,
                                 int w_begin = 0, int w_count = -1) {
  if (!ctx || (!d_points && n) || (!d_scalars && n) || !d_out) return MP_ERR_INVALID_ARG;
  if (n >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "MSM size %llu too large", (unsigned long long)n);
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  int c = window_bits > 0 ? window_bits : msm_pick_window(n);
  if (c < 2 || c > 16) return ctx->fail(MP_ERR_INVALID_ARG, "window_bits %d out of range [2,16]", c);
  ctx->last_window = c;
  ctx->last_ec_adds = scheduled_ec_adds(n, c, ncomp) * (uint64_t)(w_count < 0 ? msm_num_windows(c) - w_begin : w_count) / msm_num_windows(c);
  if (w_begin < 0 || (w_count >= 0 && w_begin + w_count > msm_num_windows(c)) || w_count == 0)
    return ctx->fail(MP_ERR_INVALID_ARG, "window range [%d, +%d) outside [0, %d)", w_begin, w_count, msm_num_windows(c));
  affine* mont = (affine*)ctx->scratch(mp_ctx::kSlotPointsMont, sizeof(affine) * n * ncomp);
  xyzz* res = (xyzz*)ctx->scratch(mp_ctx::kSlotMsmOut, sizeof(xyzz) * ncomp);
  int* bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  if (!mont || !res || !bad) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaError_t e;
  if ((e = cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "memset");
  if ((e = points_to_mont((const uint32_t*)d_points, mont, n * ncomp, bad, ctx->stream)) != cudaSuccess)
    return ctx->cuda_fail(e, "points_to_mont");
  ctx->launches += n ? 1 : 0;
  MsmJob job{0, 0, (uint32_t)n};
  if ((e = msm_run(ctx->ws, (const uint32_t*)d_scalars, n, mont, ncomp, &job, 1, c, res, ctx->stream, w_begin, w_count)) != cudaSuccess)
    return ctx->cuda_fail(e, "msm_run");
  ctx->launches += msm_last_launches(ctx->ws);
  if ((e = xyzz_to_canonical(res, (uint32_t*)d_out, ncomp, ctx->stream)) != cudaSuccess)
    return ctx->cuda_fail(e, "xyzz_to_canonical");
  ctx->launches += 1;
  return MP_OK;
}

static int32_t msm_host_common(mp_ctx* ctx, const uint8_t* points, const uint8_t* scalars, uint64_t n,
                               int ncomp, int32_t window_bits, uint8_t* out) {
  if (!ctx || (!points && n) || (!scalars && n) || !out) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  size_t pbytes = (size_t)n * 64 * ncomp, sbytes = (size_t)n * 32;
  uint8_t* d_in = (uint8_t*)ctx->scratch(mp_ctx::kSlotStageIn, pbytes + sbytes + 256);
  uint8_t* d_out = (uint8_t*)ctx->scratch(mp_ctx::kSlotStageOut, 64 * ncomp);
  if (!d_in || !d_out) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaError_t e;
  if (n) {
    if ((e = cudaMemcpyAsync(d_in, points, pbytes, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
      return ctx->cuda_fail(e, "H2D points");
    if ((e = cudaMemcpyAsync(d_in + pbytes, scalars, sbytes, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
      return ctx->cuda_fail(e, "H2D scalars");
  }
  int32_t st = msm_device_common(ctx, d_in, d_in + pbytes, n, ncomp, window_bits, d_out);
  if (st != MP_OK) return st;
  int bad = 0;
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  if ((e = cudaMemcpyAsync(out, d_out, 64 * ncomp, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
    return ctx->cuda_fail(e, "D2H result");
  if ((e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
    return ctx->cuda_fail(e, "D2H flag");
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "MSM execution");
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "an input point is not on the Stark curve");
  return MP_OK;
}

extern "C" int32_t mp_msm_g1(mp_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, uint64_t n,
                             int32_t window_bits, uint8_t* out) {
  return msm_host_common(ctx, bases, scalars, n, 1, window_bits, out);
}
extern "C" int32_t mp_ct_msm(mp_ctx* ctx, const uint8_t* deck, const uint8_t* scalars, uint64_t n,
                             int32_t window_bits, uint8_t* out) {
  return msm_host_common(ctx, deck, scalars, n, 2, window_bits, out);
}
extern "C" int32_t mp_msm_g1_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                    int32_t window_bits, void* d_out) {
  return msm_device_common(ctx, d_bases, d_scalars, n, 1, window_bits, d_out);
}
extern "C" int32_t mp_ct_msm_device(mp_ctx* ctx, const void* d_deck, const void* d_scalars, uint64_t n,
                                    int32_t window_bits, void* d_out) {
  return msm_device_common(ctx, d_deck, d_scalars, n, 2, window_bits, d_out);
}

// K2 of SURVEY.md section 2b / `mp_msm_batch_shared_bases` of section 8(b): a batch of MSMs over ONE point array
// and ONE scalar array -- job j = sum_{t < len_j} scalars[scalar_off_j + t] * points[point_off_j + t] -- in one
// launch sequence (one digit pass, one sort, one accumulation for all jobs).  The multi-exponentiation argument's
// diagonal products are m(m+1) such jobs over the rows of the shuffled deck (reference call site mod.rs:409-415
// -> MultiExponentiationArgument).  jobs: njobs x (scalar_off, point_off, len) as uint32.  out: njobs * ncomp points.
extern "C" int32_t mp_msm_jobs(mp_ctx* ctx, const uint8_t* points, uint64_t n_points, int32_t ncomp, const uint8_t* scalars,
                               uint64_t n_scalars, const uint32_t* jobs, uint64_t njobs, int32_t window_bits, uint8_t* out) {
  if (!ctx || !jobs || !out || (!points && n_points) || (!scalars && n_scalars)) return MP_ERR_INVALID_ARG;
  if (ncomp != 1 && ncomp != 2) return ctx->fail(MP_ERR_INVALID_ARG, "ncomp must be 1 (G1) or 2 (ciphertexts)");
  if (njobs == 0) return MP_OK;
  if (njobs >= (1u << 24) || n_points >= (1ull << 31) || n_scalars >= (1ull << 31))
    return ctx->fail(MP_ERR_INVALID_ARG, "batch too large");
  static_assert(sizeof(MsmJob) == 12, "MsmJob is three uint32");
  std::vector<MsmJob> h_jobs(njobs);
  uint64_t total = 0;
  for (uint64_t j = 0; j < njobs; j++) {
    h_jobs[j] = MsmJob{jobs[3 * j], jobs[3 * j + 1], jobs[3 * j + 2]};
    if ((uint64_t)h_jobs[j].scalar_off + h_jobs[j].len > n_scalars || (uint64_t)h_jobs[j].point_off + h_jobs[j].len > n_points)
      return ctx->fail(MP_ERR_INVALID_ARG, "job %llu reaches outside the arrays", (unsigned long long)j);
    total += h_jobs[j].len;
  }
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int c = window_bits > 0 ? window_bits : msm_pick_window(total / njobs, njobs);
  if (c < 2 || c > 16) return ctx->fail(MP_ERR_INVALID_ARG, "window_bits %d out of range [2,16]", c);
  ctx->last_window = c;
  const size_t pbytes = (size_t)n_points * 64 * ncomp, sbytes = (size_t)n_scalars * 32;
  uint8_t* d_in = (uint8_t*)ctx->scratch(mp_ctx::kSlotStageIn, pbytes + sbytes + 256);
  affine* mont = (affine*)ctx->scratch(mp_ctx::kSlotPointsMont, sizeof(affine) * n_points * ncomp);
  xyzz* res = (xyzz*)ctx->scratch(mp_ctx::kSlotMsmOut, sizeof(xyzz) * njobs * ncomp);
  uint8_t* d_out = (uint8_t*)ctx->scratch(mp_ctx::kSlotStageOut, 64 * njobs * ncomp);
  int* bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  if (!d_in || !mont || !res || !d_out || !bad) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaError_t e;
  if (pbytes && (e = cudaMemcpyAsync(d_in, points, pbytes, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "H2D points");
  if (sbytes && (e = cudaMemcpyAsync(d_in + pbytes, scalars, sbytes, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "H2D scalars");
  if ((e = cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "memset");
  if ((e = points_to_mont((const uint32_t*)d_in, mont, n_points * ncomp, bad, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "points_to_mont");
  ctx->launches += n_points ? 1 : 0;
  if ((e = msm_run(ctx->ws, (const uint32_t*)(d_in + pbytes), n_scalars, mont, ncomp, h_jobs.data(), (int)njobs, c, res, ctx->stream)) != cudaSuccess)
    return ctx->cuda_fail(e, "msm_run (jobs)");
  ctx->launches += msm_last_launches(ctx->ws);
  if ((e = xyzz_to_canonical(res, (uint32_t*)d_out, njobs * ncomp, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "xyzz_to_canonical");
  ctx->launches += 1;
  int h_bad = 0;
  if ((e = cudaMemcpyAsync(out, d_out, 64 * njobs * ncomp, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "D2H results");
  if ((e = cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "D2H flag");
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "MSM jobs");
  if (h_bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "an input point is not a canonical point of the Stark curve");
  ctx->last_ec_adds = 0;
  for (auto& j : h_jobs) ctx->last_ec_adds += scheduled_ec_adds(j.len, c, ncomp);
  return MP_OK;
}

extern "C" int32_t mp_profile_enable(mp_ctx* ctx, int32_t on) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  msm_profile_enable(ctx->ws, on != 0);
  if (MsmWorkspace* b = shuffle_bulk_workspace(ctx)) msm_profile_enable(b, on != 0);
  return MP_OK;
}
extern "C" int32_t mp_profile_collect(mp_ctx* ctx, double* accumulate_ms, uint64_t* bucket_adds, uint64_t* launches) {
  if (!ctx || !accumulate_ms || !bucket_adds || !launches) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaError_t e = msm_profile_collect(ctx->ws, accumulate_ms, bucket_adds, launches);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "mp_profile_collect");
  if (MsmWorkspace* b = shuffle_bulk_workspace(ctx)) {  // the prover's diagonal products run on their own workspace
    double ms = 0; uint64_t adds = 0, n = 0;
    e = msm_profile_collect(b, &ms, &adds, &n);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "mp_profile_collect");
    *accumulate_ms += ms; *bucket_adds += adds; *launches += n;
  }
  return MP_OK;
}

extern "C" int32_t mp_profile_collect_dominant(mp_ctx* ctx, double* ms, uint64_t* bucket_adds, uint64_t* launches) {
  if (!ctx || !ms || !bucket_adds || !launches) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  double all_ms; uint64_t all_adds, all_launches;
  // the dominant launches are the prover's diagonal products when a proof ran (bulk workspace)
  MsmWorkspace* b = shuffle_bulk_workspace(ctx);
  cudaError_t e = msm_profile_collect(ctx->ws, &all_ms, &all_adds, &all_launches, ms, bucket_adds, launches);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "mp_profile_collect_dominant");
  if (b) {
    double bms = 0; uint64_t badds = 0, bn = 0;
    e = msm_profile_collect(b, &all_ms, &all_adds, &all_launches, &bms, &badds, &bn);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "mp_profile_collect_dominant");
    if (bn && badds / bn >= (*launches ? *bucket_adds / *launches : 0)) { *ms = bms; *bucket_adds = badds; *launches = bn; }
  }
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// shuffle protocol entry points (bodies in shuffle_*.cu)
// ------------------------------------------------------------------------------------------
extern "C" int32_t mp_ctx_set_params(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g, const uint8_t* ck_g,
                                     const uint8_t* ck_h, const uint8_t* ghat) {
  return shuffle_set_params(ctx, m, n, enc_g, ck_g, ck_h, ghat);
}
extern "C" int32_t mp_params_m(mp_ctx* ctx) { return shuffle_m(ctx); }
extern "C" int32_t mp_params_n(mp_ctx* ctx) { return shuffle_n(ctx); }
extern "C" uint64_t mp_proof_len(int32_t m, int32_t n) { return shuffle_proof_len(m, n); }
extern "C" uint64_t mp_prover_randomness_len(int32_t m, int32_t n) { return shuffle_randomness_len(m, n); }
extern "C" int32_t mp_remask_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                   const uint8_t* rho, uint64_t n_cards, uint8_t* out_deck) {
  return shuffle_remask(ctx, pk, deck, perm, rho, n_cards, out_deck);
}
extern "C" int32_t mp_pedersen_commit_batch(mp_ctx* ctx, const uint8_t* values, const uint8_t* blinds, uint64_t k,
                                            uint64_t len, uint8_t* out) {
  return shuffle_commit_batch(ctx, values, blinds, k, len, out);
}
extern "C" int32_t mp_shuffle_prove(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* shuffled_deck,
                                    const uint32_t* perm, const uint8_t* rho, const uint8_t* randomness,
                                    uint8_t* proof_out) {
  return shuffle_prove(ctx, pk, deck, shuffled_deck, perm, rho, randomness, proof_out);
}
static int32_t shuffle_and_remask_common(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                         const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck,
                                         uint8_t* proof_out, const void* d_deck) {
  if (!ctx || !ctx->shuffle) return ctx ? ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called") : MP_ERR_INVALID_ARG;
  uint64_t n_cards = (uint64_t)mp_params_m(ctx) * mp_params_n(ctx);
  if (!d_deck && shuffle_uses_small_deck_path(n_cards))  // small decks: the lockstep prover with a batch of one
    return shuffle_prove_batch(ctx, pk, deck, perm, rho, randomness, 1, out_deck, proof_out, 1);
  const void* d_shuffled = nullptr;
  Transcript fs;  // its statement absorb starts while the remask kernel runs
  int32_t st = shuffle_remask(ctx, pk, deck, perm, rho, n_cards, out_deck, d_deck, &d_shuffled, &fs);
  if (st != MP_OK) return st;
  int launches = ctx->launches;
  st = shuffle_prove(ctx, pk, deck, out_deck, perm, rho, randomness, proof_out, d_shuffled, &fs);
  ctx->launches += launches;
  return st;
}
extern "C" int32_t mp_shuffle_and_remask(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                         const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck,
                                         uint8_t* proof_out) {
  return shuffle_and_remask_common(ctx, pk, deck, perm, rho, randomness, out_deck, proof_out, nullptr);
}
extern "C" int32_t mp_shuffle_and_remask_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                                  const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck,
                                                  uint8_t* proof_out, const void* d_deck) {
  return shuffle_and_remask_common(ctx, pk, deck, perm, rho, randomness, out_deck, proof_out, d_deck);
}
extern "C" int32_t mp_shuffle_verify(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* shuffled_deck,
                                     const uint8_t* proof) {
  return shuffle_verify(ctx, pk, deck, shuffled_deck, proof);
}

extern "C" int32_t mp_shuffle_and_remask_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                                               const uint8_t* rhos, const uint8_t* randomness, uint64_t batch,
                                               uint8_t* out_decks, uint8_t* proofs, int32_t host_threads) {
  return shuffle_prove_batch(ctx, pk, decks, perms, rhos, randomness, batch, out_decks, proofs, host_threads);
}
extern "C" int32_t mp_shuffle_verify_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks,
                                           const uint8_t* shuffled_decks, const uint8_t* proofs, uint64_t batch,
                                           int32_t* statuses, int32_t host_threads) {
  return shuffle_verify_batch(ctx, pk, decks, shuffled_decks, proofs, batch, statuses, host_threads);
}
extern "C" int32_t mp_shuffle_and_remask_batch_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                                                        const uint8_t* rhos, const uint8_t* randomness, uint64_t batch,
                                                        uint8_t* out_decks, uint8_t* proofs, int32_t host_threads,
                                                        const void* d_decks) {
  return shuffle_prove_batch(ctx, pk, decks, perms, rhos, randomness, batch, out_decks, proofs, host_threads, d_decks);
}
extern "C" int32_t mp_shuffle_verify_batch_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks,
                                                    const uint8_t* shuffled_decks, const uint8_t* proofs, uint64_t batch,
                                                    int32_t* statuses, int32_t host_threads, const void* d_decks,
                                                    const void* d_shuffled_decks) {
  return shuffle_verify_batch(ctx, pk, decks, shuffled_decks, proofs, batch, statuses, host_threads, d_decks, d_shuffled_decks);
}
extern "C" int32_t mp_shuffle_verify_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck,
                                              const uint8_t* shuffled_deck, const uint8_t* proof, const void* d_deck,
                                              const void* d_shuffled_deck) {
  return shuffle_verify(ctx, pk, deck, shuffled_deck, proof, d_deck, d_shuffled_deck);
}
extern "C" int32_t mp_shuffle_prove_resident(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck,
                                             const uint8_t* shuffled_deck, const uint32_t* perm, const uint8_t* rho,
                                             const uint8_t* randomness, uint8_t* proof_out, const void* d_shuffled_deck) {
  return shuffle_prove(ctx, pk, deck, shuffled_deck, perm, rho, randomness, proof_out, d_shuffled_deck);
}

// ---- batched sigma protocols (bodies in sigma.cu)
extern "C" int32_t mp_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* r,
                                 const uint8_t* omega, uint64_t n, uint8_t* out_masked, uint8_t* out_proofs, int32_t host_threads) {
  return sigma_mask_batch(ctx, shared_key, cards, r, omega, n, out_masked, out_proofs, host_threads);
}
extern "C" int32_t mp_verify_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* masked,
                                        const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  return sigma_verify_mask_batch(ctx, shared_key, cards, masked, proofs, n, statuses, host_threads);
}
extern "C" int32_t mp_remask_prove_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* alpha,
                                         const uint8_t* omega, uint64_t n, uint8_t* out_deck, uint8_t* out_proofs,
                                         int32_t host_threads) {
  return sigma_remask_prove_batch(ctx, shared_key, deck, alpha, omega, n, out_deck, out_proofs, host_threads);
}
extern "C" int32_t mp_verify_remask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* remasked,
                                          const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  return sigma_verify_remask_batch(ctx, shared_key, deck, remasked, proofs, n, statuses, host_threads);
}
extern "C" int32_t mp_reveal_batch(mp_ctx* ctx, const uint8_t* sk, const uint8_t* pk, const uint8_t* masked, const uint8_t* omega,
                                   uint64_t n, uint8_t* out_tokens, uint8_t* out_proofs, int32_t host_threads) {
  return sigma_reveal_batch(ctx, sk, pk, masked, omega, n, out_tokens, out_proofs, host_threads);
}
extern "C" int32_t mp_verify_reveal_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* tokens, const uint8_t* masked,
                                          const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  return sigma_verify_reveal_batch(ctx, pk, tokens, masked, proofs, n, statuses, host_threads);
}
extern "C" int32_t mp_key_ownership_prove_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* sks, const uint8_t* infos,
                                                const uint64_t* info_offsets, const uint8_t* omega, uint64_t n,
                                                uint8_t* out_proofs, int32_t host_threads) {
  return sigma_key_ownership_prove_batch(ctx, pks, sks, infos, info_offsets, omega, n, out_proofs, host_threads);
}
extern "C" int32_t mp_key_ownership_verify_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* infos,
                                                 const uint64_t* info_offsets, const uint8_t* proofs, uint64_t n,
                                                 int32_t* statuses, int32_t host_threads) {
  return sigma_key_ownership_verify_batch(ctx, pks, infos, info_offsets, proofs, n, statuses, host_threads);
}

// ---- wire format (host half in wire_host.hpp, device half in wire.cu)
extern "C" int32_t mp_points_compress(const uint8_t* points, uint64_t n, uint8_t* out) {
  if (n && (!points || !out)) return MP_ERR_INVALID_ARG;
  for (uint64_t i = 0; i < n; i++) wire_compress_point(points + 64 * i, out + 32 * i);
  return MP_OK;
}
extern "C" int32_t mp_points_decompress(mp_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, int32_t* statuses) {
  return wire_points_decompress(ctx, in, n, out, statuses);
}
extern "C" uint64_t mp_deck_serialized_len(uint64_t n_cards) { return wire_deck_len(n_cards); }
extern "C" int32_t mp_deck_serialize(const uint8_t* deck, uint64_t n_cards, uint8_t* out) {
  if (!out || (n_cards && !deck)) return MP_ERR_INVALID_ARG;
  memcpy(out, &n_cards, 8);
  for (uint64_t i = 0; i < 2 * n_cards; i++) wire_compress_point(deck + 64 * i, out + 8 + 32 * i);
  return MP_OK;
}
extern "C" int32_t mp_deck_deserialize(mp_ctx* ctx, const uint8_t* in, uint64_t in_len, uint8_t* out_deck, uint64_t* n_cards) {
  return wire_deck_deserialize(ctx, in, in_len, out_deck, n_cards);
}
extern "C" uint64_t mp_proof_serialized_len(int32_t m, int32_t n) { return wire_proof_len(m, n); }
extern "C" int32_t mp_proof_serialize(int32_t m, int32_t n, const uint8_t* proof, uint8_t* out) {
  if (!proof || !out || m < 1 || n < 1) return MP_ERR_INVALID_ARG;
  wire_proof_serialize(m, n, proof, out);
  return MP_OK;
}
extern "C" int32_t mp_proof_deserialize(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* in, uint8_t* out_proof) {
  return wire_proof_deserialize(ctx, m, n, in, out_proof);
}

extern "C" int32_t mp_msm_num_windows(int32_t window_bits) {
  return (window_bits >= 2 && window_bits <= 16) ? msm_num_windows(window_bits) : 0;
}
extern "C" int32_t mp_msm_g1_windows_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                            int32_t window_bits, int32_t w_begin, int32_t w_count, void* d_out) {
  if (window_bits < 2 || window_bits > 16) return ctx ? ctx->fail(MP_ERR_INVALID_ARG, "explicit window_bits required") : MP_ERR_INVALID_ARG;
  return msm_device_common(ctx, d_bases, d_scalars, n, 1, window_bits, d_out, w_begin, w_count);
}

// ------------------------------------------------------------------------------------------
// debug / parity hooks
// ------------------------------------------------------------------------------------------
__global__ void k_dbg_fq_mul(const fq* a, const fq* b, fq* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fq_mul(a[i], b[i]);
}
__global__ void k_dbg_point_add(const uint32_t* p, const uint32_t* q, uint32_t* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xyzz acc = xyzz_from_affine(affine_from_canonical(p + i * 16));
  xyzz_madd(acc, affine_from_canonical(q + i * 16));
  affine a = xyzz_to_affine(acc);
  if (affine_is_identity(a)) { for (int k = 0; k < 16; k++) out[i * 16 + k] = 0; }
  else affine_to_canonical(a, out + i * 16);
}
__global__ void k_dbg_scalar_mul(const uint32_t* p, const uint32_t* k, uint32_t* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine P = affine_from_canonical(p + i * 16);
  xyzz acc = xyzz_identity();
  for (int bit = 255; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((k[i * 8 + (bit >> 5)] >> (bit & 31)) & 1) xyzz_madd(acc, P);
  }
  affine a = xyzz_to_affine(acc);
  if (affine_is_identity(a)) { for (int w = 0; w < 16; w++) out[i * 16 + w] = 0; }
  else affine_to_canonical(a, out + i * 16);
}

template <typename K>
static int32_t dbg_map(mp_ctx* ctx, const uint8_t* a, size_t abytes, const uint8_t* b, size_t bbytes,
                       uint64_t n, uint8_t* out, size_t obytes, K launch) {
  if (!ctx || !a || !b || !out) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  uint8_t* d = (uint8_t*)ctx->scratch(mp_ctx::kSlotStageIn, (abytes + bbytes + obytes) * n + 256);
  if (!d) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  uint8_t *da = d, *db = d + abytes * n, *dout = db + bbytes * n;
  cudaMemcpyAsync(da, a, abytes * n, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(db, b, bbytes * n, cudaMemcpyHostToDevice, ctx->stream);
  launch(da, db, dout);
  cudaMemcpyAsync(out, dout, obytes * n, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "debug kernel");
  ctx->launches = 1;
  return MP_OK;
}

extern "C" int32_t mp_dbg_fq_mul(mp_ctx* ctx, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out) {
  return dbg_map(ctx, a, 32, b, 32, n, out, 32, [&](uint8_t* da, uint8_t* db, uint8_t* dout) {
    k_dbg_fq_mul<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const fq*)da, (const fq*)db, (fq*)dout, n);
  });
}
extern "C" int32_t mp_dbg_point_add(mp_ctx* ctx, const uint8_t* p, const uint8_t* q, uint64_t n, uint8_t* out) {
  return dbg_map(ctx, p, 64, q, 64, n, out, 64, [&](uint8_t* da, uint8_t* db, uint8_t* dout) {
    k_dbg_point_add<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const uint32_t*)da, (const uint32_t*)db, (uint32_t*)dout, n);
  });
}
extern "C" int32_t mp_dbg_scalar_mul(mp_ctx* ctx, const uint8_t* p, const uint8_t* k, uint64_t n, uint8_t* out) {
  return dbg_map(ctx, p, 64, k, 32, n, out, 64, [&](uint8_t* da, uint8_t* db, uint8_t* dout) {
    k_dbg_scalar_mul<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const uint32_t*)da, (const uint32_t*)db, (uint32_t*)dout, n);
  });
}

// host-side probe: milliseconds to absorb `n_points` 64-byte points into a transcript (the
// statement absorb of a proof is 4 * cards + n + m + 4 points)
extern "C" double mp_dbg_transcript_ms(uint64_t n_points) {
  std::vector<uint8_t> pts(n_points * 64, 7);
  auto t0 = std::chrono::steady_clock::now();
  Transcript fs;
  fs.begin();
  fs.feed_points64(pts.data(), n_points);
  fs.end();
  volatile uint32_t sink = fs.challenge().v[0];
  (void)sink;
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------------------------------
// integer-pipe microbenchmarks (roofline denominators measured on the box)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bench_imad_wide(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
  unsigned long long acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = k + a;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b + r));
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}
__global__ void __launch_bounds__(256) k_bench_imad_lo(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
  uint32_t acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = k + a;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a + k), "r"(b + r));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(128) k_bench_fq_mul(uint32_t* out, int iters, uint32_t seed) {
  fq x = fq_one(), y = fq_r2();
  x.v[0] ^= seed + threadIdx.x;
  y.v[0] ^= blockIdx.x;
  x = fq_reduce_weak(x); y = fq_reduce_weak(y);
  for (int it = 0; it < iters; it++) { x = fq_mul(x, y); y = fq_mul(y, x); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ y.v[7];
}
__global__ void __launch_bounds__(128) k_bench_fq_sqr(uint32_t* out, int iters, uint32_t seed) {
  fq x = fq_one(), y = fq_r2();
  x.v[0] ^= seed + threadIdx.x;
  y.v[0] ^= blockIdx.x;
  x = fq_reduce_weak(x); y = fq_reduce_weak(y);
  for (int it = 0; it < iters; it++) { x = fq_sqr(x); y = fq_sqr(y); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ y.v[7];
}
// carry-chained wide multiply-adds (the form fq_mul uses): 8 chains of 4 lo/hi pairs
__global__ void __launch_bounds__(256) k_bench_imad_wide_cc(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
  uint32_t acc[9];
#pragma unroll
  for (int k = 0; k < 9; k++) acc[k] = k + a;
  for (int it = 0; it < iters; it++) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int r = 0; r < 16; r++) {
      MP_ROW_CHAIN(acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7], acc[8], a, a + 1, a + 2, a + 3, b + r);
    }
#endif
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 9; k++) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 1:1 mix of independent wide multiply-adds (FMA pipe) and 3-input adds (ALU pipe)
__global__ void __launch_bounds__(256) k_bench_mix(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
  unsigned long long acc[8];
  uint32_t s[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { acc[k] = k + a; s[k] = k * b; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b + r));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(s[k]) : "r"(a + r));
      }
    }
  }
  unsigned long long x = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) x ^= acc[k] + s[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)x ^ (uint32_t)(x >> 32);
}
// FP64 pipe probes (is the DFMA pipe free next to the integer multiplier?  52-bit-limb products through
// fma.rz.f64 are the known alternative to 32x32 IMAD.WIDE for wide modular arithmetic):
// MODE 0: DFMA only; 1: DFMA + IMAD.WIDE 1:1; 2: DFMA + 32-bit add 1:1.  Counted op = one DFMA.
template <int MODE>
__global__ void __launch_bounds__(256) k_bench_dfma(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
  double f[8];
  unsigned long long acc[8];
  uint32_t s[8];
  const double m = 1.0 + 1e-9 * (double)(threadIdx.x & 7), c = 1e-3 * (double)(blockIdx.x & 3);
#pragma unroll
  for (int k = 0; k < 8; k++) { f[k] = 1.0 + k; acc[k] = k + a; s[k] = k * b; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(f[k]) : "d"(m), "d"(c));
        if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b + r));
        if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[k]) : "r"(a + r));
      }
    }
  }
  unsigned long long x = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) x ^= acc[k] + s[k] + (unsigned long long)__double_as_longlong(f[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)x ^ (uint32_t)(x >> 32);
}
__global__ void __launch_bounds__(128) k_bench_madd(uint32_t* out, int iters, uint32_t seed) {
  // G in canonical form -> Montgomery; acc walks G, 2G, 3G, ... (no special cases hit)
  const uint32_t g[16] = {0xc943cfcau, 0x3d723d8bu, 0x0d1819e0u, 0xdeacfd9bu, 0x5a40f0c7u, 0x7beced41u,
                          0x8599971bu, 0x01ef15c1u, 0x36e8dc1fu, 0x2873000cu, 0x1abe43a3u, 0xde53ecd1u,
                          0xdf46ec62u, 0xb7be4801u, 0x0aa49730u, 0x00566806u};
  affine P = affine_from_canonical(g);
  xyzz acc = xyzz_dbl_affine(P);
  if ((seed + threadIdx.x) == 0xffffffffu) acc = xyzz_dbl(acc);
  for (int it = 0; it < iters; it++) xyzz_madd(acc, P);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.X.v[0] ^ acc.ZZZ.v[3];
}

extern "C" int32_t mp_dbg_bench(mp_ctx* ctx, int32_t which, int32_t iters, float* ms, double* ops) {
  if (!ctx || !ms || !ops || iters <= 0) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  const int blocks = 148 * 8;
  uint32_t* d = (uint32_t*)ctx->scratch(mp_ctx::kSlotStageOut, sizeof(uint32_t) * blocks * 256);
  if (!d) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {  // first pass = warm-up
    cudaEventRecord(e0, ctx->stream);
    switch (which) {
      case 0: k_bench_imad_wide<<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      case 1: k_bench_imad_lo<<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      case 2: k_bench_fq_mul<<<blocks, 128, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 128 * iters * 2; break;
      case 3: k_bench_madd<<<blocks, 128, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 128 * iters; break;
      case 4: k_bench_fq_sqr<<<blocks, 128, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 128 * iters * 2; break;
      case 5: k_bench_imad_wide_cc<<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      case 6: k_bench_mix<<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      case 7: k_bench_dfma<0><<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      case 8: k_bench_dfma<1><<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      case 9: k_bench_dfma<2><<<blocks, 256, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 256 * iters * 64; break;
      default: cudaEventDestroy(e0); cudaEventDestroy(e1); return MP_ERR_INVALID_ARG;
    }
    cudaEventRecord(e1, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); return ctx->cuda_fail(e, "bench kernel"); }
  }
  cudaEventElapsedTime(ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->launches = 2;
  return MP_OK;
}
