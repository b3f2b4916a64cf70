// extern "C" surface of the BLS12-377 G1 instantiation (declared in include/mpshuffle_bls12_377.h).
//
// the reference's protocol is generic over `C: ProjectiveCurve`
// (src/discrete_log_cards/mod.rs:86) and its only benchmark harness instantiates it over
// `ark_bls12_377::G1Projective` (examples/parameter_selection.rs:25-26).  This translation unit and a
// second copy of msm.cu are compiled with -DMP_CURVE_BLS12_377 -Dmp=mp_bls12_377: the same Pippenger
// pipeline (msm.cu), the same XYZZ group law (ec.cuh, a = 0 branch) over the 12-limb field of
// fq_bls12_377.cuh, linked into libmpshuffle.so next to the Stark-curve build.  Built so far: the group
// layer under the protocol -- variable-base MSM, ciphertext (2-component) MSM, batched MSM jobs, fixed-base
// batched Pedersen commitments -- i.e. the kernels all of ShuffleArgument::{prove,verify} reduce to, and
// verify_shuffle on top of them (host-scalar form); the prover and the device-resident drivers are Stark-only.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <initializer_list>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mpshuffle_bls12_377.h"
#include "ctx.cuh"
#include "msm.cuh"
#include "shuffle.cuh"
#include "shuffle_host.hpp"
#include "wire_host.hpp"

#ifndef MP_CURVE_BLS12_377
#error "capi_bls12_377.cu must be compiled with -DMP_CURVE_BLS12_377 -Dmp=mp_bls12_377"
#endif

using namespace mp;

static constexpr size_t kFe = 4 * kFqLimbs;  // 48 bytes per coordinate
static constexpr size_t kPt = 2 * kFe;       // 96 bytes per affine point

struct mp377_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  MsmWorkspace* ws = nullptr;
  // protocol-driver context of this curve (parameters, fixed-base tables, staging): the engine context type the
  // Stark build calls mp_ctx, compiled a second time under its own name (Makefile: -Dmp_ctx=mp377_pctx)
  mp_ctx* prover = nullptr;
  std::string err;
  int launches = 0;
  uint64_t last_ec_adds = 0;
  int last_window = 0;
  // commit key (h, G_1 .. G_len) as a window-major fixed-base table
  uint32_t ck_nb = 0;
  int ck_c = 0;
  struct Buf { void* ptr = nullptr; size_t cap = 0; };
  enum { kStageIn = 0, kStageOut, kPointsMont, kMsmOut, kFlags, kCkBases, kCkTable, kScal, kSlots };
  Buf bufs[kSlots];

  void* scratch(int slot, size_t bytes) {
    Buf& b = bufs[slot];
    if (b.cap < bytes) {
      if (b.ptr) cudaFree(b.ptr);
      b.ptr = nullptr;
      b.cap = 0;
      size_t want = bytes + bytes / 8 + 256;
      if (cudaMalloc(&b.ptr, want) != cudaSuccess) return nullptr;
      b.cap = want;
    }
    return b.ptr;
  }
  int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  int32_t cuda_fail(cudaError_t e, const char* where) {
    return fail(MP_ERR_CUDA, "CUDA error in %s: %s", where, cudaGetErrorString(e));
  }
};

#define CK377(call, where)                                   \
  do {                                                       \
    cudaError_t _e = (call);                                 \
    if (_e != cudaSuccess) return ctx->cuda_fail(_e, where); \
  } while (0)

extern "C" int32_t mp377_ctx_create(mp377_ctx** out, int32_t device) {
  if (!out) return MP_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
    fprintf(stderr, "mpshuffle: no usable CUDA device %d (%s); there is no CPU fallback\n", device,
            e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    return MP_ERR_CUDA;
  }
  if (cudaSetDevice(device) != cudaSuccess) return MP_ERR_CUDA;
  mp377_ctx* ctx = new mp377_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MP_ERR_CUDA; }
  ctx->ws = msm_workspace_create();
  *out = ctx;
  return MP_OK;
}
extern "C" void mp377_ctx_destroy(mp377_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  msm_workspace_destroy(ctx->ws);
  if (ctx->prover) ctx_destroy(ctx->prover);
  for (auto& b : ctx->bufs)
    if (b.ptr) cudaFree(b.ptr);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
extern "C" void* mp377_ctx_stream(mp377_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int32_t mp377_ctx_sync(mp377_ctx* ctx) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  CK377(cudaStreamSynchronize(ctx->stream), "mp377_ctx_sync");
  return MP_OK;
}
extern "C" const char* mp377_last_error_string(mp377_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" int32_t mp377_last_kernel_launches(mp377_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t mp377_last_msm_ec_adds(mp377_ctx* ctx) { return ctx ? ctx->last_ec_adds : 0; }
extern "C" int32_t mp377_last_msm_window(mp377_ctx* ctx) { return ctx ? ctx->last_window : 0; }
extern "C" int32_t mp377_msm_num_windows(int32_t window_bits) {
  return window_bits >= 2 && window_bits <= 16 ? msm_num_windows(window_bits) : 0;
}

// ------------------------------------------------------------------------------------------
// MSM entry points
// ------------------------------------------------------------------------------------------
static uint64_t scheduled_ec_adds(uint64_t n, int c, int ncomp) {
  const uint64_t W = msm_num_windows(c), B = 1ull << (c - 1);
  return (uint64_t)ncomp * (W * (n + 2 * B) + W * (uint64_t)c);
}

static int32_t msm_device_common(mp377_ctx* ctx, const void* d_points, const void* d_scalars, uint64_t n, int ncomp,
                                 int32_t window_bits, void* d_out, int w_begin = 0, int w_count = -1) {
  if (!ctx || (!d_points && n) || (!d_scalars && n) || !d_out) return MP_ERR_INVALID_ARG;
  if (n >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "MSM size %llu too large", (unsigned long long)n);
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int c = window_bits > 0 ? window_bits : msm_pick_window(n);
  if (c < 2 || c > 16) return ctx->fail(MP_ERR_INVALID_ARG, "window_bits %d out of range [2,16]", c);
  ctx->last_window = c;
  if (w_begin < 0 || (w_count >= 0 && w_begin + w_count > msm_num_windows(c)) || w_count == 0)
    return ctx->fail(MP_ERR_INVALID_ARG, "window range [%d, +%d) outside [0, %d)", w_begin, w_count, msm_num_windows(c));
  ctx->last_ec_adds = scheduled_ec_adds(n, c, ncomp) * (uint64_t)(w_count < 0 ? msm_num_windows(c) - w_begin : w_count) / msm_num_windows(c);
  affine* mont = (affine*)ctx->scratch(mp377_ctx::kPointsMont, sizeof(affine) * n * ncomp);
  xyzz* res = (xyzz*)ctx->scratch(mp377_ctx::kMsmOut, sizeof(xyzz) * ncomp);
  int* bad = (int*)ctx->scratch(mp377_ctx::kFlags, 256);
  if (!mont || !res || !bad) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  CK377(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream), "memset");
  CK377(points_to_mont((const uint32_t*)d_points, mont, n * ncomp, bad, ctx->stream), "points_to_mont");
  ctx->launches += n ? 1 : 0;
  MsmJob job{0, 0, (uint32_t)n};
  CK377(msm_run(ctx->ws, (const uint32_t*)d_scalars, n, mont, ncomp, &job, 1, c, res, ctx->stream, w_begin, w_count), "msm_run");
  ctx->launches += msm_last_launches(ctx->ws);
  CK377(xyzz_to_canonical(res, (uint32_t*)d_out, ncomp, ctx->stream), "xyzz_to_canonical");
  ctx->launches += 1;
  return MP_OK;
}

static int32_t msm_host_common(mp377_ctx* ctx, const uint8_t* points, const uint8_t* scalars, uint64_t n, int ncomp,
                               int32_t window_bits, uint8_t* out) {
  if (!ctx || (!points && n) || (!scalars && n) || !out) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  const size_t pbytes = (size_t)n * kPt * ncomp, sbytes = (size_t)n * 32;
  uint8_t* d_in = (uint8_t*)ctx->scratch(mp377_ctx::kStageIn, pbytes + sbytes + 256);
  uint8_t* d_out = (uint8_t*)ctx->scratch(mp377_ctx::kStageOut, kPt * ncomp);
  if (!d_in || !d_out) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  if (n) {
    CK377(cudaMemcpyAsync(d_in, points, pbytes, cudaMemcpyHostToDevice, ctx->stream), "H2D points");
    CK377(cudaMemcpyAsync(d_in + pbytes, scalars, sbytes, cudaMemcpyHostToDevice, ctx->stream), "H2D scalars");
  }
  int32_t st = msm_device_common(ctx, d_in, d_in + pbytes, n, ncomp, window_bits, d_out);
  if (st != MP_OK) return st;
  int bad = 0;
  int* d_bad = (int*)ctx->scratch(mp377_ctx::kFlags, 256);
  CK377(cudaMemcpyAsync(out, d_out, kPt * ncomp, cudaMemcpyDeviceToHost, ctx->stream), "D2H result");
  CK377(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H flag");
  CK377(cudaStreamSynchronize(ctx->stream), "MSM execution");
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "an input point is not a canonical point of BLS12-377 G1");
  return MP_OK;
}

extern "C" int32_t mp377_msm_g1(mp377_ctx* ctx, const uint8_t* bases, const uint8_t* scalars, uint64_t n,
                                int32_t window_bits, uint8_t* out) {
  return msm_host_common(ctx, bases, scalars, n, 1, window_bits, out);
}
extern "C" int32_t mp377_ct_msm(mp377_ctx* ctx, const uint8_t* deck, const uint8_t* scalars, uint64_t n,
                                int32_t window_bits, uint8_t* out) {
  return msm_host_common(ctx, deck, scalars, n, 2, window_bits, out);
}
extern "C" int32_t mp377_msm_g1_device(mp377_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                       int32_t window_bits, void* d_out) {
  return msm_device_common(ctx, d_bases, d_scalars, n, 1, window_bits, d_out);
}
extern "C" int32_t mp377_ct_msm_device(mp377_ctx* ctx, const void* d_deck, const void* d_scalars, uint64_t n,
                                       int32_t window_bits, void* d_out) {
  return msm_device_common(ctx, d_deck, d_scalars, n, 2, window_bits, d_out);
}

// Batch of MSMs over ONE point array and ONE scalar array (kernel family K2 of SURVEY.md 2b, the
// `mp_msm_batch_shared_bases` of section 8(b)): job j = sum_t scalars[scalar_off_j + t] * points[point_off_j + t],
// t < len_j.  The multi-exponentiation argument's diagonal products are m(m+1) such jobs over the rows of the
// shuffled deck (reference call site mod.rs:409-415 -> MultiExponentiationArgument); one launch sequence
// evaluates all of them.  jobs: njobs x (scalar_off, point_off, len) as uint32.  out: njobs * ncomp points.
extern "C" int32_t mp377_msm_jobs(mp377_ctx* ctx, const uint8_t* points, uint64_t n_points, int32_t ncomp,
                                  const uint8_t* scalars, uint64_t n_scalars, const uint32_t* jobs, uint64_t njobs,
                                  int32_t window_bits, uint8_t* out) {
  if (!ctx || !jobs || !out || (!points && n_points) || (!scalars && n_scalars)) return MP_ERR_INVALID_ARG;
  if (ncomp != 1 && ncomp != 2) return ctx->fail(MP_ERR_INVALID_ARG, "ncomp must be 1 (G1) or 2 (ciphertexts)");
  if (njobs == 0) return MP_OK;
  if (njobs >= (1u << 24) || n_points >= (1ull << 31) || n_scalars >= (1ull << 31))
    return ctx->fail(MP_ERR_INVALID_ARG, "batch too large");
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
  const size_t pbytes = (size_t)n_points * kPt * ncomp, sbytes = (size_t)n_scalars * 32;
  uint8_t* d_in = (uint8_t*)ctx->scratch(mp377_ctx::kStageIn, pbytes + sbytes + 256);
  affine* mont = (affine*)ctx->scratch(mp377_ctx::kPointsMont, sizeof(affine) * n_points * ncomp);
  xyzz* res = (xyzz*)ctx->scratch(mp377_ctx::kMsmOut, sizeof(xyzz) * njobs * ncomp);
  uint8_t* d_out = (uint8_t*)ctx->scratch(mp377_ctx::kStageOut, kPt * njobs * ncomp);
  int* bad = (int*)ctx->scratch(mp377_ctx::kFlags, 256);
  if (!d_in || !mont || !res || !d_out || !bad) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  if (pbytes) CK377(cudaMemcpyAsync(d_in, points, pbytes, cudaMemcpyHostToDevice, ctx->stream), "H2D points");
  if (sbytes) CK377(cudaMemcpyAsync(d_in + pbytes, scalars, sbytes, cudaMemcpyHostToDevice, ctx->stream), "H2D scalars");
  CK377(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream), "memset");
  CK377(points_to_mont((const uint32_t*)d_in, mont, n_points * ncomp, bad, ctx->stream), "points_to_mont");
  ctx->launches += n_points ? 1 : 0;
  CK377(msm_run(ctx->ws, (const uint32_t*)(d_in + pbytes), n_scalars, mont, ncomp, h_jobs.data(), (int)njobs, c, res, ctx->stream),
        "msm_run (jobs)");
  ctx->launches += msm_last_launches(ctx->ws);
  CK377(xyzz_to_canonical(res, (uint32_t*)d_out, njobs * ncomp, ctx->stream), "xyzz_to_canonical");
  ctx->launches += 1;
  int h_bad = 0;
  CK377(cudaMemcpyAsync(out, d_out, kPt * njobs * ncomp, cudaMemcpyDeviceToHost, ctx->stream), "D2H results");
  CK377(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H flag");
  CK377(cudaStreamSynchronize(ctx->stream), "MSM jobs");
  if (h_bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "an input point is not a canonical point of BLS12-377 G1");
  ctx->last_ec_adds = 0;
  for (auto& j : h_jobs) ctx->last_ec_adds += scheduled_ec_adds(j.len, c, ncomp);
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// verify_shuffle over this curve (BarnettSmartProtocol::verify_shuffle, reference src/lib.rs:191-197, impl
// mod.rs:420-443, instantiated as in examples/parameter_selection.rs:25-29).  The host half is the curve-generic
// csrc/shuffle_host.hpp -- transcript, challenges, every verifier check rewritten as "sum scalar * point == O"
// (CPU-tested for both curves in tests/test_host_verify_plan.py) -- with the O(N) scalars computed on the host;
// the group work is two launch sequences of the batched MSM above: the eight commitment-space equations as
// eight G1 jobs, the two ciphertext equations as two 2-component jobs.  (The Stark-curve verifier additionally
// moves the O(N) scalar work to device kernels and keeps decks resident; that driver is not built for this
// curve yet.)
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// shuffle_and_remask over this curve (BarnettSmartProtocol::shuffle_and_remask, reference src/lib.rs:181-188, impl
// mod.rs:380-418, instantiated as in examples/parameter_selection.rs:25-29 -- the call the reference's only
// benchmark harness times).  The protocol driver is the SAME source as the Stark build's lockstep prover
// (shuffle_setup.cu: parameters, fixed-base tables, remasking; shuffle_prove_batch.cu: the rounds of the argument
// with host-side scalar algebra and every group operation in batched MSM launches), compiled a second time against
// the 12-limb field; proofs are byte-identical to oracle/py/bayer_groth.py under curve("bls12_377").
// ------------------------------------------------------------------------------------------
static int32_t prover_status(mp377_ctx* ctx, int32_t st) {
  if (st < 0 && ctx->prover) ctx->err = ctx->prover->err;
  if (ctx->prover) ctx->launches = ctx->prover->launches;
  return st;
}
extern "C" int32_t mp377_ctx_set_params(mp377_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g, const uint8_t* ck_g,
                                        const uint8_t* ck_h, const uint8_t* ghat) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  if (!ctx->prover && ctx_create(&ctx->prover, ctx->device) != MP_OK) return ctx->fail(MP_ERR_CUDA, "cannot create the prover context");
  return prover_status(ctx, shuffle_set_params(ctx->prover, m, n, enc_g, ck_g, ck_h, ghat));
}
extern "C" uint64_t mp377_prover_randomness_len(int32_t m, int32_t n) { return shuffle_randomness_len(m, n); }
extern "C" int32_t mp377_remask_batch(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm, const uint8_t* rho,
                                      uint64_t n_cards, uint8_t* out_deck) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  if (!ctx->prover) return ctx->fail(MP_ERR_NO_PARAMS, "mp377_ctx_set_params has not been called");
  return prover_status(ctx, shuffle_remask(ctx->prover, pk, deck, perm, rho, n_cards, out_deck));
}
extern "C" int32_t mp377_shuffle_and_remask_batch(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint32_t* perms,
                                                  const uint8_t* rhos, const uint8_t* randomness, uint64_t batch, uint8_t* out_decks,
                                                  uint8_t* proofs, int32_t host_threads) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  if (!ctx->prover) return ctx->fail(MP_ERR_NO_PARAMS, "mp377_ctx_set_params has not been called");
  return prover_status(ctx, shuffle_prove_batch(ctx->prover, pk, decks, perms, rhos, randomness, batch, out_decks, proofs, host_threads));
}
extern "C" int32_t mp377_shuffle_and_remask(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                            const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck, uint8_t* proof_out) {
  return mp377_shuffle_and_remask_batch(ctx, pk, deck, perm, rho, randomness, 1, out_deck, proof_out, 1);
}

// ------------------------------------------------------------------------------------------
// Batched sigma protocols either side of the shuffle over this curve (SURVEY.md section 8(f) ranks 1 + 3):
// BarnettSmartProtocol::{mask, verify_mask, remask, verify_remask, compute_reveal_token, verify_reveal,
// prove_key_ownership, verify_key_ownership} (reference src/lib.rs:88-175, impl mod.rs:132-354) for n items per call.
// Same source as the Stark build (sigma.cu: one k_lincomb launch per pass, sigma_host.hpp: the per-proof Fiat-Shamir
// transcripts on host threads), compiled against the 12-limb field; the verifiers first put every untrusted point
// through the G1 membership test, as the reference's deserialiser would.
// ------------------------------------------------------------------------------------------
static constexpr size_t kSigPt = 96, kSigCp = 2 * kSigPt + 32, kSigSchnorr = kSigPt + 32;
#define NEED_PROVER(ctx)                                                                                     \
  do {                                                                                                       \
    if (!(ctx)) return MP_ERR_INVALID_ARG;                                                                   \
    if (!(ctx)->prover) return (ctx)->fail(MP_ERR_NO_PARAMS, "mp377_ctx_set_params has not been called");    \
  } while (0)
// G1 membership of the points a sigma verifier is handed.  `arrays`: contiguous points, `per` of them per item
// (per == 0: call-level points such as keys -- a bad one fails the call); `proofs`: the leading `lead` bytes of each of
// n records of `rec` bytes.  item_bad[i] is set for items with a point that is not a canonical point of G1: those fail
// alone with MP_VERIFY_MALFORMED, as the reference's deserialiser would fail them, and the rest of the batch is checked.
struct SigmaPts { const uint8_t* p; uint64_t count; uint64_t per; };
static int32_t sigma_points_in_g1(mp377_ctx* ctx, std::initializer_list<SigmaPts> arrays, const uint8_t* proofs, uint64_t n,
                                  size_t rec, size_t lead, std::vector<uint8_t>& item_bad) {
  item_bad.assign(n, 0);
  std::vector<uint8_t> all;
  std::vector<int64_t> owner;   // item index, or -1 for a call-level point
  for (const auto& a : arrays) {
    if (!a.p || !a.count) continue;
    all.insert(all.end(), a.p, a.p + a.count * kSigPt);
    for (uint64_t k = 0; k < a.count; k++) owner.push_back(a.per ? (int64_t)(k / a.per) : -1);
  }
  if (proofs)
    for (uint64_t i = 0; i < n; i++) {
      all.insert(all.end(), proofs + rec * i, proofs + rec * i + lead);
      for (size_t k = 0; k < lead / kSigPt; k++) owner.push_back((int64_t)i);
    }
  if (all.empty()) return MP_OK;
  std::vector<int32_t> st(owner.size());
  const int32_t rc = mp377_subgroup_check(ctx, all.data(), owner.size(), st.data());
  if (rc == MP_OK) return MP_OK;
  if (rc != MP_ERR_NOT_ON_CURVE && rc != MP_ERR_NOT_IN_SUBGROUP) return rc;
  for (size_t k = 0; k < owner.size(); k++) {
    if (!st[k]) continue;
    if (owner[k] < 0) return st[k] == 1 ? MP_ERR_NOT_ON_CURVE : MP_ERR_NOT_IN_SUBGROUP;   // ctx->err was set by the check
    item_bad[(size_t)owner[k]] = 1;
  }
  return MP_OK;
}
// an off-curve point of a flagged item would fail the whole inner call through its call-level points only; the
// inner verifier flags off-curve item points itself, and the items flagged here are overridden afterwards
static int32_t sigma_finish(mp377_ctx* ctx, int32_t rc, const std::vector<uint8_t>& item_bad, int32_t* statuses) {
  rc = prover_status(ctx, rc);
  if (rc == MP_OK)
    for (size_t i = 0; i < item_bad.size(); i++)
      if (item_bad[i]) statuses[i] = MP_VERIFY_MALFORMED;
  return rc;
}
extern "C" int32_t mp377_mask_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* r,
                                    const uint8_t* omega, uint64_t n, uint8_t* out_masked, uint8_t* out_proofs, int32_t host_threads) {
  NEED_PROVER(ctx);
  return prover_status(ctx, sigma_mask_batch(ctx->prover, shared_key, cards, r, omega, n, out_masked, out_proofs, host_threads));
}
extern "C" int32_t mp377_verify_mask_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* masked,
                                           const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  NEED_PROVER(ctx);
  if (!shared_key || (n && (!cards || !masked || !proofs || !statuses))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  std::vector<uint8_t> item_bad;
  int32_t sg = sigma_points_in_g1(ctx, {{shared_key, 1, 0}, {cards, n, 1}, {masked, 2 * n, 2}}, proofs, n, kSigCp, 2 * kSigPt, item_bad);
  if (sg != MP_OK) return sg;
  return sigma_finish(ctx, sigma_verify_mask_batch(ctx->prover, shared_key, cards, masked, proofs, n, statuses, host_threads), item_bad, statuses);
}
extern "C" int32_t mp377_remask_prove_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* alpha,
                                            const uint8_t* omega, uint64_t n, uint8_t* out_deck, uint8_t* out_proofs,
                                            int32_t host_threads) {
  NEED_PROVER(ctx);
  return prover_status(ctx, sigma_remask_prove_batch(ctx->prover, shared_key, deck, alpha, omega, n, out_deck, out_proofs, host_threads));
}
extern "C" int32_t mp377_verify_remask_batch(mp377_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* remasked,
                                             const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  NEED_PROVER(ctx);
  if (!shared_key || (n && (!deck || !remasked || !proofs || !statuses))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  std::vector<uint8_t> item_bad;
  int32_t sg = sigma_points_in_g1(ctx, {{shared_key, 1, 0}, {deck, 2 * n, 2}, {remasked, 2 * n, 2}}, proofs, n, kSigCp, 2 * kSigPt, item_bad);
  if (sg != MP_OK) return sg;
  return sigma_finish(ctx, sigma_verify_remask_batch(ctx->prover, shared_key, deck, remasked, proofs, n, statuses, host_threads), item_bad, statuses);
}
extern "C" int32_t mp377_reveal_batch(mp377_ctx* ctx, const uint8_t* sk, const uint8_t* pk, const uint8_t* masked,
                                      const uint8_t* omega, uint64_t n, uint8_t* out_tokens, uint8_t* out_proofs, int32_t host_threads) {
  NEED_PROVER(ctx);
  return prover_status(ctx, sigma_reveal_batch(ctx->prover, sk, pk, masked, omega, n, out_tokens, out_proofs, host_threads));
}
extern "C" int32_t mp377_verify_reveal_batch(mp377_ctx* ctx, const uint8_t* pk, const uint8_t* tokens, const uint8_t* masked,
                                             const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  NEED_PROVER(ctx);
  if (!pk || (n && (!tokens || !masked || !proofs || !statuses))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  std::vector<uint8_t> item_bad;
  int32_t sg = sigma_points_in_g1(ctx, {{pk, 1, 0}, {tokens, n, 1}, {masked, 2 * n, 2}}, proofs, n, kSigCp, 2 * kSigPt, item_bad);
  if (sg != MP_OK) return sg;
  return sigma_finish(ctx, sigma_verify_reveal_batch(ctx->prover, pk, tokens, masked, proofs, n, statuses, host_threads), item_bad, statuses);
}
extern "C" int32_t mp377_key_ownership_prove_batch(mp377_ctx* ctx, const uint8_t* pks, const uint8_t* sks, const uint8_t* infos,
                                                   const uint64_t* info_offsets, const uint8_t* omega, uint64_t n,
                                                   uint8_t* out_proofs, int32_t host_threads) {
  NEED_PROVER(ctx);
  return prover_status(ctx, sigma_key_ownership_prove_batch(ctx->prover, pks, sks, infos, info_offsets, omega, n, out_proofs, host_threads));
}
extern "C" int32_t mp377_key_ownership_verify_batch(mp377_ctx* ctx, const uint8_t* pks, const uint8_t* infos,
                                                    const uint64_t* info_offsets, const uint8_t* proofs, uint64_t n,
                                                    int32_t* statuses, int32_t host_threads) {
  NEED_PROVER(ctx);
  if (n && (!pks || !info_offsets || !proofs || !statuses)) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  std::vector<uint8_t> item_bad;
  int32_t sg = sigma_points_in_g1(ctx, {{pks, n, 1}}, proofs, n, kSigSchnorr, kSigPt, item_bad);
  if (sg != MP_OK) return sg;
  return sigma_finish(ctx, sigma_key_ownership_verify_batch(ctx->prover, pks, infos, info_offsets, proofs, n, statuses, host_threads), item_bad, statuses);
}

// ------------------------------------------------------------------------------------------
// Subgroup membership.  E(F_q) has cofactor h = 0x170b5d44300000000000000000000000; the protocol lives in the
// order-r subgroup G1, and every verifier scalar is reduced mod r, so points with a cofactor-torsion component must
// not reach the verifier (small-subgroup malleability).  In the reference every point arrives through ark-ec 0.3
// `CanonicalDeserialize`, which rejects them (is_in_correct_subgroup_assuming_on_curve).  Test used here -- the one
// ark-bls12-377 uses too (M. Scott, "A note on group membership tests for G1, G2 and GT on BLS pairing-friendly
// curves"): with the endomorphism phi(x, y) = (beta x, y), beta a primitive cube root of unity in F_q,
//     P in G1  <=>  phi(P) == -[u^2] P,      u = 0x8508c00000000001 (the BLS parameter),
// i.e. one 127-bit double-and-add (Hamming weight 22) per point instead of a 253-bit one.  Checked against
// big-int arithmetic in tests/test_oracle_bls12_377.py for subgroup points, random curve points, pure torsion
// points and G1 + torsion.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k377_subgroup_check(const affine* __restrict__ pts, uint64_t n, int* __restrict__ bad,
                                                           uint64_t per_item) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  affine P;
  {
    const uint4* s = reinterpret_cast<const uint4*>(pts + i);
    uint4* d = reinterpret_cast<uint4*>(&P);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(affine) / 16); k++) d[k] = __ldg(s + k);
  }
  if (affine_is_identity(P)) return;  // the identity is in every subgroup
  const uint32_t u2[4] = {0x00000001u, 0x0a118000u, 0x90000001u, 0x452217ccu};  // u^2, 127 bits
  xyzz acc = xyzz_identity();
#pragma unroll 1
  for (int bit = 126; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((u2[bit >> 5] >> (bit & 31)) & 1) xyzz_madd(acc, P);
  }
  // phi(P) == -acc   <=>   acc.X == beta x ZZ  and  acc.Y == -(y ZZZ)      (acc = O cannot equal -phi(P) != O)
  fq beta;  // beta * R mod q
  {
    const uint32_t b[12] = {0x5a7b8727u, 0x2c766f92u, 0x253d58b5u, 0x03d7f6b0u, 0xec122131u, 0x838ec0deu,
                            0xf658bb10u, 0xbd5eb3e9u, 0x6ed3e52eu, 0x6942bd12u, 0xdd04ed6au, 0x01673786u};
#pragma unroll
    for (int k = 0; k < 12; k++) beta.v[k] = b[k];
  }
  bool ok = !xyzz_is_identity(acc);
  if (ok) {
    const fq lx = fq_reduce_full(acc.X), rx = fq_reduce_full(fq_mul(fq_mul(beta, P.x), acc.ZZ));
    const fq ly = fq_reduce_full(acc.Y), ry = fq_reduce_full(fq_neg2(fq_mul(P.y, acc.ZZZ)));
    ok = fq_eq_raw(lx, rx) && fq_eq_raw(ly, ry);
  }
  if (!ok) atomicExch(bad + (per_item ? i / per_item : 0), 1);
}

// points: canonical, n * 96 bytes (host).  Sets *in_subgroup = 1 iff every point is a canonical point of the curve AND
// lies in G1.  statuses (optional, n entries): 0 = in G1, 1 = not on the curve / not canonical, 2 = on the curve but
// outside G1.
extern "C" int32_t mp377_subgroup_check(mp377_ctx* ctx, const uint8_t* points, uint64_t n, int32_t* statuses) {
  if (!ctx || (!points && n)) return MP_ERR_INVALID_ARG;
  if (n == 0) return MP_OK;
  if (n >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "too many points");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  uint8_t* d_in = (uint8_t*)ctx->scratch(mp377_ctx::kStageIn, n * kPt + 256);
  affine* mont = (affine*)ctx->scratch(mp377_ctx::kPointsMont, sizeof(affine) * n);
  int* flags = (int*)ctx->scratch(mp377_ctx::kScal, 2 * n * sizeof(int) + 256);
  if (!d_in || !mont || !flags) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  CK377(cudaMemcpyAsync(d_in, points, n * kPt, cudaMemcpyHostToDevice, ctx->stream), "H2D points");
  CK377(cudaMemsetAsync(flags, 0, 2 * n * sizeof(int), ctx->stream), "memset");
  CK377(points_to_mont_items((const uint32_t*)d_in, mont, n, flags, 1, ctx->stream), "points_to_mont");
  k377_subgroup_check<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(mont, n, flags + n, 1);
  CK377(cudaGetLastError(), "k377_subgroup_check");
  ctx->launches = 2;
  std::vector<int> h(2 * n);
  CK377(cudaMemcpyAsync(h.data(), flags, 2 * n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H flags");
  CK377(cudaStreamSynchronize(ctx->stream), "subgroup check");
  int32_t worst = MP_OK;
  for (uint64_t i = 0; i < n; i++) {
    const int32_t st = h[i] ? 1 : (h[n + i] ? 2 : 0);
    if (statuses) statuses[i] = st;
    if (st == 1) worst = MP_ERR_NOT_ON_CURVE;
    else if (st == 2 && worst == MP_OK) worst = MP_ERR_NOT_IN_SUBGROUP;
  }
  if (worst == MP_ERR_NOT_ON_CURVE) return ctx->fail(worst, "a point is not a canonical point of BLS12-377 G1's curve");
  if (worst == MP_ERR_NOT_IN_SUBGROUP) return ctx->fail(worst, "a point is on the curve but outside the order-r subgroup G1");
  return MP_OK;
}

extern "C" uint64_t mp377_proof_len(int32_t m, int32_t n) { return shuffle_proof_len(m, n); }

extern "C" int32_t mp377_shuffle_verify(mp377_ctx* ctx, int32_t m, int32_t n, const uint8_t* enc_g, const uint8_t* ck_g,
                                        const uint8_t* ck_h, const uint8_t* ghat, const uint8_t* pk, const uint8_t* deck,
                                        const uint8_t* shuffled_deck, const uint8_t* proof) {
  if (!ctx || !enc_g || !ck_g || !ck_h || !ghat || !pk || !deck || !shuffled_deck || !proof) return MP_ERR_INVALID_ARG;
  if (m < 1 || n < 2 || (uint64_t)m * n >= (1ull << 26)) return ctx->fail(MP_ERR_INVALID_ARG, "unsupported (m, n) = (%d, %d)", m, n);
  const size_t N = (size_t)m * n;
  int launches = 0;
  // gsum = g_1 + .. + g_n (commitments to constant vectors are c * gsum); also validates the key
  ShuffleParamsHost S;
  S.m = m;
  S.n = n;
  {
    std::vector<uint8_t> ones((size_t)n * 32, 0);
    for (int j = 0; j < n; j++) ones[32 * (size_t)j] = 1;
    int32_t st = msm_host_common(ctx, ck_g, ones.data(), (uint64_t)n, 1, 0, S.gsum);
    if (st != MP_OK) return st;
    launches += ctx->launches;
  }
  S.ck64.resize((size_t)(n + 1) * kPt);
  memcpy(S.ck64.data(), ck_h, kPt);
  memcpy(S.ck64.data() + kPt, ck_g, (size_t)n * kPt);
  memcpy(S.enc_g, enc_g, kPt);
  memcpy(S.ghat, ghat, kPt);
  const Layout L(m, n);
  if (!proof_scalars_canonical(proof, L))
    return ctx->fail(MP_ERR_NOT_CANONICAL, "a scalar of the proof is not below the group order");
  {
    // every untrusted point -- public key, both decks, the 11m + 8 proof points -- and the parameters must lie in G1
    // (what `CanonicalDeserialize` enforces on the reference's side of this call)
    std::vector<uint8_t> all;
    all.reserve((4 * N + 11 * (size_t)m + 8 + (size_t)n + 4) * kPt);
    auto put = [&](const uint8_t* p, size_t count) { all.insert(all.end(), p, p + count * kPt); };
    put(pk, 1); put(enc_g, 1); put(ghat, 1); put(ck_h, 1); put(ck_g, (size_t)n);
    put(deck, 2 * N); put(shuffled_deck, 2 * N);
    put(proof, L.za / kPt); put(proof + L.svpts, 3); put(proof + L.mepts, (L.mea - L.mepts) / kPt);
    int32_t sg = mp377_subgroup_check(ctx, all.data(), all.size() / kPt, nullptr);
    if (sg != MP_OK) return sg;
    launches += ctx->launches;
  }
  const Challenges ch = derive_challenges(&S, pk, deck, shuffled_deck, N, proof, L);
  // the eight commitment-space equations
  TermList tl;
  HostChecks hc;
  append_g1_checks(tl, &S, proof, L, ch, &hc);
  static_assert(sizeof(MsmJob) == 12, "MsmJob is three uint32");
  std::vector<uint8_t> g1_out((size_t)kG1Checks * kPt);
  int32_t st = mp377_msm_jobs(ctx, tl.pts.data(), tl.count(), 1, (const uint8_t*)tl.scal.data(), tl.count(),
                              (const uint32_t*)tl.jobs.data(), (uint64_t)kG1Checks, 0, g1_out.data());
  if (st != MP_OK) return st;
  launches += ctx->launches;
  bool g1_id[kG1Checks];
  for (int j = 0; j < kG1Checks; j++) g1_id[j] = all_zero(g1_out.data() + (size_t)j * kPt, kPt);
  // the two ciphertext equations, each one contiguous 2-component job (layout: assemble_ct_jobs)
  std::vector<uint8_t> cts;
  std::vector<uint32_t> scal;
  uint32_t ct_jobs[6];
  fr bstar;
  assemble_ct_jobs(&S, pk, deck, shuffled_deck, proof, L, ch, cts, scal, ct_jobs, &bstar);
  const size_t nct = scal.size() / 8;
  uint8_t ct_out[2 * kCtBytes];
  st = mp377_msm_jobs(ctx, cts.data(), nct, 2, (const uint8_t*)scal.data(), nct, ct_jobs, 2, 0, ct_out);
  if (st != MP_OK) return st;
  launches += ctx->launches;
  ctx->launches = launches;
  return verdict(hc, bstar, g1_id, all_zero(ct_out, sizeof ct_out));
}

// ------------------------------------------------------------------------------------------
// wire format, serialising half (ark-serialize 0.3 compressed encodings; host byte handling as on the Stark
// curve, csrc/wire_host.hpp): 48-byte compressed points.  Deserialising needs a square root in F_q per point
// (q - 1 = 2^46 * odd) and is not built for this curve yet.
// ------------------------------------------------------------------------------------------
extern "C" int32_t mp377_points_compress(const uint8_t* points, uint64_t n, uint8_t* out) {
  if ((!points || !out) && n) return MP_ERR_INVALID_ARG;
  for (uint64_t i = 0; i < n; i++) wire_compress_point(points + kWirePt * i, out + kWireFe * i);
  return MP_OK;
}
extern "C" uint64_t mp377_deck_serialized_len(uint64_t n_cards) { return wire_deck_len(n_cards); }
extern "C" int32_t mp377_deck_serialize(const uint8_t* deck, uint64_t n_cards, uint8_t* out) {
  if (!out || (!deck && n_cards)) return MP_ERR_INVALID_ARG;
  memcpy(out, &n_cards, 8);  // u64 little-endian length prefix of Vec<MaskedCard>
  for (uint64_t i = 0; i < 2 * n_cards; i++) wire_compress_point(deck + kWirePt * i, out + 8 + kWireFe * i);
  return MP_OK;
}
extern "C" uint64_t mp377_proof_serialized_len(int32_t m, int32_t n) { return wire_proof_len(m, n); }
extern "C" int32_t mp377_proof_serialize(int32_t m, int32_t n, const uint8_t* proof, uint8_t* out) {
  if (!proof || !out || m < 1 || n < 1) return MP_ERR_INVALID_ARG;
  wire_proof_serialize(m, n, proof, out);
  return MP_OK;
}

// Deserialising half (wire.cu compiled for this curve: one square root in F_q per point, two-adicity 46, then the G1
// membership test ark-serialize applies to every deserialised point).  statuses: 0 ok, 1 malformed encoding, 2 x is
// not the abscissa of a curve point, 3 on the curve but outside G1.
static int32_t wire_status(mp377_ctx* ctx, int32_t st) {
  if (!ctx->prover) return st;
  if (st < 0) ctx->err = ctx->prover->err;
  ctx->launches = ctx->prover->launches;
  return st;
}
static int32_t need_wire_ctx(mp377_ctx* ctx) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  if (!ctx->prover && ctx_create(&ctx->prover, ctx->device) != MP_OK) return ctx->fail(MP_ERR_CUDA, "cannot create the protocol context");
  return MP_OK;
}
extern "C" int32_t mp377_points_decompress(mp377_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, int32_t* statuses) {
  int32_t rc = need_wire_ctx(ctx);
  if (rc != MP_OK) return rc;
  rc = wire_status(ctx, wire_points_decompress(ctx->prover, in, n, out, statuses));
  if (rc != MP_OK || n == 0) return rc;
  const int launches = ctx->launches;
  std::vector<int32_t> sg(n);
  rc = mp377_subgroup_check(ctx, out, n, sg.data());
  ctx->launches += launches;
  if (rc == MP_OK) return MP_OK;
  for (uint64_t i = 0; i < n; i++)
    if (sg[i]) {
      if (statuses) statuses[i] = 3;
      memset(out + kPt * i, 0, kPt);
    }
  return rc;
}
extern "C" int32_t mp377_deck_deserialize(mp377_ctx* ctx, const uint8_t* in, uint64_t in_len, uint8_t* out_deck, uint64_t* n_cards) {
  int32_t rc = need_wire_ctx(ctx);
  if (rc != MP_OK) return rc;
  rc = wire_status(ctx, wire_deck_deserialize(ctx->prover, in, in_len, out_deck, n_cards));
  if (rc != MP_OK || *n_cards == 0) return rc;
  const int launches = ctx->launches;
  rc = mp377_subgroup_check(ctx, out_deck, 2 * *n_cards, nullptr);
  ctx->launches += launches;
  return rc;
}
extern "C" int32_t mp377_proof_deserialize(mp377_ctx* ctx, int32_t m, int32_t n, const uint8_t* in, uint8_t* out_proof) {
  int32_t rc = need_wire_ctx(ctx);
  if (rc != MP_OK) return rc;
  rc = wire_status(ctx, wire_proof_deserialize(ctx->prover, m, n, in, out_proof));
  if (rc != MP_OK) return rc;
  const int launches = ctx->launches;
  std::vector<uint8_t> pts;
  const uint8_t* p = out_proof;
  for (const WireRun& r : wire_proof_runs(m, n)) {
    const size_t len = (r.points ? kPt : 32) * r.count;
    if (r.points) pts.insert(pts.end(), p, p + len);
    p += len;
  }
  rc = mp377_subgroup_check(ctx, pts.data(), pts.size() / kPt, nullptr);
  ctx->launches += launches;
  return rc;
}

// Window-range split of one MSM across GPUs: the partial
//   sum_{w in [w_begin, w_begin + w_count)} 2^(c (w - w_begin)) * (window sum w),
// so that  MSM = sum over ranks of 2^(c * w_begin_r) * partial_r  (see mental-poker_b200/dist.py).
extern "C" int32_t mp377_msm_g1_windows_device(mp377_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                               int32_t window_bits, int32_t w_begin, int32_t w_count, void* d_out) {
  if (window_bits < 2 || window_bits > 16) return ctx ? ctx->fail(MP_ERR_INVALID_ARG, "explicit window_bits required") : MP_ERR_INVALID_ARG;
  return msm_device_common(ctx, d_bases, d_scalars, n, 1, window_bits, d_out, w_begin, w_count);
}

extern "C" int32_t mp377_profile_enable(mp377_ctx* ctx, int32_t on) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  msm_profile_enable(ctx->ws, on != 0);
  return MP_OK;
}
extern "C" int32_t mp377_profile_collect(mp377_ctx* ctx, double* accumulate_ms, uint64_t* bucket_adds, uint64_t* launches) {
  if (!ctx || !accumulate_ms || !bucket_adds || !launches) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  CK377(msm_profile_collect(ctx->ws, accumulate_ms, bucket_adds, launches), "mp377_profile_collect");
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// Pedersen commitments over a constant key (PedersenCommitment::{setup, commit}; reference type at
// src/discrete_log_cards/mod.rs:18,89, setup call mod.rs:111): fixed-base table mode of msm.cu
// ------------------------------------------------------------------------------------------
extern "C" int32_t mp377_set_commit_key(mp377_ctx* ctx, const uint8_t* ck, uint64_t len) {
  if (!ctx || !ck) return MP_ERR_INVALID_ARG;
  if (len == 0 || len >= (1u << 24)) return ctx->fail(MP_ERR_INVALID_ARG, "commit key length %llu out of range", (unsigned long long)len);
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const uint32_t nb = (uint32_t)len + 1;
  const int c = msm_pick_table_window(nb);
  const int W = msm_num_windows(c);
  uint8_t* d_in = (uint8_t*)ctx->scratch(mp377_ctx::kStageIn, (size_t)nb * kPt);
  affine* bases = (affine*)ctx->scratch(mp377_ctx::kCkBases, sizeof(affine) * nb);
  affine* table = (affine*)ctx->scratch(mp377_ctx::kCkTable, sizeof(affine) * nb * W);
  int* bad = (int*)ctx->scratch(mp377_ctx::kFlags, 256);
  if (!d_in || !bases || !table || !bad) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  ctx->ck_nb = 0;
  CK377(cudaMemcpyAsync(d_in, ck, (size_t)nb * kPt, cudaMemcpyHostToDevice, ctx->stream), "H2D commit key");
  CK377(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream), "memset");
  CK377(points_to_mont((const uint32_t*)d_in, bases, nb, bad, ctx->stream), "points_to_mont");
  CK377(msm_build_table(ctx->ws, bases, nb, 0, nb, c, table, ctx->stream), "msm_build_table");
  ctx->launches = 3;
  int h_bad = 0;
  CK377(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "D2H flag");
  CK377(cudaStreamSynchronize(ctx->stream), "commit key table");
  if (h_bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a commit-key point is not a canonical point of BLS12-377 G1");
  ctx->ck_nb = nb;
  ctx->ck_c = c;
  return MP_OK;
}

extern "C" int32_t mp377_pedersen_commit_batch(mp377_ctx* ctx, const uint8_t* values, const uint8_t* blinds, uint64_t k,
                                               uint64_t len, uint8_t* out) {
  if (!ctx || (k && (!blinds || !out)) || (k && len && !values)) return MP_ERR_INVALID_ARG;
  if (!ctx->ck_nb) return ctx->fail(MP_ERR_NO_PARAMS, "mp377_set_commit_key has not been called");
  if (len + 1 > ctx->ck_nb) return ctx->fail(MP_ERR_INVALID_ARG, "vector length %llu exceeds the commit key length %u",
                                             (unsigned long long)len, ctx->ck_nb - 1);
  if (k == 0) return MP_OK;
  if (k * (len + 1) >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "too many commitments in one batch");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  // job j: scalars [blind_j, values_j[0 .. len)] against table columns [0, len]  (column 0 is h)
  const size_t row = (len + 1) * 32;
  std::vector<uint8_t> h_scal(k * row);
  std::vector<MsmJob> jobs(k);
  for (uint64_t j = 0; j < k; j++) {
    memcpy(h_scal.data() + j * row, blinds + j * 32, 32);
    if (len) memcpy(h_scal.data() + j * row + 32, values + j * len * 32, len * 32);
    jobs[j] = MsmJob{(uint32_t)(j * (len + 1)), 0u, (uint32_t)(len + 1)};
  }
  uint32_t* d_scal = (uint32_t*)ctx->scratch(mp377_ctx::kScal, k * row);
  xyzz* res = (xyzz*)ctx->scratch(mp377_ctx::kMsmOut, sizeof(xyzz) * k);
  uint8_t* d_out = (uint8_t*)ctx->scratch(mp377_ctx::kStageOut, kPt * k);
  if (!d_scal || !res || !d_out) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  CK377(cudaMemcpyAsync(d_scal, h_scal.data(), k * row, cudaMemcpyHostToDevice, ctx->stream), "H2D scalars");
  CK377(msm_run(ctx->ws, d_scal, k * (len + 1), (const affine*)ctx->bufs[mp377_ctx::kCkTable].ptr, 1, jobs.data(), (int)k,
                ctx->ck_c, res, ctx->stream, 0, -1, ctx->ck_nb), "msm_run (fixed-base)");
  ctx->launches += msm_last_launches(ctx->ws);
  CK377(xyzz_to_canonical(res, (uint32_t*)d_out, k, ctx->stream), "xyzz_to_canonical");
  ctx->launches += 1;
  CK377(cudaMemcpyAsync(out, d_out, kPt * k, cudaMemcpyDeviceToHost, ctx->stream), "D2H commitments");
  CK377(cudaStreamSynchronize(ctx->stream), "commit batch");
  return MP_OK;
}

// ------------------------------------------------------------------------------------------
// debug / parity helpers: device field and group arithmetic exposed one operation per thread, so that
// the PTX carry chains (which the host tests cannot reach) are checked against the oracle directly
// ------------------------------------------------------------------------------------------
__global__ void k377_dbg_fq_mul(const fq* a, const fq* b, fq* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fq_mul(a[i], b[i]);
}
__global__ void k377_dbg_point_add(const uint32_t* p, const uint32_t* q, uint32_t* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int PW = 2 * kFqLimbs;
  xyzz acc = xyzz_from_affine(affine_from_canonical(p + i * PW));
  xyzz_madd(acc, affine_from_canonical(q + i * PW));
  affine a = xyzz_to_affine(acc);
  if (affine_is_identity(a)) { for (int k = 0; k < PW; k++) out[i * PW + k] = 0; }
  else affine_to_canonical(a, out + i * PW);
}
__global__ void k377_dbg_scalar_mul(const uint32_t* p, const uint32_t* k, uint32_t* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int PW = 2 * kFqLimbs;
  affine P = affine_from_canonical(p + i * PW);
  xyzz acc = xyzz_identity();
  for (int bit = 255; bit >= 0; bit--) {
    acc = xyzz_dbl(acc);
    if ((k[i * 8 + (bit >> 5)] >> (bit & 31)) & 1) xyzz_madd(acc, P);
  }
  affine a = xyzz_to_affine(acc);
  if (affine_is_identity(a)) { for (int w = 0; w < PW; w++) out[i * PW + w] = 0; }
  else affine_to_canonical(a, out + i * PW);
}

template <typename K>
static int32_t dbg_map(mp377_ctx* ctx, const uint8_t* a, size_t abytes, const uint8_t* b, size_t bbytes, uint64_t n,
                       uint8_t* out, size_t obytes, K launch) {
  if (!ctx || !a || !b || !out) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  uint8_t* d = (uint8_t*)ctx->scratch(mp377_ctx::kStageIn, (abytes + bbytes + obytes) * n + 256);
  if (!d) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  uint8_t *da = d, *db = d + abytes * n, *dout = db + bbytes * n;
  cudaMemcpyAsync(da, a, abytes * n, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(db, b, bbytes * n, cudaMemcpyHostToDevice, ctx->stream);
  launch(da, db, dout);
  cudaMemcpyAsync(out, dout, obytes * n, cudaMemcpyDeviceToHost, ctx->stream);
  CK377(cudaStreamSynchronize(ctx->stream), "debug kernel");
  ctx->launches = 1;
  return MP_OK;
}
extern "C" int32_t mp377_dbg_fq_mul(mp377_ctx* ctx, const uint8_t* a, const uint8_t* b, uint64_t n, uint8_t* out) {
  return dbg_map(ctx, a, kFe, b, kFe, n, out, kFe, [&](uint8_t* da, uint8_t* db, uint8_t* dout) {
    k377_dbg_fq_mul<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const fq*)da, (const fq*)db, (fq*)dout, n);
  });
}
extern "C" int32_t mp377_dbg_point_add(mp377_ctx* ctx, const uint8_t* p, const uint8_t* q, uint64_t n, uint8_t* out) {
  return dbg_map(ctx, p, kPt, q, kPt, n, out, kPt, [&](uint8_t* da, uint8_t* db, uint8_t* dout) {
    k377_dbg_point_add<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const uint32_t*)da, (const uint32_t*)db, (uint32_t*)dout, n);
  });
}
extern "C" int32_t mp377_dbg_scalar_mul(mp377_ctx* ctx, const uint8_t* p, const uint8_t* k, uint64_t n, uint8_t* out) {
  return dbg_map(ctx, p, kPt, k, 32, n, out, kPt, [&](uint8_t* da, uint8_t* db, uint8_t* dout) {
    k377_dbg_scalar_mul<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>((const uint32_t*)da, (const uint32_t*)db, (uint32_t*)dout, n);
  });
}

// integer-pipe microbenchmarks of the 12-limb field: 0 = fq_mul, 1 = XYZZ mixed addition
__global__ void __launch_bounds__(128) k377_bench_fq_mul(uint32_t* out, int iters, uint32_t seed) {
  fq x = fq_one(), y = fq_r2();
  x.v[0] ^= seed + threadIdx.x;
  y.v[0] ^= blockIdx.x;
  x = fq_reduce_weak(x); y = fq_reduce_weak(y);
  for (int it = 0; it < iters; it++) { x = fq_mul(x, y); y = fq_mul(y, x); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x.v[0] ^ y.v[7];
}
__global__ void __launch_bounds__(128) k377_bench_madd(uint32_t* out, int iters, uint32_t seed) {
  // the generator in canonical form -> Montgomery; acc walks 2G, 3G, ... (no special cases hit)
  const uint32_t g[24] = {0xb21be9efu, 0xeab9b16eu, 0xffcd394eu, 0xd5481512u, 0xbd37cb5cu, 0x188282c8u,
                          0xaa9d41bbu, 0x85951e2cu, 0xbf87ff54u, 0xc8fc6225u, 0xfe740a67u, 0x008848deu,
                          0x559c8ea6u, 0xfd82de55u, 0x34a9591au, 0xc2fe3d36u, 0x4fb82305u, 0x6d182ad4u,
                          0xca3e52d9u, 0xbd7fb348u, 0x30afeec4u, 0x1f674f5du, 0xc5102effu, 0x01914a69u};
  affine P = affine_from_canonical(g);
  xyzz acc = xyzz_dbl_affine(P);
  if ((seed + threadIdx.x) == 0xffffffffu) acc = xyzz_dbl(acc);
  for (int it = 0; it < iters; it++) xyzz_madd(acc, P);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.X.v[0] ^ acc.ZZZ.v[3];
}
extern "C" int32_t mp377_dbg_bench(mp377_ctx* ctx, int32_t which, int32_t iters, float* ms, double* ops) {
  if (!ctx || !ms || !ops || iters <= 0) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  const int blocks = 148 * 8;
  uint32_t* d = (uint32_t*)ctx->scratch(mp377_ctx::kStageOut, sizeof(uint32_t) * blocks * 128);
  if (!d) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {  // first pass = warm-up
    cudaEventRecord(e0, ctx->stream);
    if (which == 0) { k377_bench_fq_mul<<<blocks, 128, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 128 * iters * 2; }
    else if (which == 1) { k377_bench_madd<<<blocks, 128, 0, ctx->stream>>>(d, iters, 1234u); *ops = (double)blocks * 128 * iters; }
    else { cudaEventDestroy(e0); cudaEventDestroy(e1); return MP_ERR_INVALID_ARG; }
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
