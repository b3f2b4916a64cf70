// Batched sigma protocols either side of the shuffle in a Barnett-Smart round (SURVEY.md section
// 8(f), rank 1): mask / remask / reveal with their Chaum-Pedersen proofs and the Schnorr key-ownership
// proof, for n independent items per call.  Reference call sites: src/discrete_log_cards/mod.rs:132-354
// (one proof per call, each 2-4 scalar multiplications inside the un-vendored proof-essentials crate).
//
// Division of labour as for the shuffle: the host owns the per-proof Fiat-Shamir transcripts
// (sigma_host.hpp; n small Blake2s + ChaCha20 evaluations spread over host threads) and the response
// scalars; every group operation runs in ONE kernel family, k_lincomb: a thread evaluates
//     out = k0*P0 + k1*P1 + f0*g + f1*pk + A - S
// with P0, P1, A, S points of a per-call arena (variable bases: signed 4-bit windows over both scalars,
// one shared doubling chain), g / pk through the 8-bit fixed-base window tables the remask kernel already uses (32 table
// additions per scalar).  Provers read results back as canonical points; verifiers only need "is it
// the identity", so their check jobs never invert.  Jobs are laid out kind-major (all c1 jobs, then
// all c2 jobs, ...) so a warp runs one shape.
#include <thread>

#include "shuffle_internal.cuh"
#include "sigma_host.hpp"

namespace mp {

static constexpr uint32_t kNone = 0xffffffffu;
static constexpr int kLcPointWords = 2 * kFqLimbs;  // canonical x || y of one point, 32-bit words
static constexpr size_t PB = kPointBytes, CB = kCtBytes;  // C-ABI point / ciphertext: 64 / 128 bytes (Stark), 96 / 192 (BLS12-377)
struct LcJob {
  uint32_t var_pt[2];  // arena indices of the variable bases (kNone = unused)
  uint32_t var_sc[2];  // scalar indices for them
  uint32_t fix_sc[2];  // scalar indices for the fixed bases g (0) and pk (1)
  uint32_t add_pt, sub_pt;  // arena points added with coefficient +1 / -1
  uint32_t out_pt;     // arena slot that receives the (affine) result, or kNone
};
// Jobs are regular: job i of a kind has field = base + stride * i (or none).  The host describes each
// kind with 18 words and k_make_jobs expands them on the device -- the first version built and uploaded
// 36 bytes per job (14 MB for a deck's worth of mask proofs).
enum LcField { fVarPt0, fVarPt1, fVarSc0, fVarSc1, fFixSc0, fFixSc1, fAddPt, fSubPt, fOutPt, kLcFields };
struct JobKind {
  uint32_t base[kLcFields], stride[kLcFields];
  JobKind() {
    for (int f = 0; f < kLcFields; f++) { base[f] = kNone; stride[f] = 0; }
  }
  JobKind& set(LcField f, uint64_t b, uint32_t st = 1) {
    base[f] = (uint32_t)b;
    stride[f] = st;
    return *this;
  }
};
static_assert(sizeof(LcJob) == kLcFields * sizeof(uint32_t), "LcJob is an array of its fields");

__global__ void __launch_bounds__(256) k_make_jobs(const JobKind* __restrict__ kinds, uint32_t nkinds, uint32_t n,
                                                   uint32_t* __restrict__ jobs) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)nkinds * n) return;
  const uint32_t k = (uint32_t)(g / n), i = (uint32_t)(g % n);
#pragma unroll
  for (int f = 0; f < kLcFields; f++) {
    const uint32_t b = kinds[k].base[f];
    jobs[g * kLcFields + f] = b == kNone ? kNone : b + kinds[k].stride[f] * i;
  }
}

static constexpr int kTabWin8 = 32, kTabDigits8 = 255, kTabSize8 = kTabWin8 * kTabDigits8;  // layout of ShuffleState::d_tab

__device__ __forceinline__ void ld_scalar(const uint32_t* scal, uint32_t idx, uint32_t k[8]) {
  const uint4* p = reinterpret_cast<const uint4*>(scal + (size_t)idx * 8);
  uint4 lo = __ldg(p), hi = __ldg(p + 1);
  k[0] = lo.x; k[1] = lo.y; k[2] = lo.z; k[3] = lo.w; k[4] = hi.x; k[5] = hi.y; k[6] = hi.z; k[7] = hi.w;
}
__device__ __forceinline__ affine ld_point(const affine* p) {
  affine r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(affine) / 16); i++) d[i] = s[i];
  return r;
}
// out-of-line group operations, operands and result BY VALUE: see msm.cu xyzz_add_v (NVVM merges the stack slots of
// address-taken structs passed by pointer to noinline functions)
static __device__ __noinline__ xyzz madd_val(const xyzz acc, const affine q) { xyzz r = acc; xyzz_madd(r, q); return r; }
static __device__ __noinline__ xyzz dbl_val(const xyzz acc) { return xyzz_dbl(acc); }
static __device__ __noinline__ xyzz add_val(const xyzz acc, const xyzz q) { xyzz r = acc; xyzz_add(r, q); return r; }
#define madd_call(acc, q) ((acc) = madd_val((acc), (q)))
#define dbl_call(acc) ((acc) = dbl_val((acc)))
#define add_call(acc, q) ((acc) = add_val((acc), (q)))

// flags: 1 = write canonical bytes to out_canon[job], 2 = write identity flag to out_flag[job]
__global__ void __launch_bounds__(128) k_lincomb(const LcJob* __restrict__ jobs, uint32_t njobs, affine* __restrict__ arena,
                                                 const uint32_t* __restrict__ scal, const affine* __restrict__ tab, int flags,
                                                 uint32_t* __restrict__ out_canon, uint8_t* __restrict__ out_flag) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= njobs) return;
  const LcJob j = jobs[g];
  xyzz acc = xyzz_identity();
  if (j.var_pt[0] != kNone) {
    // Variable bases: fixed signed 4-bit windows, MSB first.  Every lane runs the same schedule -- four
    // doublings, then one table addition per base -- where a per-bit double-and-add makes the whole warp
    // pay for a mixed addition whenever ANY lane has the bit set (measured: 24.7 of 32 lanes active).
    // Per base: multiples 1..8 in XYZZ (1 KB of local memory), digits in [-7, 8] recoded once.
    const bool two = j.var_pt[1] != kNone;
    xyzz mult[2][8];     // mult[t][q] = (q + 1) * P_t
    uint32_t dig[2][8];  // 64 nibbles per scalar: (magnitude - 1) | sign << 3, or 0xf for a zero digit
    uint32_t hi = 0;     // bit t: scalar t carries out of nibble 63 (only a non-canonical 256-bit value can)
    int top = -1;
#pragma unroll 1
    for (int t = 0; t < (two ? 2 : 1); t++) {
      uint32_t k[8];
      ld_scalar(scal, j.var_sc[t], k);
      const affine P = ld_point(arena + j.var_pt[t]);
      xyzz cur = xyzz_from_affine(P);
      mult[t][0] = cur;
      xyzz dbl = cur;
      dbl_call(dbl);
      mult[t][1] = dbl;
#pragma unroll 1
      for (int q = 2; q < 8; q++) {
        madd_call(dbl, P);  // (q + 1) P = q P + P
        mult[t][q] = dbl;
      }
      uint32_t carry = 0;
#pragma unroll 1
      for (int w = 0; w < 64; w++) {
        uint32_t raw = ((k[w >> 3] >> ((w & 7) * 4)) & 0xfu) + carry;
        uint32_t enc;
        if (raw == 16u) { enc = 0xfu; carry = 1u; }                  // nibble f + carry: digit 0, carry on
        else if (raw > 8u) { enc = (16u - raw - 1u) | 8u; carry = 1u; }  // digit raw - 16 in [-7, -1]
        else { enc = raw ? raw - 1u : 0xfu; carry = 0u; }            // digit raw in [0, 8]
        if (enc != 0xfu && w > top) top = w;
        const uint32_t sh = (w & 7) * 4;
        if (sh == 0) dig[t][w >> 3] = 0;
        dig[t][w >> 3] |= enc << sh;
      }
      hi |= carry << t;
    }
    if (hi) {  // digit 64 = 1: start from P_t and run all 64 windows
      if (hi & 1u) add_call(acc, mult[0][0]);
      if (hi & 2u) add_call(acc, mult[1][0]);
      top = 63;
    }
#pragma unroll 1
    for (int w = top; w >= 0; w--) {
      if (!xyzz_is_identity(acc)) { dbl_call(acc); dbl_call(acc); dbl_call(acc); dbl_call(acc); }
#pragma unroll 1
      for (int t = 0; t < (two ? 2 : 1); t++) {
        const uint32_t enc = (dig[t][w >> 3] >> ((w & 7) * 4)) & 0xfu;
        if (enc != 0xfu) {
          xyzz e = mult[t][enc & 7u];
          if (enc & 8u) e = xyzz_neg(e);
          add_call(acc, e);
        }
      }
    }
  }
#pragma unroll 1
  for (int f = 0; f < 2; f++) {
    if (j.fix_sc[f] == kNone) continue;
    uint32_t k[8];
    ld_scalar(scal, j.fix_sc[f], k);
    const affine* T = tab + (size_t)f * kTabSize8;
#pragma unroll 1
    for (int w = 0; w < kTabWin8; w++) {
      const uint32_t d = (k[w >> 2] >> ((w & 3) * 8)) & 0xffu;
      if (d) {
        affine e;
        const uint4* s = reinterpret_cast<const uint4*>(T + w * kTabDigits8 + (d - 1));
        uint4* dst = reinterpret_cast<uint4*>(&e);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(affine) / 16); q++) dst[q] = __ldg(s + q);
        madd_call(acc, e);
      }
    }
  }
  if (j.add_pt != kNone) madd_call(acc, ld_point(arena + j.add_pt));
  if (j.sub_pt != kNone) madd_call(acc, affine_neg(ld_point(arena + j.sub_pt)));
  if (flags & 2) out_flag[g] = xyzz_is_identity(acc) ? 1 : 0;
  if ((flags & 1) || j.out_pt != kNone) {
    const affine r = xyzz_to_affine(acc);  // identity -> (0, 0)
    if (j.out_pt != kNone) {
      uint4* d = reinterpret_cast<uint4*>(arena + j.out_pt);
      const uint4* s = reinterpret_cast<const uint4*>(&r);
#pragma unroll
      for (int q = 0; q < (int)(sizeof(affine) / 16); q++) d[q] = s[q];
    }
    if (flags & 1) {
      uint32_t w[kLcPointWords];
      if (affine_is_identity(r)) {
#pragma unroll
        for (int q = 0; q < kLcPointWords; q++) w[q] = 0;
      } else {
        affine_to_canonical(r, w);
      }
      uint4* o = reinterpret_cast<uint4*>(out_canon + (size_t)g * kLcPointWords);
#pragma unroll
      for (int q = 0; q < kLcPointWords / 4; q++) o[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
  }
}

// Resident blocks per SM of k_lincomb, capped through an (unused) dynamic shared memory request.  A thread keeps up to
// 2 KB of multiples tables in local memory; at the 3 blocks per SM the registers allow, the resident threads' frames
// are 148 x 384 x 2.7 KB = 154 MB -- more than the L2 holds, so the table reads of the two-base jobs went to HBM
// (ncu, 65 536 verify_reveal jobs: 3.9 GB of DRAM traffic at 3 blocks per SM, 0.27 GB at 2, 0.05 GB at 1; round 1
// measured 8.6 GB).  The kernel is bound by its dependent multiplication chains, not by that traffic: the eight
// launches of the six batched calls take 36.9 / 36.7 / 46.0 ms at 3 / 2 / 1 blocks per SM, and a 128-register build at
// 4 blocks per SM was slower (and moved 8.4 GB).  Default 2; MP_LINCOMB_BLOCKS = 0 removes the cap.
static size_t lincomb_smem() {
  static const size_t bytes = [] {
    const char* e = getenv("MP_LINCOMB_BLOCKS");
    const int blocks = e ? atoi(e) : 2;
    if (blocks <= 0) return (size_t)0;
    const size_t b = (size_t)(227 * 1024) / (size_t)blocks - 1024;   // 1 KB per block is reserved by the system
    cudaFuncSetAttribute(k_lincomb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
    return b;
  }();
  return bytes;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
enum SigmaSlot { sSigCanon = sKaraOut + 1, sSigArena, sSigScal, sSigJobs, sSigOut, sSigFlags, sSigKinds, sSigStage0, sSigStage1, sSigStage2, sSigBadPts };

// One batched call: a point arena (uploaded canonical points -> Montgomery, plus reserved result slots),
// a scalar array, and launches of k_lincomb over job lists expanded on the device.  Caller buffers are
// pageable and often strided (160-byte proof records): they are uploaded whole, once, and picked apart
// by device-to-device 2D copies -- a strided host-to-device copy of 32-byte rows is an order of magnitude
// slower than the contiguous copy of the same records.
struct SigmaCall {
  mp_ctx* ctx;
  ShuffleState* S;
  cudaStream_t st;
  uint64_t n_in = 0, n_arena = 0, n_scal = 0;
  uint8_t* d_canon = nullptr;
  affine* d_arena = nullptr;
  uint32_t* d_scal = nullptr;
  int* d_bad = nullptr;
  int* d_bad_pts = nullptr;  // verifiers: one flag per input point, so that a malformed item fails alone

  int32_t init(mp_ctx* c, uint64_t in_points, uint64_t reserved_points, uint64_t scalars, bool per_point_flags = false) {
    ctx = c;
    S = c->shuffle;
    st = c->stream;
    n_in = in_points;
    n_arena = in_points + reserved_points;
    n_scal = scalars;
    d_canon = (uint8_t*)ctx->scratch(sSigCanon, n_in * PB + 64);
    d_arena = (affine*)ctx->scratch(sSigArena, n_arena * sizeof(affine) + 64);
    d_scal = (uint32_t*)ctx->scratch(sSigScal, n_scal * 32 + 64);
    d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
    NEED(d_canon); NEED(d_arena); NEED(d_scal); NEED(d_bad);
    CK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    if (per_point_flags && n_in) {
      d_bad_pts = (int*)ctx->scratch(sSigBadPts, n_in * sizeof(int));
      NEED(d_bad_pts);
      CK(cudaMemsetAsync(d_bad_pts, 0, n_in * sizeof(int), st));
    }
    return MP_OK;
  }
  // per-point "not a canonical point of the curve" flags of the ingested points (after the last run)
  int32_t bad_points(std::vector<int>& h) {
    h.assign(n_in, 0);
    if (!d_bad_pts || !n_in) return MP_OK;
    CK(cudaMemcpyAsync(h.data(), d_bad_pts, n_in * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(stream_wait(ctx, st));
    return MP_OK;
  }
  // uploads a caller buffer whole into staging slot `which` (0..2); *d receives the device copy
  int32_t stage(int which, const uint8_t* src, size_t bytes, const uint8_t** d) {
    uint8_t* p = (uint8_t*)ctx->scratch(sSigStage0 + which, bytes + 64);
    NEED(p);
    CK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st));
    *d = p;
    return MP_OK;
  }
  // `records` rows of `width` bytes, `pitch` bytes apart in a staged (device) buffer -> arena slot / scalar `first`
  int32_t points_from(uint64_t first, const uint8_t* d_src, uint64_t records, size_t width, size_t pitch) {
    if (!records) return MP_OK;
    CK(cudaMemcpy2DAsync(d_canon + first * PB, width, d_src, pitch, width, records, cudaMemcpyDeviceToDevice, st));
    return MP_OK;
  }
  int32_t scalars_from(uint64_t first, const uint8_t* d_src, uint64_t count, size_t pitch) {
    if (!count) return MP_OK;
    CK(cudaMemcpy2DAsync(d_scal + first * 8, 32, d_src, pitch, 32, count, cudaMemcpyDeviceToDevice, st));
    return MP_OK;
  }
  // contiguous caller buffers go straight to their place
  int32_t put_points(uint64_t first, const uint8_t* src, uint64_t count) {
    if (!count) return MP_OK;
    CK(cudaMemcpyAsync(d_canon + first * PB, src, count * PB, cudaMemcpyHostToDevice, st));
    return MP_OK;
  }
  int32_t put_scalars(uint64_t first, const uint8_t* src, uint64_t count) {
    if (!count) return MP_OK;
    CK(cudaMemcpyAsync(d_scal + first * 8, src, count * 32, cudaMemcpyHostToDevice, st));
    return MP_OK;
  }
  int32_t ingest() {  // canonical -> Montgomery, canonical + on-curve validation
    if (d_bad_pts) CK(points_to_mont_items((const uint32_t*)d_canon, d_arena, n_in, d_bad_pts, 1, st));
    else CK(points_to_mont((const uint32_t*)d_canon, d_arena, n_in, d_bad, st));
    ctx->launches += 1;
    return MP_OK;
  }
  // runs kinds.size() * n jobs (kind-major); canonical results (64 B per job) to h_canon and/or identity
  // flags to h_flags (host pointers, ideally pinned)
  int32_t run(const std::vector<JobKind>& kinds, uint64_t n, uint8_t* h_canon, uint8_t* h_flags) {
    const uint32_t nk = (uint32_t)kinds.size(), nj = (uint32_t)(nk * n);
    if (!nj) return MP_OK;
    JobKind* d_kinds = (JobKind*)ctx->scratch(sSigKinds, sizeof(JobKind) * 16);
    LcJob* d_jobs = (LcJob*)ctx->scratch(sSigJobs, (size_t)nj * sizeof(LcJob));
    uint32_t* d_out = h_canon ? (uint32_t*)ctx->scratch(sSigOut, (size_t)nj * PB) : nullptr;
    uint8_t* d_flags = h_flags ? (uint8_t*)ctx->scratch(sSigFlags, (size_t)nj + 64) : nullptr;
    NEED(d_kinds); NEED(d_jobs);
    if (nk > 16) return ctx->fail(MP_ERR_INVALID_ARG, "internal: too many job kinds");
    if (h_canon) NEED(d_out);
    if (h_flags) NEED(d_flags);
    CK(cudaMemcpyAsync(d_kinds, kinds.data(), sizeof(JobKind) * nk, cudaMemcpyHostToDevice, st));
    k_make_jobs<<<(nj + 255) / 256, 256, 0, st>>>(d_kinds, nk, (uint32_t)n, (uint32_t*)d_jobs);
    k_lincomb<<<(nj + 127) / 128, 128, lincomb_smem(), st>>>(d_jobs, nj, d_arena, d_scal, S->d_tab, (h_canon ? 1 : 0) | (h_flags ? 2 : 0),
                                                             d_out, d_flags);
    CK(cudaGetLastError());
    ctx->launches += 2;
    if (h_canon) CK(cudaMemcpyAsync(h_canon, d_out, (size_t)nj * PB, cudaMemcpyDeviceToHost, st));
    if (h_flags) CK(cudaMemcpyAsync(h_flags, d_flags, (size_t)nj, cudaMemcpyDeviceToHost, st));
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(stream_wait(ctx, st));
    if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "an input point is not a canonical point of the curve");
    return MP_OK;
  }
};

static int32_t sigma_begin(mp_ctx* ctx, uint64_t n) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  if (!ctx->shuffle || ctx->shuffle->m == 0) return ctx->fail(MP_ERR_NO_PARAMS, "mp_ctx_set_params has not been called");
  if (n >= (1ull << 27)) return ctx->fail(MP_ERR_INVALID_ARG, "batch too large");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  return MP_OK;
}
static int host_thread_count(int32_t host_threads, uint64_t n) {
  int t = host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency();
  t = std::max(1, std::min(t, 64));
  return (int)std::min<uint64_t>((uint64_t)t, std::max<uint64_t>(1, n / 64));
}
// fn(i) for i < n on `threads` host threads, contiguous ranges
template <typename F>
static void for_items(uint64_t n, int threads, F&& fn) {
  if (threads <= 1) {
    for (uint64_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t] {
      const uint64_t b = n * t / threads, e = n * (t + 1) / threads;
      for (uint64_t i = b; i < e; i++) fn(i);
    });
  for (auto& th : pool) th.join();
}

// ---- Chaum-Pedersen over the fixed parameters (g, pk): mask and remask ------------------------
// remask = false: statement (c1, c2 - card), outputs masked = (r g, card + r pk)       (mod.rs:182-214)
// remask = true : statement (a g, a pk) = out - in, outputs out = in + (a g, a pk)      (mod.rs:242-272)
static int32_t cp_fixed_prove(mp_ctx* ctx, bool remask, const uint8_t* pk, const uint8_t* in, const uint8_t* wit,
                              const uint8_t* omega, uint64_t n, uint8_t* out, uint8_t* proofs, int32_t host_threads) {
  int32_t rc = sigma_begin(ctx, n);
  if (rc != MP_OK) return rc;
  if (!pk || (n && (!in || !wit || !omega || !out || !proofs))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MP_OK;
  if ((rc = shuffle_ensure_pk_table(ctx, pk)) != MP_OK) return rc;
  const uint64_t per = remask ? 2 : 1;  // input points per item
  SigmaCall sc;
  if ((rc = sc.init(ctx, per * n, 0, 2 * n)) != MP_OK) return rc;
  if ((rc = sc.put_points(0, in, per * n)) != MP_OK) return rc;
  if ((rc = sc.put_scalars(0, wit, n)) != MP_OK) return rc;       // scalars [0, n): witness w
  if ((rc = sc.put_scalars(n, omega, n)) != MP_OK) return rc;     // scalars [n, 2n): omega
  if ((rc = sc.ingest()) != MP_OK) return rc;
  // kinds: 0 s0 = w g | 1 s1 = w pk | 2 a = omega g | 3 b = omega pk | 4 out.c1 | 5 out.c2
  //   mask:   out.c1 = s0 (no extra job), out.c2 = card + w pk
  //   remask: out.c1 = in.c1 + w g,       out.c2 = in.c2 + w pk
  std::vector<JobKind> kinds(remask ? 6 : 5);
  kinds[0].set(fFixSc0, 0);
  kinds[1].set(fFixSc1, 0);
  kinds[2].set(fFixSc0, n);
  kinds[3].set(fFixSc1, n);
  if (remask) {
    kinds[4].set(fFixSc0, 0).set(fAddPt, 0, 2);
    kinds[5].set(fFixSc1, 0).set(fAddPt, 1, 2);
  } else {
    kinds[4].set(fFixSc1, 0).set(fAddPt, 0);
  }
  ShuffleState* S = ctx->shuffle;
  uint8_t* res = pinned(S, kinds.size() * n * PB);
  if (!res) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  if ((rc = sc.run(kinds, n, res, nullptr)) != MP_OK) return rc;
  const char* seed = remask ? kSeedRemasking : kSeedMasking;
  const Transcript seeded(seed, strlen(seed));
  auto R = [&](int kind, uint64_t i) { return res + ((size_t)kind * n + i) * PB; };
  for_items(n, host_thread_count(host_threads, n), [&](uint64_t i) {
    const fr c = cp_challenge(seeded, S->enc_g, pk, R(0, i), R(1, i), R(2, i), R(3, i));
    const fr r = fr_add(fr_from_bytes(omega + 32 * i), fr_mul(c, fr_from_bytes(wit + 32 * i)));
    uint8_t* p = proofs + kCpProofLen * i;
    memcpy(p, R(2, i), PB);
    memcpy(p + PB, R(3, i), PB);
    fr_to_bytes(r, p + 2 * PB);
    if (remask) {
      memcpy(out + CB * i, R(4, i), PB);
      memcpy(out + CB * i + PB, R(5, i), PB);
    } else {
      memcpy(out + CB * i, R(0, i), PB);
      memcpy(out + CB * i + PB, R(4, i), PB);
    }
  });
  return MP_OK;
}

static int32_t cp_fixed_verify(mp_ctx* ctx, bool remask, const uint8_t* pk, const uint8_t* in, const uint8_t* out,
                               const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  int32_t rc = sigma_begin(ctx, n);
  if (rc != MP_OK) return rc;
  if (!pk || (n && (!in || !out || !proofs || !statuses))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MP_OK;
  if ((rc = shuffle_ensure_pk_table(ctx, pk)) != MP_OK) return rc;
  // arena: [inputs: n or 2n] [outputs: 2n] [a, b: 2n] | reserved statement slots [s0: n] [s1: n]
  const uint64_t per = remask ? 2 : 1;
  const uint64_t oIn = 0, oOut = per * n, oAB = oOut + 2 * n, oS = oAB + 2 * n;
  SigmaCall sc;
  if ((rc = sc.init(ctx, oS, 2 * n, 2 * n, true)) != MP_OK) return rc;
  const uint8_t* d_proofs;
  if ((rc = sc.put_points(oIn, in, per * n)) != MP_OK) return rc;
  if ((rc = sc.put_points(oOut, out, 2 * n)) != MP_OK) return rc;
  if ((rc = sc.stage(0, proofs, n * kCpProofLen, &d_proofs)) != MP_OK) return rc;
  if ((rc = sc.points_from(oAB, d_proofs, n, 2 * PB, kCpProofLen)) != MP_OK) return rc;
  if ((rc = sc.scalars_from(0, d_proofs + 2 * PB, n, kCpProofLen)) != MP_OK) return rc;  // scalars [0, n): responses r
  if ((rc = sc.ingest()) != MP_OK) return rc;
  // pass 1: the statement.  mask: s0 = out.c1 (no job), s1 = out.c2 - card;  remask: s = out - in
  std::vector<JobKind> kinds;
  if (remask) {
    kinds.resize(2);
    for (int k = 0; k < 2; k++) kinds[k].set(fAddPt, oOut + k, 2).set(fSubPt, oIn + k, 2).set(fOutPt, oS + (uint64_t)k * n);
  } else {
    kinds.resize(1);
    kinds[0].set(fAddPt, oOut + 1, 2).set(fSubPt, oIn).set(fOutPt, oS + n);
  }
  ShuffleState* S = ctx->shuffle;
  const size_t stmt_bytes = kinds.size() * n * PB;
  uint8_t* pin = pinned(S, stmt_bytes + 2 * n + 64);
  if (!pin) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  uint8_t *stmt = pin, *flags = pin + stmt_bytes;
  if ((rc = sc.run(kinds, n, stmt, nullptr)) != MP_OK) return rc;
  // challenges (host), uploaded negated:  r g - c s0 - a == O  and  r pk - c s1 - b == O
  const char* seed = remask ? kSeedRemasking : kSeedMasking;
  const Transcript seeded(seed, strlen(seed));
  std::vector<uint8_t> negc((size_t)n * 32), canon_ok((size_t)n);
  for_items(n, host_thread_count(host_threads, n), [&](uint64_t i) {
    const uint8_t* p = proofs + kCpProofLen * i;
    const uint8_t* s0 = remask ? stmt + PB * i : out + CB * i;
    const uint8_t* s1 = remask ? stmt + PB * (n + i) : stmt + PB * i;
    const fr c = cp_challenge(seeded, S->enc_g, pk, s0, s1, p, p + PB);
    fr_to_bytes(fr_neg(c), negc.data() + 32 * i);
    canon_ok[i] = fr_bytes_canonical(p + 2 * PB);
  });
  if ((rc = sc.put_scalars(n, negc.data(), n)) != MP_OK) return rc;  // scalars [n, 2n): -c
  kinds.assign(2, JobKind());
  kinds[0].set(fFixSc0, 0).set(fVarSc0, n).set(fSubPt, oAB, 2);
  if (remask) kinds[0].set(fVarPt0, oS); else kinds[0].set(fVarPt0, oOut, 2);
  kinds[1].set(fFixSc1, 0).set(fVarPt0, oS + n).set(fVarSc0, n).set(fSubPt, oAB + 1, 2);
  if ((rc = sc.run(kinds, n, nullptr, flags)) != MP_OK) return rc;
  std::vector<int> bad;
  if ((rc = sc.bad_points(bad)) != MP_OK) return rc;
  for (uint64_t i = 0; i < n; i++) {
    // an item with a point off the curve / not canonical fails on its own, as the reference's deserialiser would fail it
    bool malformed = bad[oOut + 2 * i] || bad[oOut + 2 * i + 1] || bad[oAB + 2 * i] || bad[oAB + 2 * i + 1];
    for (uint64_t k = 0; k < per; k++) malformed = malformed || bad[oIn + per * i + k];
    statuses[i] = malformed ? MP_VERIFY_MALFORMED : (flags[i] && flags[n + i] && canon_ok[i]) ? MP_OK : MP_VERIFY_CHAUM_PEDERSEN;
  }
  return MP_OK;
}

int32_t sigma_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* r, const uint8_t* omega,
                         uint64_t n, uint8_t* out_masked, uint8_t* out_proofs, int32_t host_threads) {
  return cp_fixed_prove(ctx, false, shared_key, cards, r, omega, n, out_masked, out_proofs, host_threads);
}
int32_t sigma_verify_mask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* cards, const uint8_t* masked,
                                const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  return cp_fixed_verify(ctx, false, shared_key, cards, masked, proofs, n, statuses, host_threads);
}
int32_t sigma_remask_prove_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* alpha,
                                 const uint8_t* omega, uint64_t n, uint8_t* out_deck, uint8_t* out_proofs, int32_t host_threads) {
  return cp_fixed_prove(ctx, true, shared_key, deck, alpha, omega, n, out_deck, out_proofs, host_threads);
}
int32_t sigma_verify_remask_batch(mp_ctx* ctx, const uint8_t* shared_key, const uint8_t* deck, const uint8_t* remasked,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  return cp_fixed_verify(ctx, true, shared_key, deck, remasked, proofs, n, statuses, host_threads);
}

// ---- reveal tokens of one player (mod.rs:301-354): parameters (masked.c1, g), statement (token, pk) ----
int32_t sigma_reveal_batch(mp_ctx* ctx, const uint8_t* sk, const uint8_t* pk, const uint8_t* masked, const uint8_t* omega,
                           uint64_t n, uint8_t* out_tokens, uint8_t* out_proofs, int32_t host_threads) {
  int32_t rc = sigma_begin(ctx, n);
  if (rc != MP_OK) return rc;
  if (!sk || !pk || (n && (!masked || !omega || !out_tokens || !out_proofs))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MP_OK;
  SigmaCall sc;
  if ((rc = sc.init(ctx, n, 0, n + 1)) != MP_OK) return rc;
  const uint8_t* d_masked;
  if ((rc = sc.stage(0, masked, n * CB, &d_masked)) != MP_OK) return rc;
  if ((rc = sc.points_from(0, d_masked, n, PB, CB)) != MP_OK) return rc;  // c1 of every card
  if ((rc = sc.put_scalars(0, omega, n)) != MP_OK) return rc;
  if ((rc = sc.put_scalars(n, sk, 1)) != MP_OK) return rc;
  if ((rc = sc.ingest()) != MP_OK) return rc;
  std::vector<JobKind> kinds(3);  // token = sk c1 | a = omega c1 | b = omega g
  kinds[0].set(fVarPt0, 0).set(fVarSc0, n, 0);
  kinds[1].set(fVarPt0, 0).set(fVarSc0, 0);
  kinds[2].set(fFixSc0, 0);
  ShuffleState* S = ctx->shuffle;
  uint8_t* res = pinned(S, 3 * n * PB);
  if (!res) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  if ((rc = sc.run(kinds, n, res, nullptr)) != MP_OK) return rc;
  const Transcript seeded(kSeedReveal, strlen(kSeedReveal));
  const fr skf = fr_from_bytes(sk);
  for_items(n, host_thread_count(host_threads, n), [&](uint64_t i) {
    const uint8_t *tok = res + PB * i, *a = res + PB * (n + i), *b = res + PB * (2 * n + i);
    const fr c = cp_challenge(seeded, masked + CB * i, S->enc_g, tok, pk, a, b);
    const fr r = fr_add(fr_from_bytes(omega + 32 * i), fr_mul(c, skf));
    memcpy(out_tokens + PB * i, tok, PB);
    uint8_t* p = out_proofs + kCpProofLen * i;
    memcpy(p, a, PB);
    memcpy(p + PB, b, PB);
    fr_to_bytes(r, p + 2 * PB);
  });
  return MP_OK;
}

int32_t sigma_verify_reveal_batch(mp_ctx* ctx, const uint8_t* pk, const uint8_t* tokens, const uint8_t* masked,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  int32_t rc = sigma_begin(ctx, n);
  if (rc != MP_OK) return rc;
  if (!pk || (n && (!tokens || !masked || !proofs || !statuses))) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MP_OK;
  // arena: [c1: n] [tokens: n] [a, b: 2n] [pk]
  const uint64_t oC1 = 0, oTok = n, oAB = 2 * n, oPk = 4 * n;
  SigmaCall sc;
  if ((rc = sc.init(ctx, 4 * n + 1, 0, 2 * n, true)) != MP_OK) return rc;
  const uint8_t *d_masked, *d_proofs;
  if ((rc = sc.stage(0, masked, n * CB, &d_masked)) != MP_OK) return rc;
  if ((rc = sc.stage(1, proofs, n * kCpProofLen, &d_proofs)) != MP_OK) return rc;
  if ((rc = sc.points_from(oC1, d_masked, n, PB, CB)) != MP_OK) return rc;
  if ((rc = sc.put_points(oTok, tokens, n)) != MP_OK) return rc;
  if ((rc = sc.points_from(oAB, d_proofs, n, 2 * PB, kCpProofLen)) != MP_OK) return rc;
  if ((rc = sc.put_points(oPk, pk, 1)) != MP_OK) return rc;
  if ((rc = sc.scalars_from(0, d_proofs + 2 * PB, n, kCpProofLen)) != MP_OK) return rc;  // scalars [0, n): r
  if ((rc = sc.ingest()) != MP_OK) return rc;
  ShuffleState* S = ctx->shuffle;
  const Transcript seeded(kSeedReveal, strlen(kSeedReveal));
  std::vector<uint8_t> negc((size_t)n * 32), canon_ok((size_t)n);
  for_items(n, host_thread_count(host_threads, n), [&](uint64_t i) {  // overlaps the uploads
    const uint8_t* p = proofs + kCpProofLen * i;
    const fr c = cp_challenge(seeded, masked + CB * i, S->enc_g, tokens + PB * i, pk, p, p + PB);
    fr_to_bytes(fr_neg(c), negc.data() + 32 * i);
    canon_ok[i] = fr_bytes_canonical(p + 2 * PB);
  });
  if ((rc = sc.put_scalars(n, negc.data(), n)) != MP_OK) return rc;  // scalars [n, 2n): -c
  std::vector<JobKind> kinds(2);  // r c1 - c token - a  |  r g - c pk - b
  kinds[0].set(fVarPt0, oC1).set(fVarSc0, 0).set(fVarPt1, oTok).set(fVarSc1, n).set(fSubPt, oAB, 2);
  kinds[1].set(fFixSc0, 0).set(fVarPt0, oPk, 0).set(fVarSc0, n).set(fSubPt, oAB + 1, 2);
  uint8_t* flags = pinned(S, 2 * n + 64);
  if (!flags) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  if ((rc = sc.run(kinds, n, nullptr, flags)) != MP_OK) return rc;
  std::vector<int> bad;
  if ((rc = sc.bad_points(bad)) != MP_OK) return rc;
  if (bad[oPk]) return ctx->fail(MP_ERR_NOT_ON_CURVE, "the public key is not a canonical point of the curve");
  for (uint64_t i = 0; i < n; i++) {
    const bool malformed = bad[oC1 + i] || bad[oTok + i] || bad[oAB + 2 * i] || bad[oAB + 2 * i + 1];
    statuses[i] = malformed ? MP_VERIFY_MALFORMED : (flags[i] && flags[n + i] && canon_ok[i]) ? MP_OK : MP_VERIFY_CHAUM_PEDERSEN;
  }
  return MP_OK;
}

// ---- Schnorr key ownership (mod.rs:132-165) --------------------------------------------------------
int32_t sigma_key_ownership_prove_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* sks, const uint8_t* infos,
                                        const uint64_t* info_off, const uint8_t* omega, uint64_t n, uint8_t* out_proofs,
                                        int32_t host_threads) {
  int32_t rc = sigma_begin(ctx, n);
  if (rc != MP_OK) return rc;
  if (n && (!pks || !sks || !info_off || !omega || !out_proofs)) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MP_OK;
  SigmaCall sc;
  if ((rc = sc.init(ctx, 0, 0, n)) != MP_OK) return rc;
  if ((rc = sc.put_scalars(0, omega, n)) != MP_OK) return rc;
  std::vector<JobKind> kinds(1);  // commit = omega g
  kinds[0].set(fFixSc0, 0);
  ShuffleState* S = ctx->shuffle;
  uint8_t* res = pinned(S, n * PB);
  if (!res) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  if ((rc = sc.run(kinds, n, res, nullptr)) != MP_OK) return rc;
  for_items(n, host_thread_count(host_threads, n), [&](uint64_t i) {
    const uint8_t* commit = res + PB * i;
    const fr c = schnorr_challenge(infos ? infos + info_off[i] : nullptr, (size_t)(info_off[i + 1] - info_off[i]), S->enc_g,
                                   pks + PB * i, commit);
    const fr op = fr_sub(fr_from_bytes(omega + 32 * i), fr_mul(c, fr_from_bytes(sks + 32 * i)));
    uint8_t* p = out_proofs + kSchnorrProofLen * i;
    memcpy(p, commit, PB);
    fr_to_bytes(op, p + PB);
  });
  return MP_OK;
}

int32_t sigma_key_ownership_verify_batch(mp_ctx* ctx, const uint8_t* pks, const uint8_t* infos, const uint64_t* info_off,
                                         const uint8_t* proofs, uint64_t n, int32_t* statuses, int32_t host_threads) {
  int32_t rc = sigma_begin(ctx, n);
  if (rc != MP_OK) return rc;
  if (n && (!pks || !info_off || !proofs || !statuses)) return ctx->fail(MP_ERR_INVALID_ARG, "null argument");
  if (n == 0) return MP_OK;
  SigmaCall sc;  // arena: [pk: n] [commit: n];  scalars: [opening: n] [c: n]
  if ((rc = sc.init(ctx, 2 * n, 0, 2 * n, true)) != MP_OK) return rc;
  const uint8_t* d_proofs;
  if ((rc = sc.put_points(0, pks, n)) != MP_OK) return rc;
  if ((rc = sc.stage(0, proofs, n * kSchnorrProofLen, &d_proofs)) != MP_OK) return rc;
  if ((rc = sc.points_from(n, d_proofs, n, PB, kSchnorrProofLen)) != MP_OK) return rc;
  if ((rc = sc.scalars_from(0, d_proofs + PB, n, kSchnorrProofLen)) != MP_OK) return rc;
  if ((rc = sc.ingest()) != MP_OK) return rc;
  ShuffleState* S = ctx->shuffle;
  std::vector<uint8_t> cs((size_t)n * 32), canon_ok((size_t)n);
  for_items(n, host_thread_count(host_threads, n), [&](uint64_t i) {
    const uint8_t* p = proofs + kSchnorrProofLen * i;
    const fr c = schnorr_challenge(infos ? infos + info_off[i] : nullptr, (size_t)(info_off[i + 1] - info_off[i]), S->enc_g,
                                   pks + PB * i, p);
    fr_to_bytes(c, cs.data() + 32 * i);
    canon_ok[i] = fr_bytes_canonical(p + PB);
  });
  if ((rc = sc.put_scalars(n, cs.data(), n)) != MP_OK) return rc;
  std::vector<JobKind> kinds(1);  // opening g + c pk - commit == O
  kinds[0].set(fFixSc0, 0).set(fVarPt0, 0).set(fVarSc0, n).set(fSubPt, n);
  uint8_t* flags = pinned(S, n + 64);
  if (!flags) return ctx->fail(MP_ERR_CUDA, "pinned allocation failed");
  if ((rc = sc.run(kinds, n, nullptr, flags)) != MP_OK) return rc;
  std::vector<int> bad;
  if ((rc = sc.bad_points(bad)) != MP_OK) return rc;
  for (uint64_t i = 0; i < n; i++)
    statuses[i] = (bad[i] || bad[n + i]) ? MP_VERIFY_MALFORMED : (flags[i] && canon_ok[i]) ? MP_OK : MP_VERIFY_SCHNORR;
  return MP_OK;
}

}  // namespace mp
