// Karatsuba evaluation of the prover's diagonal ciphertext products E_k (kernel family K2 of
// SURVEY.md 2b; multi-exponentiation argument, Appendix B.5; reference call site
// src/discrete_log_cards/mod.rs:409-415).  diag_plan.hpp holds the combinatorics; this file the
// device side:
//
//   k_kara_points    leaf point rows   PL[leaf][t] = sum_{u in U(leaf)} deck'[(m-1-u)*2n + t]   (mixed adds)
//   k_batch_to_affine  XYZZ -> affine Montgomery with one inversion per 16 points (Montgomery's trick)
//   k_kara_scalars   leaf scalar rows  SL[leaf][l] = sum_{u in U(leaf)} b_{u+1}[l]  (mod group order)
//   msm_run          one batched Pippenger launch sequence: 3^L + m jobs of n terms, 2 components
//   k_kara_combine   E_k = sum of +-R_job over the plan's signed contribution list
//
// The point rows depend on no challenge: diag_karatsuba_points() is queued right after the deck
// upload and runs while the host hashes the statement.
#include "comm.cuh"
#include "diag.cuh"

#include "shuffle_internal.cuh"

namespace mp {

struct DiagDevice {
  DiagPlan plan;
  uint32_t *d_mask = nullptr, *d_val = nullptr, *d_row_start = nullptr, *d_entries = nullptr;
  ~DiagDevice() {
    if (d_mask) cudaFree(d_mask);
    if (d_val) cudaFree(d_val);
    if (d_row_start) cudaFree(d_row_start);
    if (d_entries) cudaFree(d_entries);
  }
};
void diag_device_destroy(DiagDevice* d) { delete d; }

static __device__ __forceinline__ xyzz ld_xyzz(const xyzz* p) {
  xyzz r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = s[i];
  return r;
}
static __device__ __forceinline__ void st_xyzz(xyzz* p, const xyzz& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = s[i];
}
static __device__ __forceinline__ affine ld_affine(const affine* p) {
  affine r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < 4; i++) d[i] = __ldg(s + i);
  return r;
}
// by value: see msm.cu xyzz_add_v (NVVM merges stack slots of by-pointer noinline arguments)
static __device__ __noinline__ xyzz xyzz_add_val(const xyzz acc, const xyzz q) { xyzz r = acc; xyzz_add(r, q); return r; }
#define xyzz_add_call(acc, q) ((acc) = xyzz_add_val((acc), (q)))

// Block b: leaf = b / chunks, 128 consecutive (column, component) slots of that leaf's row.  Rows of
// the deck are 2n consecutive affine points, so a warp reads 32 consecutive 64-byte records per step.
__global__ void __launch_bounds__(128) k_kara_points(const affine* __restrict__ deck, const uint32_t* __restrict__ leaf_mask,
                                                     const uint32_t* __restrict__ leaf_val, uint32_t m, uint32_t M,
                                                     uint32_t n2, uint32_t chunks, xyzz* __restrict__ out) {
  const uint32_t leaf = blockIdx.x / chunks;
  const uint32_t t = (blockIdx.x % chunks) * 128 + threadIdx.x;
  if (t >= n2) return;
  const uint32_t val = leaf_val[leaf], free_bits = ~leaf_mask[leaf] & (M - 1);
  xyzz acc = xyzz_identity();
  uint32_t sub = 0;
  do {  // all u = val | sub, sub a subset of the free bits
    const uint32_t u = val | sub;
    if (u < m) {
      affine p = ld_affine(deck + (size_t)(m - 1 - u) * n2 + t);
      xyzz_madd(acc, p);
    }
    sub = (sub - free_bits) & free_bits;
  } while (sub != 0);
  st_xyzz(out + (size_t)leaf * n2 + t, acc);
}

// Thread i normalises entries i, i + T, i + 2T, ... (kGroup of them) with ONE field inversion.
static constexpr int kGroup = 16;
__global__ void __launch_bounds__(128) k_batch_to_affine(const xyzz* __restrict__ in, affine* __restrict__ out,
                                                         uint64_t count, uint64_t T) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  fq pre[kGroup];  // pre[k] = z_0 * .. * z_k, z = ZZ * ZZZ (1 for the identity)
  fq acc = fq_one();
#pragma unroll 1
  for (int k = 0; k < kGroup; k++) {
    const uint64_t e = i + (uint64_t)k * T;
    if (e < count) {
      xyzz p = ld_xyzz(in + e);
      if (!xyzz_is_identity(p)) acc = fq_mul(acc, fq_mul(p.ZZ, p.ZZZ));
    }
    pre[k] = acc;
  }
  fq inv = fq_inv(acc);
#pragma unroll 1
  for (int k = kGroup - 1; k >= 0; k--) {
    const uint64_t e = i + (uint64_t)k * T;
    if (e >= count) continue;
    xyzz p = ld_xyzz(in + e);
    affine a;
    if (xyzz_is_identity(p)) {
      a.x = fq_zero();
      a.y = fq_zero();
    } else {
      fq z = fq_mul(p.ZZ, p.ZZZ);
      fq iz = k > 0 ? fq_mul(inv, pre[k - 1]) : inv;  // 1 / z_k
      inv = fq_mul(inv, z);                            // drop z_k from the running inverse
      a.x = fq_reduce_full(fq_mul(p.X, fq_mul(iz, p.ZZZ)));
      a.y = fq_reduce_full(fq_mul(p.Y, fq_mul(iz, p.ZZ)));
    }
    uint4* d = reinterpret_cast<uint4*>(out + e);
    const uint4* s = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int q = 0; q < 4; q++) d[q] = s[q];
  }
}

// rows: canonical scalars, row j at rows + j*n*8 words (row 0 = a0, row v+1 = S_v)
__global__ void __launch_bounds__(256) k_kara_scalars(const uint32_t* __restrict__ rows, const uint32_t* __restrict__ leaf_mask,
                                                      const uint32_t* __restrict__ leaf_val, uint32_t m, uint32_t M,
                                                      uint32_t n, uint64_t total, uint32_t* __restrict__ out) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const uint32_t leaf = (uint32_t)(g / n), l = (uint32_t)(g % n);
  const uint32_t val = leaf_val[leaf], free_bits = ~leaf_mask[leaf] & (M - 1);
  fr acc = fr_zero();
  uint32_t sub = 0;
  do {
    const uint32_t u = val | sub;
    if (u < m) {
      const uint4* s = reinterpret_cast<const uint4*>(rows + ((size_t)(u + 1) * n + l) * 8);
      uint4 lo = __ldg(s), hi = __ldg(s + 1);
      fr v;
      v.v[0] = lo.x; v.v[1] = lo.y; v.v[2] = lo.z; v.v[3] = lo.w;
      v.v[4] = hi.x; v.v[5] = hi.y; v.v[6] = hi.z; v.v[7] = hi.w;
      acc = fr_add(acc, v);  // residues below the group order: same addition in canonical form
    }
    sub = (sub - free_bits) & free_bits;
  } while (sub != 0);
  uint4* d = reinterpret_cast<uint4*>(out + g * 8);
  d[0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
  d[1] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
}

// Block per (k, component): signed sum of the contributing job results (R index = job*2 + comp).
static constexpr int kCombThreads = 128;
__global__ void __launch_bounds__(kCombThreads) k_kara_combine(const xyzz* __restrict__ R, const uint32_t* __restrict__ row_start,
                                                               const uint32_t* __restrict__ entries, xyzz* __restrict__ E) {
  __shared__ xyzz smem[kCombThreads];
  const uint32_t k = blockIdx.x >> 1, comp = blockIdx.x & 1;
  xyzz acc = xyzz_identity();
  for (uint32_t e = row_start[k] + threadIdx.x; e < row_start[k + 1]; e += kCombThreads) {
    const uint32_t ent = entries[e];
    xyzz v = ld_xyzz(R + (size_t)(ent & 0x7fffffffu) * 2 + comp);
    if (ent >> 31) v = xyzz_neg(v);
    xyzz_add_call(acc, v);
  }
  smem[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kCombThreads / 2; s >= 1; s >>= 1) {
    if ((int)threadIdx.x < s) {
      xyzz a = smem[threadIdx.x], b = smem[threadIdx.x + s];
      xyzz_add_call(a, b);
      smem[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_xyzz(E + blockIdx.x, smem[0]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int32_t diag_device(mp_ctx* ctx, DiagDevice** out) {
  ShuffleState* S = ctx->shuffle;
  if (S->diag && S->diag->plan.m == S->m) { *out = S->diag; return MP_OK; }
  delete S->diag;
  S->diag = nullptr;
  DiagDevice* d = new DiagDevice();
  d->plan = diag_plan_build(S->m);
  const DiagPlan& p = d->plan;
  auto up = [&](uint32_t** dst, const std::vector<uint32_t>& v) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(v.size(), 1) * 4);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
  };
  cudaError_t e = up(&d->d_mask, p.leaf_mask);
  if (e == cudaSuccess) e = up(&d->d_val, p.leaf_val);
  if (e == cudaSuccess) e = up(&d->d_row_start, p.row_start);
  if (e == cudaSuccess) e = up(&d->d_entries, p.entries);
  if (e != cudaSuccess) { delete d; return ctx->cuda_fail(e, "diag plan upload"); }
  S->diag = d;
  *out = d;
  return MP_OK;
}

bool diag_karatsuba_selected(int m, int n, int c_table) {
  if (const char* e = getenv("MP_DIAG_KARATSUBA")) return atoi(e) != 0 && m >= 2;
  if (m < 8) return false;
  int L = 0;
  while ((1 << L) < m) L++;
  uint64_t leaves = 1;
  for (int b = 0; b < L; b++) leaves *= 3;
  const int c_leaf = msm_pick_window((uint64_t)n, leaves + m);
  return diag_use_karatsuba(m, n, msm_num_windows(c_table), msm_num_windows(c_leaf), c_leaf);
}

int32_t diag_karatsuba_points(mp_ctx* ctx, const affine* d_deck2, cudaStream_t st) {
  ShuffleState* S = ctx->shuffle;
  DiagDevice* D;
  int32_t rc = diag_device(ctx, &D);
  if (rc != MP_OK) return rc;
  const uint32_t m = (uint32_t)S->m, n2 = 2u * (uint32_t)S->n, nleaf = D->plan.nleaf();
  const uint64_t count = (uint64_t)nleaf * n2;
  xyzz* tmp = (xyzz*)ctx->scratch(sKaraTmp, count * sizeof(xyzz));
  affine* pts = (affine*)ctx->scratch(sKaraPts, count * sizeof(affine));
  NEED(tmp); NEED(pts);
  const uint32_t chunks = (n2 + 127) / 128;
  if ((uint64_t)nleaf * chunks >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "deck too large for the diagonal plan");
  k_kara_points<<<nleaf * chunks, 128, 0, st>>>(d_deck2, D->d_mask, D->d_val, m, 1u << D->plan.levels, n2, chunks, tmp);
  const uint64_t T = (count + kGroup - 1) / kGroup;
  k_batch_to_affine<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(tmp, pts, count, T);
  CK(cudaGetLastError());
  ctx->launches += 2;
  return MP_OK;
}

int32_t diag_karatsuba_products(mp_ctx* ctx, const uint32_t* d_rows_canon, xyzz* d_E, cudaStream_t st, MsmWorkspace* ws) {
  ShuffleState* S = ctx->shuffle;
  DiagDevice* D;
  int32_t rc = diag_device(ctx, &D);
  if (rc != MP_OK) return rc;
  const DiagPlan& P = D->plan;
  const uint32_t m = (uint32_t)S->m, n = (uint32_t)S->n, nleaf = P.nleaf();
  const uint64_t nscal = (uint64_t)(nleaf + 1) * n;  // leaf rows, then a0
  const affine* pts = (const affine*)ctx->scratch(sKaraPts, (uint64_t)nleaf * 2 * n * sizeof(affine));
  uint32_t* scal = (uint32_t*)ctx->scratch(sKaraScal, nscal * 32);
  // One large proof across GPUs (mp_shuffle_and_remask_multi): the leaf products are independent
  // MSMs, so rank r of G evaluates jobs [r * per, (r + 1) * per) and the 256-byte results are all-gathered in place
  // on this stream; everything else of the proof runs identically on every rank.
  const int G = comm_collective(ctx) ? comm_size(ctx) : 1, rank = G > 1 ? comm_rank(ctx) : 0;
  const uint64_t njobs = P.njobs(), per = (njobs + G - 1) / G;
  xyzz* R = (xyzz*)ctx->scratch(sKaraOut, per * G * 2 * sizeof(xyzz));
  NEED(pts); NEED(scal); NEED(R);
  const uint64_t total = (uint64_t)nleaf * n;
  k_kara_scalars<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_rows_canon, D->d_mask, D->d_val, m, 1u << P.levels, n, total, scal);
  CK(cudaMemcpyAsync(scal + total * 8, d_rows_canon, (size_t)n * 32, cudaMemcpyDeviceToDevice, st));
  ctx->launches += 1;
  std::vector<MsmJob> jobs((size_t)njobs);
  for (uint32_t l = 0; l < nleaf; l++) jobs[l] = MsmJob{l * n, l * n, n};
  for (uint32_t i = 1; i <= m; i++) jobs[nleaf + i - 1] = MsmJob{nleaf * n, P.single[m - i] * n, n};  // <C_i, a0>, C_i = P_{m-i}
  const int c = msm_pick_window((uint64_t)n, jobs.size());
  // a launch sequence stays below the 2^32-entry limit of the sort (entries = terms * windows)
  const uint64_t max_jobs = std::max<uint64_t>(1, ((1ull << 31) / (uint64_t)msm_num_windows(c)) / n);
  const uint64_t mine0 = std::min<uint64_t>(njobs, rank * per), mine1 = std::min<uint64_t>(njobs, (rank + 1) * per);
  for (uint64_t j0 = mine0; j0 < mine1; j0 += max_jobs) {
    const uint64_t cnt = std::min<uint64_t>(max_jobs, mine1 - j0);
    CK(msm_run(ws, scal, nscal, pts, 2, jobs.data() + j0, (int)cnt, c, R + 2 * j0, st));
    ctx->launches += msm_last_launches(ws);
  }
  if (G > 1) {
    int32_t rc2 = comm_allgather(ctx, R, per * 2 * sizeof(xyzz), st);
    if (rc2 != MP_OK) return rc2;
  }
  k_kara_combine<<<4 * m, kCombThreads, 0, st>>>(R, D->d_row_start, D->d_entries, d_E);
  CK(cudaGetLastError());
  ctx->launches += 1;
  return MP_OK;
}

}  // namespace mp
