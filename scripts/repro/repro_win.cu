// TEST INFRASTRUCTURE (root-cause harness): k_reduce_win of msm.cu against the closed form
//   out = sum_s T_s + L * sum_s s * S_s
// evaluated on the host with the same field / group code: a frozen copy of round 1's kernel over by-reference and
// over by-value helpers (with the per-thread intermediates run / above written out), and the combine kernels msm.cu
// ships now (k_reduce_group, k_reduce_group_quad).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DMP_CURVE_BLS12_377 -Dmp=mp_bls12_377] \
//        --expt-relaxed-constexpr scripts/repro/repro_win.cu -o scripts/repro/repro_win_{377,stark}
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../mental-poker_b200/csrc/msm.cu"
using namespace mp;

// ---- frozen copy of round 1's k_reduce_win (msm.cu no longer has it: both curves now share k_reduce_group[_quad]).
// BYREF = the round-1 helpers, xyzz passed by reference to __noinline__ functions: miscompiled on the 12-limb build.
// !BYREF = the same kernel over by-value helpers: correct on both builds.
static constexpr int kWinThreads = 256;
__device__ __noinline__ void add_ref(xyzz& acc, const xyzz& q) { xyzz_add(acc, q); }
__device__ __noinline__ void dbl_ref(xyzz& acc) { acc = xyzz_dbl(acc); }
template <bool BYREF> __device__ __forceinline__ void t_add(xyzz& acc, const xyzz& q) { if (BYREF) add_ref(acc, q); else acc = xyzz_add_v(acc, q); }
template <bool BYREF> __device__ __forceinline__ void t_dbl(xyzz& acc) { if (BYREF) dbl_ref(acc); else acc = xyzz_dbl_v(acc); }
__device__ __forceinline__ xyzz xyzz_shfl_down(const xyzz& v, int delta) {
  xyzz r;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4 * kXyzzVec; i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta);
  return r;
}
template <bool BYREF>
__device__ xyzz block_sum_xyzz(xyzz v, xyzz* smem) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    xyzz o = xyzz_shfl_down(v, d);
    if (lane < d) t_add<BYREF>(v, o);
  }
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int w = 1; w < kWinThreads / 32; w++) t_add<BYREF>(v, smem[w]);
  __syncthreads();
  return v;
}
struct Trace { xyzz run, inc, above, U; };
template <bool BYREF>
__global__ void __launch_bounds__(kWinThreads) k_reduce_win_r1(const xyzz* __restrict__ segS, const xyzz* __restrict__ segT, uint32_t nseg,
                                                              uint32_t L, int ncomp, xyzz* __restrict__ win_out, Trace* tr) {
  __shared__ xyzz smem[kWinThreads / 32];
  const uint32_t comp = blockIdx.x % ncomp;
  const uint64_t win = blockIdx.x / ncomp;
  const xyzz* S = segS + win * nseg * ncomp + comp;
  const xyzz* T = segT + win * nseg * ncomp + comp;
  const uint32_t ipt = (nseg + kWinThreads - 1) / kWinThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  xyzz run = xyzz_identity(), lsum = xyzz_identity(), tsum = xyzz_identity();
  for (int q = (int)ipt - 1; q >= 0; q--) {
    uint32_t i = threadIdx.x * ipt + q;
    if (i + 1 < nseg) { xyzz a = xyzz_load(S + (uint64_t)(i + 1) * ncomp); t_add<BYREF>(run, a); }
    t_add<BYREF>(lsum, run);
    if (i < nseg) { xyzz tv = xyzz_load(T + (uint64_t)i * ncomp); t_add<BYREF>(tsum, tv); }
  }
  if (tr) tr[threadIdx.x].run = run;
  xyzz inc = run;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    xyzz o = xyzz_shfl_down(inc, d);
    if (lane + d < 32) t_add<BYREF>(inc, o);
  }
  if (lane == 0) smem[warp] = inc;
  xyzz above = xyzz_shfl_down(inc, 1);
  if (lane == 31) above = xyzz_identity();
  __syncthreads();
  for (int w = warp + 1; w < kWinThreads / 32; w++) t_add<BYREF>(above, smem[w]);
  __syncthreads();
  if (tr) tr[threadIdx.x].above = above;
  for (uint32_t k = 1; k < ipt; k <<= 1) t_dbl<BYREF>(above);
  t_add<BYREF>(lsum, above);
  xyzz U = block_sum_xyzz<BYREF>(lsum, smem);
  xyzz Tt = block_sum_xyzz<BYREF>(tsum, smem);
  if (threadIdx.x == 0) {
    for (uint32_t k = 1; k < L; k <<= 1) t_dbl<BYREF>(U);
    t_add<BYREF>(Tt, U);
    xyzz_store(win_out + blockIdx.x, Tt);
  }
}

static bool same_point(const xyzz& a, const xyzz& b) {
  affine x = xyzz_to_affine(a), y = xyzz_to_affine(b);
  return fq_eq_raw(x.x, y.x) && fq_eq_raw(x.y, y.y);
}

int main() {
#ifdef MP_CURVE_BLS12_377
  const char* gx = "008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef";
  const char* gy = "01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6";
#else
  const char* gx = "01ef15c18599971b7beced415a40f0c7deacfd9b0d1819e03d723d8bc943cfca";
  const char* gy = "005668060aa49730b7be4801df46ec62de53ecd11abe43a32873000c36e8dc1f";
#endif
  uint32_t w[2 * kFqLimbs];
  auto parse = [&](const char* h, uint32_t* o) { for (int i = 0; i < kFqLimbs; i++) { unsigned v; sscanf(h + (kFqLimbs - 1 - i) * 8, "%8x", &v); o[i] = v; } };
  parse(gx, w); parse(gy, w + kFqLimbs);
  affine G = affine_from_canonical(w);
  int total_bad = 0;
  for (uint32_t nseg : {64u, 128u, 512u}) {
    const uint32_t L = 16;
    std::vector<xyzz> S(nseg), T(nseg);
    xyzz cur = xyzz_dbl_affine(G);
    for (uint32_t s = 0; s < nseg; s++) {  // distinct multiples of G, a few identities mixed in
      xyzz_madd(cur, G); cur = xyzz_dbl(cur); S[s] = (s % 17 == 5) ? xyzz_identity() : cur;
      xyzz_madd(cur, G); T[s] = (s % 23 == 7) ? xyzz_identity() : cur;
    }
    // closed form on the host
    xyzz want = xyzz_identity(), runS = xyzz_identity(), wsum = xyzz_identity();
    for (uint32_t s = 0; s < nseg; s++) xyzz_add(want, T[s]);
    for (uint32_t s = nseg - 1; s >= 1; s--) { xyzz_add(runS, S[s]); xyzz_add(wsum, runS); }  // sum_s s*S_s
    for (uint32_t k = 1; k < L; k <<= 1) wsum = xyzz_dbl(wsum);
    xyzz_add(want, wsum);
    // expected per-thread values of the scan
    const uint32_t ipt = (nseg + kWinThreads - 1) / kWinThreads;
    std::vector<xyzz> e_run(kWinThreads, xyzz_identity()), e_excl(kWinThreads, xyzz_identity());
    for (uint32_t t = 0; t < (uint32_t)kWinThreads; t++)
      for (uint32_t q = 0; q < ipt; q++) { uint32_t i = t * ipt + q; if (i + 1 < nseg) xyzz_add(e_run[t], S[i + 1]); }
    for (int t = kWinThreads - 2; t >= 0; t--) { e_excl[t] = e_excl[t + 1]; xyzz_add(e_excl[t], e_run[t + 1]); }
    xyzz *dS, *dT, *dOut; Trace* dTr;
    cudaMalloc(&dS, sizeof(xyzz) * nseg); cudaMalloc(&dT, sizeof(xyzz) * nseg); cudaMalloc(&dOut, sizeof(xyzz) * 4);
    cudaMalloc(&dTr, sizeof(Trace) * kWinThreads);
    cudaMemset(dTr, 0, sizeof(Trace) * kWinThreads);  // only .run and .above are written by the kernel
    cudaMemcpy(dS, S.data(), sizeof(xyzz) * nseg, cudaMemcpyHostToDevice);
    cudaMemcpy(dT, T.data(), sizeof(xyzz) * nseg, cudaMemcpyHostToDevice);
    xyzz got;
    for (int mode = 0; mode < 2; mode++) {
      cudaMemset(dOut, 0, sizeof(xyzz) * 4);
      if (mode == 0) k_reduce_win_r1<true><<<1, kWinThreads>>>(dS, dT, nseg, L, 1, dOut, dTr);
      else k_reduce_win_r1<false><<<1, kWinThreads>>>(dS, dT, nseg, L, 1, dOut, nullptr);
      cudaError_t e = cudaMemcpy(&got, dOut, sizeof(xyzz), cudaMemcpyDeviceToHost);
      bool ok = same_point(got, want);
      if (mode == 1) total_bad += !ok;   // the by-value form must be right; the by-reference form documents the defect
      printf("limbs=%d nseg=%u round-1 k_reduce_win, helpers %s: %s (cuda=%s)\n", kFqLimbs, nseg,
             mode ? "BY VALUE    " : "BY REFERENCE", ok ? "ok" : "WRONG", cudaGetErrorString(e));
    }
    std::vector<Trace> tr(kWinThreads);
    cudaMemcpy(tr.data(), dTr, sizeof(Trace) * kWinThreads, cudaMemcpyDeviceToHost);
    int brun = 0, babove = 0, first_run = -1, first_above = -1;
    for (int t = 0; t < kWinThreads; t++) {
      if (!same_point(tr[t].run, e_run[t])) { if (first_run < 0) first_run = t; brun++; }
      if (!same_point(tr[t].above, e_excl[t])) { if (first_above < 0) first_above = t; babove++; }
    }
    printf("   trace: run wrong in %d threads (first %d), exclusive suffix `above` wrong in %d threads (first %d)\n", brun, first_run, babove, first_above);
    // the group-fold path on the same data
    {
      xyzz *ping, *pong; size_t lvl = nseg / 4;
      cudaMalloc(&ping, sizeof(xyzz) * 2 * lvl); cudaMalloc(&pong, sizeof(xyzz) * 2 * lvl);
      const xyzz *curS = dS, *curT = dT; uint32_t cur_nseg = nseg, cur_L = L;
      for (int level = 0; cur_nseg > 1; level++) {
        const uint32_t Gp = cur_nseg < 4 ? cur_nseg : 4; xyzz* buf = (level & 1) ? pong : ping; const bool last = cur_nseg == Gp;
        k_reduce_group<<<(cur_nseg / Gp + 63) / 64, 64>>>(curS, curT, 1, cur_nseg, Gp, cur_L, 1, buf, last ? dOut : buf + lvl);
        curS = buf; curT = buf + lvl; cur_nseg /= Gp; cur_L *= Gp;
      }
      cudaMemcpy(&got, dOut, sizeof(xyzz), cudaMemcpyDeviceToHost);
      printf("   k_reduce_group levels: %s\n", same_point(got, want) ? "ok" : "WRONG");
      total_bad += !same_point(got, want);
      curS = dS; curT = dT; cur_nseg = nseg; cur_L = L;
      for (int level = 0; cur_nseg > 1; level++) {
        const uint32_t Gp = cur_nseg < 4 ? cur_nseg : 4; xyzz* buf = (level & 1) ? pong : ping; const bool last = cur_nseg == Gp;
        k_reduce_group_quad<<<(cur_nseg / Gp * 4 + 127) / 128, 128>>>(curS, curT, 1, cur_nseg, Gp, cur_L, 1, buf, last ? dOut : buf + lvl);
        curS = buf; curT = buf + lvl; cur_nseg /= Gp; cur_L *= Gp;
      }
      cudaMemcpy(&got, dOut, sizeof(xyzz), cudaMemcpyDeviceToHost);
      printf("   k_reduce_group_quad levels: %s\n", same_point(got, want) ? "ok" : "WRONG");
      total_bad += !same_point(got, want);
    }
  }
  printf(total_bad ? "FAIL %d\n" : "PASS\n", total_bad);
  return 0;
}
