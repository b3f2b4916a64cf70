// TEST INFRASTRUCTURE (root-cause harness): runs the Montgomery-trick table normalisation
// of msm.cu on the device with every intermediate value written out, and the same word-level algorithm on the
// host (fq/ec code is __host__ __device__), and reports the first value that differs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMP_CURVE_BLS12_377 -Dmp=mp_bls12_377 \
//        --expt-relaxed-constexpr scripts/repro/repro_table.cu -o scripts/repro/repro_table_377
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../mental-poker_b200/csrc/msm.cu"
using namespace mp;

// round 1's calling convention of the out-of-line multiplication (pointers); fq_mul() now passes by value
#ifdef MP_CURVE_BLS12_377
static __device__ __noinline__ void fq_mul_ptr(fq* r, const fq* a, const fq* b) { const fq x = *a, y = *b; const fq z = fq_mul_inline(x, y); *r = z; }
#else
static __device__ __noinline__ void fq_mul_ptr(fq* r, const fq* a, const fq* b) { const fq x = *a, y = *b; const fq z = fq_mul(x, y); *r = z; }
#endif
template <bool BYPTR> __device__ __forceinline__ fq t_mul(const fq& a, const fq& b) {
  if (BYPTR) { fq r; fq_mul_ptr(&r, &a, &b); return r; }
  return fq_mul(a, b);
}

struct Dump { fq acc_after[64]; fq z[64]; fq inv0; fq iz[64]; fq inv_after[64]; affine a[64]; };

template <bool BYPTR>
__global__ void k_trace(const xyzz* tmp, int W, Dump* d) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  fq pre[64];
  fq acc = fq_one();
  for (int w = 0; w < W; w++) {
    xyzz p = xyzz_load(tmp + w);
    fq z = t_mul<BYPTR>(p.ZZ, p.ZZZ);
    d->z[w] = z;
    if (!xyzz_is_identity(p)) acc = t_mul<BYPTR>(acc, z);
    pre[w] = acc;
    d->acc_after[w] = acc;
  }
  fq inv = fq_inv(acc);
  d->inv0 = inv;
  for (int w = W - 1; w >= 0; w--) {
    xyzz p = xyzz_load(tmp + w);
    fq z = t_mul<BYPTR>(p.ZZ, p.ZZZ);
    fq iz = w > 0 ? t_mul<BYPTR>(inv, pre[w - 1]) : inv;
    inv = t_mul<BYPTR>(inv, z);
    d->iz[w] = iz;
    d->inv_after[w] = inv;
    d->a[w].x = fq_reduce_full(t_mul<BYPTR>(p.X, t_mul<BYPTR>(iz, p.ZZZ)));
    d->a[w].y = fq_reduce_full(t_mul<BYPTR>(p.Y, t_mul<BYPTR>(iz, p.ZZ)));
  }
}

__global__ void k_one_mul(const fq* a, const fq* b, fq* out) { *out = fq_mul(*a, *b); }
// the prefix-product loop with the running value kept in global memory (nothing loop-carried in registers)
__global__ void k_prefix_gmem(const xyzz* tmp, int W, fq* acc_io, fq* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  *acc_io = fq_one();
  for (int w = 0; w < W; w++) {
    xyzz p = xyzz_load(tmp + w);
    fq z = fq_mul(p.ZZ, p.ZZZ);
    fq a;
    for (int k = 0; k < kFqLimbs; k++) a.v[k] = ((volatile uint32_t*)acc_io->v)[k];
    fq r = fq_mul(a, z);
    for (int k = 0; k < kFqLimbs; k++) ((volatile uint32_t*)acc_io->v)[k] = r.v[k];
    out[w] = r;
  }
}

static bool eq(const fq& a, const fq& b) { return fq_eq_raw(a, b); }
static bool eqmod(const fq& a, const fq& b) { return fq_eq_raw(fq_reduce_full(a), fq_reduce_full(b)); }

int main() {
  // generator of the curve this file was compiled for, canonical little-endian words
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
  const int c = 13, W = (kScalarBits + c - 1) / c;
  std::vector<xyzz> tmp(W);
  xyzz cur = xyzz_from_affine(G);
  for (int i = 0; i < W; i++) { tmp[i] = cur; if (i + 1 < W) for (int k = 0; k < c; k++) cur = xyzz_dbl(cur); }
  // host run of the same algorithm
  Dump h; memset(&h, 0, sizeof h);
  {
    fq pre[64]; fq acc = fq_one();
    for (int i = 0; i < W; i++) { fq z = fq_mul(tmp[i].ZZ, tmp[i].ZZZ); h.z[i] = z; acc = fq_mul(acc, z); pre[i] = acc; h.acc_after[i] = acc; }
    fq inv = fq_inv(acc); h.inv0 = inv;
    for (int i = W - 1; i >= 0; i--) {
      fq z = fq_mul(tmp[i].ZZ, tmp[i].ZZZ); fq iz = i > 0 ? fq_mul(inv, pre[i - 1]) : inv; inv = fq_mul(inv, z);
      h.iz[i] = iz; h.inv_after[i] = inv;
      h.a[i].x = fq_reduce_full(fq_mul(tmp[i].X, fq_mul(iz, tmp[i].ZZZ)));
      h.a[i].y = fq_reduce_full(fq_mul(tmp[i].Y, fq_mul(iz, tmp[i].ZZ)));
    }
  }
  xyzz* d_tmp; Dump* d_dump;
  cudaMalloc(&d_tmp, sizeof(xyzz) * W); cudaMalloc(&d_dump, sizeof(Dump));
  cudaMemcpy(d_tmp, tmp.data(), sizeof(xyzz) * W, cudaMemcpyHostToDevice);
  cudaMemset(d_dump, 0, sizeof(Dump));
  int bad = 0;
  for (int byptr = 1; byptr >= 0; byptr--) {
    cudaMemset(d_dump, 0, sizeof(Dump));
    if (byptr) k_trace<true><<<1, 32>>>(d_tmp, W, d_dump); else k_trace<false><<<1, 32>>>(d_tmp, W, d_dump);
    Dump g; cudaError_t e = cudaMemcpy(&g, d_dump, sizeof(Dump), cudaMemcpyDeviceToHost);
    printf("curve limbs=%d W=%d, prefix-product trace with the multiplication helper taking %s (cuda=%s)\n", kFqLimbs, W,
           byptr ? "POINTERS (round 1)" : "VALUES (now)", cudaGetErrorString(e));
    int wrong = 0, first = -1;
    for (int i = 0; i < W; i++) {
      if (!eq(g.z[i], h.z[i])) { printf("  z[%d] words differ (mod q equal: %d)\n", i, (int)eqmod(g.z[i], h.z[i])); wrong++; }
      if (!eq(g.acc_after[i], h.acc_after[i])) { if (first < 0) first = i; wrong++; }
    }
    if (first >= 0) printf("  acc[%d] = acc[%d] * z[%d] is the first value that differs from the host (mod q equal: %d)\n", first, first - 1, first,
                           (int)eqmod(g.acc_after[first], h.acc_after[first]));
    if (!eq(g.inv0, h.inv0)) wrong++;
    for (int i = W - 1; i >= 0; i--) {
      if (!eq(g.iz[i], h.iz[i])) wrong++;
      if (!eq(g.inv_after[i], h.inv_after[i])) wrong++;
      if (!eq(g.a[i].x, h.a[i].x) || !eq(g.a[i].y, h.a[i].y)) wrong++;
    }
    printf("  %d traced values differ from the host\n", wrong);
    if (!byptr) bad += wrong;  // the by-value form must be right; the by-pointer form documents the defect
  }
  {  // one multiplication in isolation: the first pair the trace gets wrong
    fq *d_a, *d_b, *d_o; cudaMalloc(&d_a, sizeof(fq)); cudaMalloc(&d_b, sizeof(fq)); cudaMalloc(&d_o, sizeof(fq) * 64);
    for (int i = 1; i < 4; i++) {
      cudaMemcpy(d_a, &h.acc_after[i - 1], sizeof(fq), cudaMemcpyHostToDevice);
      cudaMemcpy(d_b, &h.z[i], sizeof(fq), cudaMemcpyHostToDevice);
      k_one_mul<<<1, 1>>>(d_a, d_b, d_o);
      fq o; cudaMemcpy(&o, d_o, sizeof(fq), cudaMemcpyDeviceToHost);
      printf("isolated fq_mul(acc[%d], z[%d]) on the device: %s\n", i - 1, i, eq(o, h.acc_after[i]) ? "equals the host" : "DIFFERS");
    }
    k_prefix_gmem<<<1, 32>>>(d_tmp, W, d_a, d_o);
    fq outs[64]; cudaMemcpy(outs, d_o, sizeof(fq) * W, cudaMemcpyDeviceToHost);
    int nb = 0; for (int i = 0; i < W; i++) nb += !eq(outs[i], h.acc_after[i]);
    printf("prefix products with the running value in global memory: %d of %d differ\n", nb, W);
  }
  // also the product kernels themselves on a 3-base table
  {
    const uint32_t nb = 3;
    std::vector<affine> bases(nb);
    xyzz t = xyzz_from_affine(G);
    for (uint32_t i = 0; i < nb; i++) { bases[i] = xyzz_to_affine(t); t = xyzz_dbl(t); xyzz_madd(t, G); }
    affine* d_b; affine *d_t1, *d_t2; xyzz* d_x; fq* d_pre;
    cudaMalloc(&d_b, sizeof(affine) * nb); cudaMalloc(&d_t1, sizeof(affine) * nb * W); cudaMalloc(&d_t2, sizeof(affine) * nb * W);
    cudaMalloc(&d_x, sizeof(xyzz) * nb * W); cudaMalloc(&d_pre, sizeof(fq) * nb * 64);
    cudaMemcpy(d_b, bases.data(), sizeof(affine) * nb, cudaMemcpyHostToDevice);
    k_table_shift<<<1, 64>>>(d_b, nb, 0, nb, c, W, d_x);
    k_table_normalise_each<<<(W * nb + 63) / 64, 64>>>(d_x, nb, 0, nb, W, d_t1);
    std::vector<affine> t1(nb * W), t2(nb * W);
    cudaMemcpy(t1.data(), d_t1, sizeof(affine) * nb * W, cudaMemcpyDeviceToHost);
    for (int mode = 0; mode < 3; mode++) {
      cudaMemset(d_t2, 0, sizeof(affine) * nb * W);
      if (mode == 0) k_table_normalise<0><<<1, 64>>>(d_x, nb, 0, nb, W, d_t2, nullptr);
      if (mode == 1) k_table_normalise<1><<<1, 64>>>(d_x, nb, 0, nb, W, d_t2, d_pre);
      if (mode == 2) k_table_normalise<2><<<1, 64>>>(d_x, nb, 0, nb, W, d_t2, nullptr);
      cudaError_t e = cudaMemcpy(t2.data(), d_t2, sizeof(affine) * nb * W, cudaMemcpyDeviceToHost);
      int diff = 0, first = -1;
      for (size_t k = 0; k < t1.size(); k++) if (memcmp(&t1[k], &t2[k], sizeof(affine))) { if (first < 0) first = (int)k; diff++; }
      printf("k_table_normalise<%d> vs _each: %d of %zu entries differ (first %d = window %d base %d) cuda=%s\n", mode, diff, t1.size(), first,
             first < 0 ? -1 : first / (int)nb, first < 0 ? -1 : first % (int)nb, cudaGetErrorString(e));
      bad += diff;
    }
  }
  printf(bad ? "FAIL %d\n" : "PASS\n", bad);
  return 0;
}
