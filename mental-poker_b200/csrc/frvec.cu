// Scalar-field vector kernels (see frvec.cuh).  All of these are small next to the MSMs: they
// are written for clarity, coalesced 128-bit access and enough parallelism to fill the chip at
// N = 2^16..2^20, not tuned further.
#include "frvec.cuh"

namespace mp {

static constexpr int kPowChunk = 8;     // consecutive powers per thread
static constexpr int kPowThreads = 128;

__device__ __forceinline__ fr fr_load(const fr* p) {
  fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4 a = s[0], b = s[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void fr_store(fr* p, const fr& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  d[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
__device__ __forceinline__ void fr_store_canonical(uint32_t* p, const fr& v) {
  uint32_t w[8];
  fr_to_canonical(v, w);
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = make_uint4(w[0], w[1], w[2], w[3]);
  d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// block-wide product / sum of one fr per thread through shared memory; result in thread 0
template <bool MUL>
__device__ fr block_combine(fr v, fr* sm) {
  const int tid = threadIdx.x;
  sm[tid] = v;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (tid < s) {
      fr a = sm[tid], b = sm[tid + s];
      sm[tid] = MUL ? fr_mul(a, b) : fr_add(a, b);
    }
    __syncthreads();
  }
  return sm[0];
}

// ------------------------------------------------------------------------------------------
// powers of x (+ optional product prod (y*i + x^i - z))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPowThreads)
    k_fr_powers(FrPow2Table tab, uint64_t N, uint32_t* __restrict__ out_canon, fr* __restrict__ out_mont,
                const fr* __restrict__ yz, bool want_prod, fr* __restrict__ partials) {
  __shared__ fr sm[kPowThreads];
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t first = t * kPowChunk + 1;  // exponent of the first power this thread writes
  fr prod = fr_one();
  if (first <= N) {
    // x^first from the table of x^(2^k)
    fr cur = fr_one();
    bool started = false;
    for (int k = 0; k < 32; k++) {
      if ((first >> k) & 1) {
        cur = started ? fr_mul(cur, tab.p[k]) : tab.p[k];
        started = true;
      }
    }
    fr y, z, yi;
    if (want_prod) {
      y = yz[0];
      z = yz[1];
      yi = fr_mul(y, fr_from_u64(first));
    }
    const fr x = tab.p[0];
    for (int j = 0; j < kPowChunk; j++) {
      uint64_t e = first + j;
      if (e > N) break;
      if (j > 0) cur = fr_mul(cur, x);
      if (out_canon) fr_store_canonical(out_canon + (e - 1) * 8, cur);
      if (out_mont) fr_store(out_mont + (e - 1), cur);
      if (want_prod) {
        if (j > 0) yi = fr_add(yi, y);
        prod = fr_mul(prod, fr_sub(fr_add(yi, cur), z));
      }
    }
  }
  if (want_prod) {
    fr total = block_combine<true>(prod, sm);
    if (threadIdx.x == 0) fr_store(partials + blockIdx.x, total);
  }
}

template <bool MUL>
__global__ void __launch_bounds__(256) k_fr_reduce_final(const fr* __restrict__ partials, unsigned count,
                                                         fr* __restrict__ out, bool negate) {
  __shared__ fr sm[256];
  fr acc = MUL ? fr_one() : fr_zero();
  for (unsigned i = threadIdx.x; i < count; i += 256) {
    fr v = fr_load(partials + i);
    acc = MUL ? fr_mul(acc, v) : fr_add(acc, v);
  }
  fr total = block_combine<MUL>(acc, sm);
  if (threadIdx.x == 0) fr_store(out, negate ? fr_neg(total) : total);
}

unsigned fr_powers_blocks(uint64_t N) {
  uint64_t threads = (N + kPowChunk - 1) / kPowChunk;
  return (unsigned)((threads + kPowThreads - 1) / kPowThreads);
}

cudaError_t fr_powers(const FrPow2Table& tab, uint64_t N, uint32_t* out_canon, fr* out_mont, const fr* yz,
                      fr* partials, fr* bstar, cudaStream_t stream) {
  if (N == 0) return cudaSuccess;
  unsigned blocks = fr_powers_blocks(N);
  k_fr_powers<<<blocks, kPowThreads, 0, stream>>>(tab, N, out_canon, out_mont, yz, bstar != nullptr, partials);
  if (bstar) k_fr_reduce_final<true><<<1, 256, 0, stream>>>(partials, blocks, bstar, false);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// element-wise / row kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fr_outer_canonical(const fr* __restrict__ coef, const fr* __restrict__ a,
                                                            int m, int n, uint32_t* __restrict__ out) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)m * n) return;
  fr v = fr_mul(fr_load(coef + g / n), fr_load(a + g % n));
  fr_store_canonical(out + g * 8, v);
}
cudaError_t fr_outer_canonical(const fr* coef, const fr* a, int m, int n, uint32_t* out_canon, cudaStream_t stream) {
  uint64_t total = (uint64_t)m * n;
  if (total == 0) return cudaSuccess;
  k_fr_outer_canonical<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(coef, a, m, n, out_canon);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fr_to_canonical(const fr* __restrict__ in, uint32_t* __restrict__ out, uint64_t n) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) fr_store_canonical(out + g * 8, fr_load(in + g));
}
__global__ void __launch_bounds__(256) k_fr_from_canonical(const uint32_t* __restrict__ in, fr* __restrict__ out, uint64_t n) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  fr raw = fr_load(reinterpret_cast<const fr*>(in) + g);
  fr_store(out + g, fr_mul(raw, fr_r2()));
}
cudaError_t fr_to_canonical_vec(const fr* in, uint32_t* out_canon, uint64_t count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  k_fr_to_canonical<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(in, out_canon, count);
  return cudaGetLastError();
}
cudaError_t fr_from_canonical_vec(const uint32_t* in_canon, fr* out, uint64_t count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  k_fr_from_canonical<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(in_canon, out, count);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fr_perm_vectors(const uint32_t* __restrict__ perm, const fr* __restrict__ xpow,
                                                         uint64_t N, fr* __restrict__ a, fr* __restrict__ b) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  uint32_t p = perm[g];
  if (a) fr_store(a + g, fr_from_u64((uint64_t)p + 1));
  if (b) fr_store(b + g, fr_load(xpow + p));
}
cudaError_t fr_perm_vectors(const uint32_t* perm, const fr* xpow, uint64_t N, fr* a, fr* b, cudaStream_t stream) {
  if (N == 0) return cudaSuccess;
  k_fr_perm_vectors<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(perm, xpow, N, a, b);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fr_affine_comb(const fr* __restrict__ a, const fr* __restrict__ b,
                                                        const fr* __restrict__ yz, uint64_t N, fr* __restrict__ d) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  fr v = fr_sub(fr_add(fr_mul(yz[0], fr_load(a + g)), fr_load(b + g)), yz[1]);
  fr_store(d + g, v);
}
cudaError_t fr_affine_comb(const fr* a, const fr* b, const fr* yz, uint64_t N, fr* d, cudaStream_t stream) {
  if (N == 0) return cudaSuccess;
  k_fr_affine_comb<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(a, b, yz, N, d);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128) k_fr_column_prefix(const fr* __restrict__ D, int m, int n, fr* __restrict__ Bv) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  fr acc = fr_load(D + j);
  fr_store(Bv + j, acc);
  for (int k = 1; k < m; k++) {
    acc = fr_mul(acc, fr_load(D + (uint64_t)k * n + j));
    fr_store(Bv + (uint64_t)k * n + j, acc);
  }
}
cudaError_t fr_column_prefix_products(const fr* D, int m, int n, fr* Bv, cudaStream_t stream) {
  if (m == 0 || n == 0) return cudaSuccess;
  k_fr_column_prefix<<<(n + 127) / 128, 128, 0, stream>>>(D, m, n, Bv);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fr_scale_rows(const fr* __restrict__ in, const fr* __restrict__ coef, int rows,
                                                       int n, fr* __restrict__ out) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)rows * n) return;
  fr_store(out + g, fr_mul(fr_load(coef + g / n), fr_load(in + g)));
}
cudaError_t fr_scale_rows(const fr* in, const fr* coef, int rows, int n, fr* out, cudaStream_t stream) {
  uint64_t total = (uint64_t)rows * n;
  if (total == 0) return cudaSuccess;
  k_fr_scale_rows<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, coef, rows, n, out);
  return cudaGetLastError();
}

// grid.x covers columns, grid.y splits the rows into slabs whose partial sums are combined by
// a second pass (count can be 2^7..2^10 rows; n a few hundred columns)
__global__ void __launch_bounds__(128) k_fr_lincomb_rows(const fr* __restrict__ rows, uint64_t stride,
                                                         const fr* __restrict__ coef, int count, int n,
                                                         fr* __restrict__ out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  fr acc = fr_zero();
  for (int k = 0; k < count; k++) acc = fr_add(acc, fr_mul(fr_load(coef + k), fr_load(rows + (uint64_t)k * stride + j)));
  fr_store(out + j, acc);
}
cudaError_t fr_lincomb_rows(const fr* rows, uint64_t stride, const fr* coef, int count, int n, fr* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_fr_lincomb_rows<<<(n + 127) / 128, 128, 0, stream>>>(rows, stride, coef, count, n, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fr_dot_partial(const fr* __restrict__ a, const fr* __restrict__ b, uint64_t N,
                                                        fr* __restrict__ partials) {
  __shared__ fr sm[256];
  fr acc = fr_zero();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x)
    acc = fr_add(acc, fr_mul(fr_load(a + i), fr_load(b + i)));
  fr total = block_combine<false>(acc, sm);
  if (threadIdx.x == 0) fr_store(partials + blockIdx.x, total);
}
unsigned fr_reduce_blocks(uint64_t N) {
  uint64_t b = (N + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 592 ? 592 : b));
}
cudaError_t fr_dot(const fr* a, const fr* b, uint64_t N, fr* partials, fr* out, cudaStream_t stream) {
  unsigned blocks = fr_reduce_blocks(N);
  k_fr_dot_partial<<<blocks, 256, 0, stream>>>(a, b, N, partials);
  k_fr_reduce_final<false><<<1, 256, 0, stream>>>(partials, blocks, out, false);
  return cudaGetLastError();
}

// pair[i*rows + j] = sum_t A[i][t] * B[j][t] * ypow[t]  -- one warp per (i, j)
__global__ void __launch_bounds__(128) k_fr_bilinear_pairs(const fr* __restrict__ A, const fr* __restrict__ B,
                                                           const fr* __restrict__ ypow, int rows, int n,
                                                           fr* __restrict__ pair) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows * rows) return;
  const int i = warp / rows, j = warp % rows;
  fr acc = fr_zero();
  for (int t = lane; t < n; t += 32) {
    fr v = fr_mul(fr_load(A + (uint64_t)i * n + t), fr_load(B + (uint64_t)j * n + t));
    acc = fr_add(acc, fr_mul(v, fr_load(ypow + t)));
  }
  for (int d = 16; d >= 1; d >>= 1) {
    fr o;
#pragma unroll
    for (int w = 0; w < 8; w++) o.v[w] = __shfl_down_sync(0xffffffffu, acc.v[w], d);
    acc = fr_add(acc, o);
  }
  if (lane == 0) fr_store(pair + warp, acc);
}
// d[k] = sum_{i - j == k - (rows-1)} pair[i][j]
__global__ void __launch_bounds__(128) k_fr_diag_sums(const fr* __restrict__ pair, int rows, fr* __restrict__ d) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 2 * rows - 1) return;
  fr acc = fr_zero();
  for (int i = 0; i < rows; i++) {
    int j = i + (rows - 1) - k;
    if (j >= 0 && j < rows) acc = fr_add(acc, fr_load(pair + (uint64_t)i * rows + j));
  }
  fr_store(d + k, acc);
}
cudaError_t fr_bilinear_diagonals(const fr* A, const fr* B, const fr* ypow, int rows, int n, fr* pair_scratch, fr* d,
                                  cudaStream_t stream) {
  uint64_t warps = (uint64_t)rows * rows;
  k_fr_bilinear_pairs<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, stream>>>(A, B, ypow, rows, n, pair_scratch);
  k_fr_diag_sums<<<(2 * rows - 1 + 127) / 128, 128, 0, stream>>>(pair_scratch, rows, d);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_fr_scatter_canonical(const fr* __restrict__ in, uint64_t count,
                                                              uint32_t* __restrict__ out, uint64_t dst_off,
                                                              uint64_t dst_stride) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < count) fr_store_canonical(out + (dst_off + g * dst_stride) * 8, fr_load(in + g));
}
cudaError_t fr_scatter_canonical(const fr* in, uint64_t count, uint32_t* out_canon, uint64_t dst_off, uint64_t dst_stride,
                                 cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  k_fr_scatter_canonical<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(in, count, out_canon, dst_off, dst_stride);
  return cudaGetLastError();
}

}  // namespace mp
