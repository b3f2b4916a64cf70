// Batched windowed-Pippenger MSM for sm_100a.  See msm.cuh for the pipeline overview and
// DESIGN.md for the roofline model of each kernel.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "msm.cuh"
#include "nvtx_ranges.hpp"

namespace mp {

// ------------------------------------------------------------------------------------------
// tunables
// ------------------------------------------------------------------------------------------
static constexpr int kChunkMin = 32;      // sorted entries per accumulate thread (small launches)
static constexpr int kChunkMax = 128;     // ... when the launch still fills the chip several times over
static constexpr int kAccThreads = 128;   // accumulate block size
static constexpr int kSegLen = 16;        // buckets per k_reduce_seg thread
static constexpr uint32_t kNoDigit = 0xffffffffu;
// record sizes of the curve selected in fq.cuh (Stark: 8-limb coordinates, BLS12-377: 12)
static constexpr int kPointWords = 2 * kFqLimbs;                 // canonical x || y
static constexpr int kAffVec = (int)(sizeof(affine) / 16);       // 128-bit words per affine point
static constexpr int kXyzzVec = (int)(sizeof(xyzz) / 16);        // ... per XYZZ accumulator
static_assert(sizeof(affine) == 8 * kFqLimbs && sizeof(xyzz) == 16 * kFqLimbs && kFqLimbs % 4 == 0, "record layout");

struct MsmWorkspace {
  // growable device buffers
  void* buf[16] = {nullptr};
  size_t cap[16] = {0};
  int launches = 0;
  // optional per-kernel timing of the bucket-accumulation kernel (roofline reporting)
  bool profile = false;
  struct Timed { cudaEvent_t e0, e1; int ncomp; uint32_t* entries; /* pinned: sorted-list length */ };
  std::vector<Timed> timed;
  uint32_t* pinned_counts = nullptr;
  static constexpr int kMaxTimed = 8192;
  ~MsmWorkspace() {
    for (int i = 0; i < 16; i++)
      if (buf[i]) cudaFree(buf[i]);
    for (auto& t : timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    if (pinned_counts) cudaFreeHost(pinned_counts);
  }
  template <typename T>
  cudaError_t get(int slot, size_t count, T** out) {
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    if (cap[slot] < bytes) {
      if (buf[slot]) cudaFree(buf[slot]);
      buf[slot] = nullptr;
      cap[slot] = 0;
      size_t want = bytes + bytes / 8;
      cudaError_t e = cudaMalloc(&buf[slot], want);
      if (e != cudaSuccess) return e;
      cap[slot] = want;
    }
    *out = reinterpret_cast<T*>(buf[slot]);
    return cudaSuccess;
  }
};

MsmWorkspace* msm_workspace_create() { return new MsmWorkspace(); }
void msm_workspace_destroy(MsmWorkspace* ws) { delete ws; }
int msm_last_launches(const MsmWorkspace* ws) { return ws->launches; }
void msm_profile_enable(MsmWorkspace* ws, bool on) { ws->profile = on; }
cudaError_t msm_profile_collect(MsmWorkspace* ws, double* ms, uint64_t* adds, uint64_t* launches, double* big_ms,
                                uint64_t* big_adds, uint64_t* big_launches) {
  *ms = 0; *adds = 0; *launches = 0;
  cudaError_t err = cudaSuccess;
  std::vector<std::pair<float, uint64_t>> all;
  for (auto& t : ws->timed) {
    cudaError_t e = cudaEventSynchronize(t.e1);
    float f = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&f, t.e0, t.e1);
    if (e != cudaSuccess) err = e;
    all.emplace_back(f, (uint64_t)(*t.entries) * t.ncomp);
    *ms += f; *adds += all.back().second; *launches += 1;
    cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
  }
  ws->timed.clear();
  // the dominant launches: those within a factor two of the largest (by additions)
  uint64_t mx = 0;
  for (auto& a : all) mx = std::max(mx, a.second);
  double bm = 0; uint64_t ba = 0, bl = 0;
  for (auto& a : all)
    if (mx && a.second * 2 >= mx) { bm += a.first; ba += a.second; bl++; }
  if (big_ms) *big_ms = bm;
  if (big_adds) *big_adds = ba;
  if (big_launches) *big_launches = bl;
  return err;
}

int msm_pick_window(uint64_t avg_len, uint64_t njobs) {
  // minimise W * (len + 2.8 * 2^(c-1)) over c, W = ceil(253 / c), plus the serial tail a SHORT top
  // window causes: with t = 253 - (W-1)c bits its digits fall into 2^(t-1) buckets only, each bucket's
  // run is cut into chunks and one k_stitch thread adds the chunk partials one after the other
  // (~10 us per addition at that occupancy, i.e. ~90 k bucket additions of chip throughput each; a
  // latency paid once per launch, so it is spread over the jobs of the launch).
  // At 2 jobs of 65 537 terms c = 13 (t = 6: 32 buckets of 2 048 entries, 0.7 ms of stitching) loses to
  // c = 11 (t = 11, 23 full windows) although the first term alone rates them equal.
  int best = 4;
  double best_cost = 1e300;
  for (int c = 4; c <= 16; c++) {
    const int W = (kScalarBits + c - 1) / c;
    const int t = kScalarBits - (W - 1) * c;
    const double top_run = (double)avg_len / (double)(1u << (t - 1));  // entries per top-window bucket
    const double stitch = top_run > 2.0 * kChunkMax ? (top_run / kChunkMax) * 90000.0 / (double)(njobs ? njobs : 1) : 0.0;
    const double cost = (double)W * ((double)avg_len + 2.8 * (double)(1u << (c - 1))) + stitch;
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

#define MP_CK(x)                          \
  do {                                    \
    cudaError_t _e = (x);                 \
    if (_e != cudaSuccess) return _e;     \
  } while (0)

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
// Out-of-line group operations for the latency-bound tail kernels (one copy of the 14-multiplication addition per
// kernel instead of one per call site).  Operands and result travel BY VALUE: the round-1 helpers took references
// (xyzz& acc, const xyzz& q), and NVVM's stack-slot merging of address-taken, loop-carried structs passed to
// __noinline__ functions produced wrong code on the 12-limb build (see fq_bls12_377.cuh fq_mul_call and
// DESIGN.md section 16) -- values have no address to merge.
__device__ __noinline__ xyzz xyzz_add_v(const xyzz acc, const xyzz q) { xyzz r = acc; xyzz_add(r, q); return r; }
__device__ __noinline__ xyzz xyzz_dbl_v(const xyzz acc) { return xyzz_dbl(acc); }
#define xyzz_add_ni(acc, q) ((acc) = xyzz_add_v((acc), (q)))
#define xyzz_dbl_ni(acc) ((acc) = xyzz_dbl_v((acc)))

__device__ __forceinline__ xyzz xyzz_load(const xyzz* p) {
  xyzz r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < kXyzzVec; i++) d[i] = s[i];
  return r;
}
__device__ __forceinline__ void xyzz_store(xyzz* p, const xyzz& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int i = 0; i < kXyzzVec; i++) d[i] = s[i];
}
__device__ __forceinline__ affine affine_load(const affine* p) {
  affine r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < kAffVec; i++) d[i] = __ldg(s + i);
  return r;
}
// ------------------------------------------------------------------------------------------
// ingest / export
// ------------------------------------------------------------------------------------------
// bad: one flag for the whole array (per_item == 0) or one flag per run of `per_item` consecutive points (the
// batch verifiers: the points of proof p occupy [p * per_item, (p + 1) * per_item))
__global__ void __launch_bounds__(128) k_points_to_mont(const uint32_t* __restrict__ in,
                                                        affine* __restrict__ out, uint64_t n,
                                                        int* __restrict__ bad, uint64_t per_item) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[kPointWords];
  const uint4* s = reinterpret_cast<const uint4*>(in + i * kPointWords);
#pragma unroll
  for (int k = 0; k < kAffVec; k++) {
    uint4 v = __ldg(s + k);
    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
  }
  affine p = affine_from_canonical(w);
  if (bad != nullptr) {
    // coordinates must be canonical (< p) -- a non-canonical alias of a curve point is rejected,
    // as ark-serialize would -- and the point must satisfy the curve equation
    fq cx, cy;
#pragma unroll
    for (int k = 0; k < kFqLimbs; k++) { cx.v[k] = w[k]; cy.v[k] = w[kFqLimbs + k]; }
    uint32_t bx, by;
    fq_sub_raw(cx, fq_kp(1), &bx);
    fq_sub_raw(cy, fq_kp(1), &by);
    if (!bx || !by || !affine_on_curve(p)) atomicExch(bad + (per_item ? i / per_item : 0), 1);
  }
  uint4* d = reinterpret_cast<uint4*>(out + i);
  const uint4* ps = reinterpret_cast<const uint4*>(&p);
#pragma unroll
  for (int k = 0; k < kAffVec; k++) d[k] = ps[k];
}

cudaError_t points_to_mont(const uint32_t* d_canonical, affine* d_out, uint64_t n, int* d_bad,
                           cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_points_to_mont<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_canonical, d_out, n, d_bad, 0);
  return cudaGetLastError();
}
cudaError_t points_to_mont_items(const uint32_t* d_canonical, affine* d_out, uint64_t n, int* d_bad_items,
                                 uint64_t per_item, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_points_to_mont<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_canonical, d_out, n, d_bad_items, per_item);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(64) k_xyzz_to_canonical(const xyzz* __restrict__ in,
                                                          uint32_t* __restrict__ out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  xyzz p = xyzz_load(in + i);
  affine a = xyzz_to_affine(p);
  uint32_t w[kPointWords];
  if (affine_is_identity(a)) {
#pragma unroll
    for (int k = 0; k < kPointWords; k++) w[k] = 0;
  } else {
    affine_to_canonical(a, w);
  }
  uint4* d = reinterpret_cast<uint4*>(out + i * kPointWords);
#pragma unroll
  for (int k = 0; k < kAffVec; k++) d[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

cudaError_t xyzz_to_canonical(const xyzz* d_in, uint32_t* d_out, uint64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_xyzz_to_canonical<<<(unsigned)((n + 63) / 64), 64, 0, stream>>>(d_in, d_out, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// fixed-base tables (kernel family K3 of SURVEY.md 2b): table[w * nb + i] = 2^(c*w) * base_i
// ------------------------------------------------------------------------------------------
// thread per base: the chain of c*(W-1) doublings, every window's value kept in XYZZ
__global__ void __launch_bounds__(64) k_table_shift(const affine* __restrict__ bases, uint32_t nb, uint32_t first,
                                                    uint32_t count, int c, int W, xyzz* __restrict__ tmp) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  affine P = affine_load(bases + first + i);
  xyzz cur = xyzz_from_affine(P);
  for (int w = 0; w < W; w++) {
    xyzz_store(tmp + (size_t)w * count + i, cur);
    if (w + 1 < W)
      for (int k = 0; k < c; k++) cur = xyzz_dbl(cur);
  }
}
// thread per base: normalise its W shifted copies to affine Montgomery with ONE inversion
// (Montgomery's trick over the W values of ZZ*ZZZ); the identity stays (0,0)
static constexpr int kMaxTableWindows = 64;
// MODE (experiment knob, scripts/sanitize.sh): 0 = per-thread prefix array in local memory, multiplication as
// compiled for this curve (a call on the 12-limb build); 1 = prefix array in global scratch; 2 = local array,
// multiplications of this kernel inlined.
#ifdef MP_CURVE_BLS12_377
#define MP_TN_MUL(MODE, a, b) ((MODE) == 2 ? fq_mul_inline(a, b) : fq_mul(a, b))
#else
#define MP_TN_MUL(MODE, a, b) fq_mul(a, b)
#endif
template <int MODE>
__global__ void __launch_bounds__(64) k_table_normalise(const xyzz* __restrict__ tmp, uint32_t nb, uint32_t first,
                                                        uint32_t count, int W, affine* __restrict__ table,
                                                        fq* __restrict__ pre_all) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  fq pre_local[MODE == 1 ? 1 : kMaxTableWindows];  // pre[w] = product of z_0 .. z_w (z = ZZ*ZZZ, or 1 for the identity)
  fq* pre = MODE == 1 ? pre_all + (size_t)i * kMaxTableWindows : pre_local;
  fq acc = fq_one();
  for (int w = 0; w < W; w++) {
    xyzz p = xyzz_load(tmp + (size_t)w * count + i);
    if (!xyzz_is_identity(p)) acc = MP_TN_MUL(MODE, acc, MP_TN_MUL(MODE, p.ZZ, p.ZZZ));
    pre[w] = acc;
  }
  fq inv = fq_inv(acc);  // 1 / (z_0 * ... * z_{W-1})
  for (int w = W - 1; w >= 0; w--) {
    xyzz p = xyzz_load(tmp + (size_t)w * count + i);
    affine a;
    if (xyzz_is_identity(p)) {
      a.x = fq_zero();
      a.y = fq_zero();
    } else {
      fq z = MP_TN_MUL(MODE, p.ZZ, p.ZZZ);
      fq iz = w > 0 ? MP_TN_MUL(MODE, inv, pre[w - 1]) : inv;  // 1 / z_w
      inv = MP_TN_MUL(MODE, inv, z);                            // drop z_w from the running inverse
      a.x = fq_reduce_full(MP_TN_MUL(MODE, p.X, MP_TN_MUL(MODE, iz, p.ZZZ)));
      a.y = fq_reduce_full(MP_TN_MUL(MODE, p.Y, MP_TN_MUL(MODE, iz, p.ZZ)));
    }
    uint4* d = reinterpret_cast<uint4*>(table + (size_t)w * nb + first + i);
    const uint4* s = reinterpret_cast<const uint4*>(&a);
#pragma unroll
    for (int k = 0; k < kAffVec; k++) d[k] = s[k];
  }
}

// thread per (window, base): one inversion per table entry.  W x more inversions than the kernel above
// (a table is built once per key: 513 bases x 43 windows x ~470 multiplications is ~1 ms), no per-thread
// prefix array
__global__ void __launch_bounds__(64) k_table_normalise_each(const xyzz* __restrict__ tmp, uint32_t nb, uint32_t first,
                                                             uint32_t count, int W, affine* __restrict__ table) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)W * count) return;
  const uint32_t w = (uint32_t)(g / count), i = (uint32_t)(g % count);
  const affine a = xyzz_to_affine(xyzz_load(tmp + g));
  uint4* d = reinterpret_cast<uint4*>(table + (size_t)w * nb + first + i);
  const uint4* s = reinterpret_cast<const uint4*>(&a);
#pragma unroll
  for (int k = 0; k < kAffVec; k++) d[k] = s[k];
}

cudaError_t msm_build_table(MsmWorkspace* ws, const affine* d_bases, uint32_t nb, uint32_t first, uint32_t count,
                            int c, affine* d_table, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const int W = (kScalarBits + c - 1) / c;
  if (W > kMaxTableWindows) return cudaErrorInvalidValue;
  xyzz* tmp;
  MP_CK(ws->get(13, (size_t)W * count, &tmp));
  k_table_shift<<<(count + 63) / 64, 64, 0, stream>>>(d_bases, nb, first, count, c, W, tmp);
  static const int trick = [] { const char* e = getenv("MP_TABLE_TRICK"); return e ? atoi(e) : (kFqLimbs <= 8 ? 1 : 0); }();
  if (trick == 2) {
    fq* pre;
    MP_CK(ws->get(15, (size_t)count * kMaxTableWindows, &pre));
    k_table_normalise<1><<<(count + 63) / 64, 64, 0, stream>>>(tmp, nb, first, count, W, d_table, pre);
  } else if (trick == 3)
    k_table_normalise<2><<<(count + 63) / 64, 64, 0, stream>>>(tmp, nb, first, count, W, d_table, nullptr);
  else if (trick)
    k_table_normalise<0><<<(count + 63) / 64, 64, 0, stream>>>(tmp, nb, first, count, W, d_table, nullptr);
  else
    k_table_normalise_each<<<(unsigned)(((uint64_t)W * count + 63) / 64), 64, 0, stream>>>(tmp, nb, first, count, W, d_table);
  return cudaGetLastError();
}

int msm_pick_table_window(uint64_t typical_len) {
  // one bucket set per job: minimise W * len + 2.8 * 2^(c-1)
  int best = 4;
  double best_cost = 1e300;
  for (int c = 4; c <= 16; c++) {
    int W = (kScalarBits + c - 1) / c;
    double cost = (double)W * (double)typical_len + 2.8 * (double)(1u << (c - 1));
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

// ------------------------------------------------------------------------------------------
// digits
// ------------------------------------------------------------------------------------------
// digits[w * ns + i] = (|d| - 1) | (d < 0) << 31, or kNoDigit when d == 0.
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars,
                                                uint32_t* __restrict__ digits, uint64_t ns, int c,
                                                int w_begin, int W) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  uint32_t s[9];
  const uint4* p = reinterpret_cast<const uint4*>(scalars + i * 8);
  uint4 lo = __ldg(p), hi = __ldg(p + 1);
  s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w;
  s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
  s[8] = 0;
  const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
  uint32_t carry = 0;
  for (int w = 0; w < w_begin + W; w++) {  // windows below w_begin only feed the carry
    int pos = w * c;
    uint32_t raw = 0;
    if (pos < 256) {
      int word = pos >> 5, sh = pos & 31;
      uint64_t two = (uint64_t)s[word] | ((uint64_t)s[word + 1] << 32);
      raw = (uint32_t)(two >> sh) & mask;
    }
    raw += carry;
    uint32_t enc;
    if (raw > half) {  // negative digit raw - 2^c, magnitude 2^c - raw in [1, half-1]
      enc = ((mask + 1u - raw) - 1u) | 0x80000000u;
      carry = 1;
    } else {
      enc = raw == 0 ? kNoDigit : (raw - 1u);
      carry = 0;
    }
    if (w >= w_begin) digits[(uint64_t)(w - w_begin) * ns + i] = enc;
  }
}

// grid.y = job; histogram of buckets
__global__ void __launch_bounds__(256) k_count(const uint32_t* __restrict__ digits,
                                               const MsmJob* __restrict__ jobs,
                                               uint32_t* __restrict__ counts, uint64_t ns, int W,
                                               uint32_t B, int Wb) {
  const MsmJob job = jobs[blockIdx.y];
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < job.len;
       t += gridDim.x * blockDim.x) {
    uint64_t si = (uint64_t)job.scalar_off + t;
    for (int w = 0; w < W; w++) {
      uint32_t d = digits[(uint64_t)w * ns + si];
      if (d != kNoDigit) {
        // Wb == W: one bucket set per window; Wb == 1 (fixed-base tables): all windows share one
        uint64_t bucket = ((uint64_t)blockIdx.y * Wb + (Wb == 1 ? 0 : w)) * B + (d & 0x7fffffffu);
        atomicAdd(&counts[bucket], 1u);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_scatter(const uint32_t* __restrict__ digits,
                                                 const MsmJob* __restrict__ jobs,
                                                 uint32_t* __restrict__ cursor,
                                                 uint32_t* __restrict__ sorted, uint64_t ns, int W,
                                                 uint32_t B, int Wb, uint32_t tab_nb) {
  const MsmJob job = jobs[blockIdx.y];
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < job.len;
       t += gridDim.x * blockDim.x) {
    uint64_t si = (uint64_t)job.scalar_off + t;
    for (int w = 0; w < W; w++) {
      uint32_t d = digits[(uint64_t)w * ns + si];
      if (d != kNoDigit) {
        uint64_t bucket = ((uint64_t)blockIdx.y * Wb + (Wb == 1 ? 0 : w)) * B + (d & 0x7fffffffu);
        uint32_t pos = atomicAdd(&cursor[bucket], 1u);
        // table mode: the point is the precomputed 2^(c*w) * base, stored window-major
        sorted[pos] = ((uint32_t)w * tab_nb + job.point_off + t) | (d & 0x80000000u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// exclusive scan of bucket sizes: tiles of 2048 (256 threads x 8)
// ------------------------------------------------------------------------------------------
static constexpr int kScanTile = 2048;

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[8];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    uint32_t s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(256) k_scan_tile_sums(const uint32_t* __restrict__ counts,
                                                        uint32_t* __restrict__ tile_sums,
                                                        uint64_t n) {
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + threadIdx.x * 8;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (base + k < n) s += counts[base + k];
  uint32_t tot;
  block_exclusive_scan_256(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of tile_sums[ntiles]; writes grand total to *total_out
__global__ void __launch_bounds__(256) k_scan_top(uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                  uint32_t* __restrict__ total_out) {
  uint32_t running = 0;
  for (uint32_t base = 0; base < ntiles; base += 256) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan_256(v, &tot);
    if (i < ntiles) tile_sums[i] = running + ex;
    running += tot;
  }
  if (threadIdx.x == 0) *total_out = running;
}

__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t* __restrict__ counts,
                                                    const uint32_t* __restrict__ tile_sums,
                                                    uint32_t* __restrict__ offsets,
                                                    uint32_t* __restrict__ cursor, uint64_t n) {
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + threadIdx.x * 8;
  uint32_t v[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[k] = (base + k < n) ? counts[base + k] : 0;
    s += v[k];
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan_256(s, &tot) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (base + k < n) {
      offsets[base + k] = ex;
      cursor[base + k] = ex;
    }
    ex += v[k];
  }
}

// ------------------------------------------------------------------------------------------
// bucket accumulation (the hot kernel)
// ------------------------------------------------------------------------------------------
// Thread g handles component (g % ncomp) of chunk t = g / ncomp: sorted entries
// [t*kChunk, min((t+1)*kChunk, E)).  Its first bucket run goes to part[g]; every later run
// (which starts inside the chunk) goes to bucket_sums[bucket * ncomp + comp].
// Every warp owns one tile of 32 / NCOMP consecutive chunks.  The tile's slice of the sorted index
// list is one contiguous run, so the warp stages it into shared memory with a single TMA bulk copy
// (cp.async.bulk, completion signalled on the warp's mbarrier) instead of per-thread strided
// loads; the 64-byte points themselves are a data-dependent gather and are fetched with 128-bit
// loads, software-prefetched one entry ahead.  The grid is NOT persistent on purpose: with
// identical code a chip-sized persistent grid (static tile -> block assignment) measured 8.08 G
// additions/s against 8.54 G/s for one tile per warp under the hardware block scheduler.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCOMP, int MINBLOCKS>
__global__ void __launch_bounds__(kAccThreads, MINBLOCKS)
    k_accumulate(const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ offsets,
                 uint64_t nbuckets, const affine* __restrict__ points, xyzz* __restrict__ bucket_sums,
                 xyzz* __restrict__ part, uint32_t* __restrict__ chunk_bucket, uint32_t kChunk) {
  extern __shared__ __align__(128) uint32_t s_idx[];  // [warp][tile] sorted indices
  constexpr uint32_t kWarps = kAccThreads / 32;
  __shared__ __align__(8) uint64_t s_bar[kWarps];
  constexpr uint32_t kChunksPerTile = 32 / NCOMP;
  const uint32_t tile_entries = kChunksPerTile * kChunk, tile_bytes = tile_entries * 4u;
  const uint64_t E = offsets[nbuckets];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t tile = (uint64_t)blockIdx.x * kWarps + warp;
  if (tile * tile_entries >= E) return;  // whole warp: nothing sorted into this tile
  const uint32_t bar = smem_u32(&s_bar[warp]);
  uint32_t* const my_tile = s_idx + (size_t)warp * tile_entries;
  if (lane == 0) {
    // the elected lane arms the mbarrier with the byte count and issues the bulk copy
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(my_tile)),
                 "l"(sorted + tile * tile_entries), "r"(tile_bytes), "r"(bar)
                 : "memory");
  }
  __syncwarp();
  const uint32_t lc = lane / NCOMP;
  const uint32_t comp = lane % NCOMP;
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar)
          : "memory");
    }
  }
  {
    const uint64_t t = tile * kChunksPerTile + lc;
    const uint64_t g = t * NCOMP + comp;
    uint64_t pos = t * kChunk;
    if (pos < E) {
      const uint32_t* my = my_tile + lc * kChunk;
      const uint64_t first_pos = pos;
      const uint64_t end = min(pos + (uint64_t)kChunk, E);
      // bucket containing entry `pos`: largest b with offsets[b] <= pos
      uint64_t lo = 0, hi = nbuckets;  // invariant: offsets[lo] <= pos < offsets[hi]
      while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= pos) lo = mid; else hi = mid;
      }
      uint64_t b = lo;
      if (comp == 0) chunk_bucket[t] = (uint32_t)b;
      uint64_t next = offsets[b + 1];
      bool first = true;
      xyzz acc = xyzz_identity();
      // indices come out of shared memory 4 at a time: every lane's chunk starts at the same
      // bank, so 128-bit reads cut the (inevitable, the tile is copied linearly) bank conflicts 4x
      const uint4* my4 = reinterpret_cast<const uint4*>(my);
      uint4 cur4 = my4[0];
      uint32_t val = cur4.x;
      affine pt = affine_load(points + (uint64_t)(val & 0x7fffffffu) * NCOMP + comp);
      while (true) {
        // prefetch the next entry's point while this one is being added
        uint32_t nval = val;
        affine npt = pt;
        if (pos + 1 < end) {
          const uint32_t j = (uint32_t)(pos + 1 - first_pos);
          if ((j & 3u) == 0) cur4 = my4[j >> 2];
          nval = (j & 3u) == 0 ? cur4.x : ((j & 3u) == 1 ? cur4.y : ((j & 3u) == 2 ? cur4.z : cur4.w));
          npt = affine_load(points + (uint64_t)(nval & 0x7fffffffu) * NCOMP + comp);
        }
        if (pos == next) {  // entering a new bucket: flush the finished run
          if (first) xyzz_store(part + g, acc);
          else xyzz_store(bucket_sums + b * NCOMP + comp, acc);
          first = false;
          acc = xyzz_identity();
          b++;
          next = offsets[b + 1];
          if (next <= pos) {
            // a run of empty buckets (sparse jobs: a 2-term commitment still owns a whole bucket set):
            // gallop, then bisect, instead of one dependent load per empty bucket.
            // invariant: offsets[lo2 + 1] <= pos < offsets[hi2 + 1]   (offsets[nbuckets] = E > pos)
            uint64_t lo2 = b, hi2 = b, step = 1;
            for (;;) {
              hi2 = min(lo2 + step, nbuckets - 1);
              if (offsets[hi2 + 1] > pos) break;
              lo2 = hi2;
              step <<= 1;
            }
            while (hi2 - lo2 > 1) {
              const uint64_t mid = (lo2 + hi2) >> 1;
              if (offsets[mid + 1] <= pos) lo2 = mid; else hi2 = mid;
            }
            b = hi2;
            next = offsets[b + 1];
          }
        }
        if (val >> 31) pt = affine_neg(pt);
        xyzz_madd(acc, pt);
        pos++;
        if (pos >= end) break;
        val = nval;
        pt = npt;
      }
      if (first) xyzz_store(part + g, acc);
      else xyzz_store(bucket_sums + b * NCOMP + comp, acc);
    }
  }
}

// Stitch: thread per (chunk, comp).  A bucket whose entries straddle chunk boundaries has its sum
// split between bucket_sums[b] (the run that started inside a chunk) and the `part` of every
// chunk whose first entry lies in the bucket.  The first such chunk ("leader") folds the parts
// into bucket_sums[b]; only buckets that contain a chunk boundary are touched at all.
__global__ void __launch_bounds__(128, 4) k_stitch(const uint32_t* __restrict__ offsets, uint64_t nbuckets,
                                                const uint32_t* __restrict__ chunk_bucket, int ncomp,
                                                xyzz* __restrict__ bucket_sums, const xyzz* __restrict__ part,
                                                uint32_t kChunk) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t t = g / ncomp;
  const uint32_t comp = (uint32_t)(g % ncomp);
  const uint64_t E = offsets[nbuckets];
  const uint64_t nchunks = (E + kChunk - 1) / kChunk;
  if (t >= nchunks) return;
  const uint32_t b = chunk_bucket[t];
  if (t > 0 && chunk_bucket[t - 1] == b) return;  // not the first chunk starting in this bucket
  xyzz acc = xyzz_identity();
  if (offsets[b] % kChunk != 0) acc = xyzz_load(bucket_sums + (uint64_t)b * ncomp + comp);
  for (uint64_t u = t; u < nchunks && chunk_bucket[u] == b; u++) {
    xyzz p = xyzz_load(part + u * ncomp + comp);
    xyzz_add(acc, p);  // inlined: nearly every thread of a single large MSM adds exactly one partial
  }
  xyzz_store(bucket_sums + (uint64_t)b * ncomp + comp, acc);
}

// Thread per (window, segment, comp): running-sum sweep over the segment's L buckets from the top,
//   S = sum of the segment's buckets,   T = sum_{i=0..L-1} (i+1) * B_i.
// Empty buckets are recognised from the offsets (bucket_sums is never zero-filled).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_reduce_seg(const uint32_t* __restrict__ offsets,
                                                    const xyzz* __restrict__ bucket_sums, uint64_t nwin,
                                                    uint32_t B, uint32_t L, int ncomp,
                                                    xyzz* __restrict__ segS, xyzz* __restrict__ segT) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nseg = B / L;
  if (g >= nwin * nseg * ncomp) return;
  uint32_t comp = (uint32_t)(g % ncomp);
  uint64_t ws = g / ncomp;  // window * nseg + seg
  uint64_t first_bucket = ws * L;
  xyzz running = xyzz_identity(), acc = xyzz_identity();
  uint32_t o1 = offsets[first_bucket + L];
  for (int i = (int)L - 1; i >= 0; i--) {
    const uint64_t b = first_bucket + i;
    const uint32_t o0 = offsets[b];
    if (o1 > o0) {
      xyzz bkt = xyzz_load(bucket_sums + b * ncomp + comp);
      xyzz_add(running, bkt);
    }
    xyzz_add(acc, running);
    o1 = o0;
  }
  xyzz_store(segS + g, running);
  xyzz_store(segT + g, acc);
}

// Hierarchical form of the same combine: thread per (window, group of G consecutive segments, comp) folds
// its group into ONE segment of length L*G (G a power of two),
//   S' = sum_i S_i,   T' = sum_i T_i + L * sum_i i * S_i     (i = index inside the group),
// because a bucket at offset j of sub-segment i has weight i*L + (j+1) in the merged segment.  Applying it
// until one segment per window is left yields T' = the window sum.  One thread per group: the form for launches
// with many windows (batches of small jobs); k_reduce_group_quad is the same recurrence for few.
__global__ void __launch_bounds__(64) k_reduce_group(const xyzz* __restrict__ segS, const xyzz* __restrict__ segT,
                                                     uint64_t nwin, uint32_t nseg, uint32_t G, uint32_t L, int ncomp,
                                                     xyzz* __restrict__ outS, xyzz* __restrict__ outT) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ngroups = nseg / G;
  if (g >= nwin * ngroups * ncomp) return;
  const uint32_t comp = (uint32_t)(g % ncomp);
  const uint64_t wg = g / ncomp;  // window * ngroups + group
  const uint64_t win = wg / ngroups, grp = wg % ngroups;
  const xyzz* S = segS + (win * nseg + grp * G) * ncomp + comp;
  const xyzz* T = segT + (win * nseg + grp * G) * ncomp + comp;
  xyzz run = xyzz_identity(), lsum = xyzz_identity(), tsum = xyzz_load(T);
  for (uint32_t i = G - 1; i >= 1; i--) {
    xyzz a = xyzz_load(S + (uint64_t)i * ncomp);
    xyzz_add_ni(run, a);
    xyzz_add_ni(lsum, run);
    xyzz tv = xyzz_load(T + (uint64_t)i * ncomp);
    xyzz_add_ni(tsum, tv);
  }
  xyzz s0 = xyzz_load(S);
  xyzz_add_ni(run, s0);  // S' includes sub-segment 0 (weight 0 in lsum)
  for (uint32_t k = 1; k < L; k <<= 1) xyzz_dbl_ni(lsum);
  xyzz_add_ni(tsum, lsum);
  xyzz_store(outS + g, run);
  xyzz_store(outT + g, tsum);
}

// ------------------------------------------------------------------------------------------
// Horner fold over a job's W window sums: the one inherently serial piece of a (non-table) MSM -- c*(W-1) ~ 240
// dependent doublings.  A doubling is ten field multiplications of which only three are on its critical path, so
// a QUAD of lanes owns one (job, comp) accumulator -- lane 0 holds X, lane 1 Y, lane 2 ZZ, lane 3 ZZZ -- and runs
// dbl-2008-s-1 in three multiplication levels, every level ONE uniform fq_mul / fq_sqr whose operands each lane
// picks for its role, results exchanged with quad-masked shuffles:
//   level 1 (squares)   XX = X^2          | V = (2Y)^2       | ZZ2 = ZZ^2 (a = 1)  | --
//   level 2             S = X V           | W = U V          | ZZ3 = V ZZ          | MM = M^2,  M = 3 XX (+ ZZ2)
//   level 3             t = M (S - X3)    | WY = W Y         | --                  | ZZZ3 = W ZZZ     X3 = MM - 2S
//   then                X3                | Y3 = t - WY      | ZZ3                 | ZZZ3
// ~2 000 cycles per doubling instead of ~6 000 for one thread.  The W - 1 additions of the window sums gather the
// accumulator into every lane of the quad and run the complete xyzz_add redundantly (all special cases kept).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ fq fq_shfl(uint32_t mask, const fq& v, int src) {
  fq r;
#pragma unroll
  for (int i = 0; i < kFqLimbs; i++) r.v[i] = __shfl_sync(mask, v.v[i], src);
  return r;
}
__device__ __forceinline__ fq fq_select(bool c, const fq& a, const fq& b) {
  fq r;
#pragma unroll
  for (int i = 0; i < kFqLimbs; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}
// one doubling of the quad's accumulator (not the identity); `home` = this lane's coordinate
__device__ __forceinline__ void quad_dbl(fq& home, int role, int base, uint32_t mask) {
  const fq U = fq_add(home, home);                                  // lane 1: 2Y  [4]
  const fq s1 = fq_sqr(role == 1 ? U : home);                       // XX | V | ZZ2 | (unused)
  const fq XX = fq_shfl(mask, s1, base), V = fq_shfl(mask, s1, base + 1);
#if MP_CURVE_A_IS_ZERO
  const fq M = fq_reduce_weak(fq_add(fq_add(XX, XX), XX));
#else
  const fq ZZ2 = fq_shfl(mask, s1, base + 2);
  const fq M = fq_reduce_weak(fq_add(fq_add(XX, XX), fq_add(XX, ZZ2)));
#endif
  const fq a2 = fq_select(role == 3, M, role == 1 ? U : home);
  const fq s2 = fq_mul(a2, fq_select(role == 3, M, V));             // S | W | ZZ3 | MM
  const fq S = fq_shfl(mask, s2, base), Wv = fq_shfl(mask, s2, base + 1), MM = fq_shfl(mask, s2, base + 3);
  const fq X3 = fq_reduce_weak(fq_sub(MM, fq_add(S, S), 4));
  const fq s3 = fq_mul(fq_select(role == 0, M, Wv), fq_select(role == 0, fq_sub(S, X3, 2), home));  // t | WY | -- | ZZZ3
  const fq t = fq_shfl(mask, s3, base);
  // Y == 0 (mod p) would make ZZ3 == 0 (mod p): the result is the identity, stored as exact zero words
  const bool dead = __shfl_sync(mask, (int)fq_is_zero_mod_p_2(s2), base + 2) != 0;
  fq out = role == 0 ? X3 : (role == 1 ? fq_reduce_weak(fq_sub(t, s3, 2)) : (role == 2 ? s2 : s3));
  home = dead ? fq_zero() : out;
}
__device__ __forceinline__ xyzz quad_gather(const fq& home, int base, uint32_t mask) {
  xyzz r;
  r.X = fq_shfl(mask, home, base); r.Y = fq_shfl(mask, home, base + 1);
  r.ZZ = fq_shfl(mask, home, base + 2); r.ZZZ = fq_shfl(mask, home, base + 3);
  return r;
}
__device__ __forceinline__ fq quad_home(const xyzz& p, int role) {
  return role == 0 ? p.X : (role == 1 ? p.Y : (role == 2 ? p.ZZ : p.ZZZ));
}

// acc += q for the quad's accumulator; q is fully present in every lane of the quad.  add-2008-s in FOUR
// multiplication levels (a thread needs fourteen multiplications in a row):
//   level 1   U1 = X1 q.ZZ      | S1 = Y1 q.ZZZ        | U2 = q.X ZZ1         | S2 = q.Y ZZZ1         P = U2-U1, R = S2-S1
//   level 2   PP = P^2          | RR = R^2             | ZZ12 = ZZ1 q.ZZ      | ZZZ12 = ZZZ1 q.ZZZ
//   level 3   Q = U1 PP         | PPP = P PP           | ZZ3 = ZZ12 PP        | (PPP)                 X3 = RR - PPP - 2Q
//   level 4   --                | t1 = R (Q - X3)      | t2 = S1 PPP          | ZZZ3 = ZZZ12 PPP      Y3 = t1 - t2
// Identities are handled before level 1; equal x-coordinates (doubling / cancellation, seen at level 3) fall back
// to the complete single-thread formula on the gathered accumulator.
__device__ __forceinline__ void quad_add(fq& home, const xyzz& q, int role, int base, uint32_t mask) {
  if (xyzz_is_identity(q)) return;                                   // quad-uniform: q is the same in all four lanes
  if (fq_is_zero_raw(fq_shfl(mask, home, base + 2))) { home = quad_home(q, role); return; }
  const fq s1 = fq_mul(home, role == 0 ? q.ZZ : (role == 1 ? q.ZZZ : (role == 2 ? q.X : q.Y)));
  const fq U1 = fq_shfl(mask, s1, base), S1 = fq_shfl(mask, s1, base + 1);
  const fq P = fq_sub(fq_shfl(mask, s1, base + 2), U1, 2), R = fq_sub(fq_shfl(mask, s1, base + 3), S1, 2);  // [4]
  const fq a2 = role == 0 ? P : (role == 1 ? R : home);
  const fq s2 = fq_mul(a2, role == 0 ? P : (role == 1 ? R : (role == 2 ? q.ZZ : q.ZZZ)));
  const fq PP = fq_shfl(mask, s2, base), RR = fq_shfl(mask, s2, base + 1);
  const fq s3 = fq_mul(role == 0 ? U1 : (role == 2 ? s2 : P), PP);
  if (__shfl_sync(mask, (int)fq_is_zero_mod_p_2(s3), base + 2)) {    // P == 0 (mod p): same x
    xyzz acc = quad_gather(home, base, mask);
    acc = xyzz_add_v(acc, q);
    home = quad_home(acc, role);
    return;
  }
  const fq Q = fq_shfl(mask, s3, base), PPP = fq_shfl(mask, s3, base + 1);
  const fq X3 = fq_reduce_weak(fq_sub(RR, fq_add(PPP, fq_add(Q, Q)), 6));
  const fq s4 = fq_mul(role == 1 ? R : (role == 2 ? S1 : s2), role == 1 ? fq_sub(Q, X3, 2) : PPP);
  const fq t2 = fq_shfl(mask, s4, base + 2);
  home = role == 0 ? X3 : (role == 1 ? fq_reduce_weak(fq_sub(s4, t2, 2)) : (role == 2 ? s3 : s4));
}

// Quad form of k_reduce_group for launches too small to fill the chip with one thread per group (a single large
// MSM has 16 windows): same recurrence, every group operation cooperative.
__global__ void __launch_bounds__(128) k_reduce_group_quad(const xyzz* __restrict__ segS, const xyzz* __restrict__ segT,
                                                           uint64_t nwin, uint32_t nseg, uint32_t G, uint32_t L, int ncomp,
                                                           xyzz* __restrict__ outS, xyzz* __restrict__ outT) {
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  const uint32_t mask = 0xfu << base;
  const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const uint32_t ngroups = nseg / G;
  if (g >= nwin * ngroups * ncomp) return;  // whole quads leave together
  const uint32_t comp = (uint32_t)(g % ncomp);
  const uint64_t wg = g / ncomp;
  const uint64_t win = wg / ngroups, grp = wg % ngroups;
  const xyzz* S = segS + (win * nseg + grp * G) * ncomp + comp;
  const xyzz* T = segT + (win * nseg + grp * G) * ncomp + comp;
  fq run = fq_zero(), lsum = fq_zero(), tsum = quad_home(xyzz_load(T), role);
#pragma unroll 1
  for (uint32_t i = G - 1; i >= 1; i--) {
    quad_add(run, xyzz_load(S + (uint64_t)i * ncomp), role, base, mask);
    quad_add(lsum, quad_gather(run, base, mask), role, base, mask);
    quad_add(tsum, xyzz_load(T + (uint64_t)i * ncomp), role, base, mask);
  }
  quad_add(run, xyzz_load(S), role, base, mask);  // S' includes sub-segment 0 (weight 0 in lsum)
  if (!fq_is_zero_raw(fq_shfl(mask, lsum, base + 2))) {
#pragma unroll 1
    for (uint32_t k = 1; k < L; k <<= 1) quad_dbl(lsum, role, base, mask);
  }
  quad_add(tsum, quad_gather(lsum, base, mask), role, base, mask);
  const xyzz rS = quad_gather(run, base, mask), rT = quad_gather(tsum, base, mask);
  if (role == 0) { xyzz_store(outS + g, rS); xyzz_store(outT + g, rT); }
}

// Quad per (job, comp); a block of 32 threads folds 8 of them.
__global__ void __launch_bounds__(32) k_fold(const xyzz* __restrict__ win_out, int njobs, int W,
                                             int c, int ncomp, xyzz* __restrict__ out) {
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  const uint32_t mask = 0xfu << base;
  const int g = blockIdx.x * 8 + (lane >> 2);
  if (g >= njobs * ncomp) return;  // whole quads leave together
  const int job = g / ncomp, comp = g % ncomp;
  fq home = quad_home(xyzz_load(win_out + ((uint64_t)(job * W + W - 1)) * ncomp + comp), role);
#pragma unroll 1
  for (int w = W - 2; w >= 0; w--) {
    const bool is_id = fq_is_zero_raw(fq_shfl(mask, home, base + 2));
    if (!is_id) {
#pragma unroll 1
      for (int k = 0; k < c; k++) quad_dbl(home, role, base, mask);
    }
    xyzz acc = quad_gather(home, base, mask);
    const xyzz v = xyzz_load(win_out + ((uint64_t)(job * W + w)) * ncomp + comp);
    acc = xyzz_add_v(acc, v);
    home = quad_home(acc, role);
  }
  const xyzz res = quad_gather(home, base, mask);
  if (role == 0) xyzz_store(out + g, res);
}

// Fold of the per-rank partials of a window-range split (comm.cu):  out = sum_r 2^(c * w_begin_r) * P_r, evaluated as
// a Horner chain from the highest rank down -- shifts.s[r] = doublings between rank r + 1's partial and rank r's.
struct FoldShifts { int s[64]; };
__global__ void __launch_bounds__(32) k_fold_ranges(const xyzz* __restrict__ parts, int nranks, int ncomp, FoldShifts shifts,
                                                    xyzz* __restrict__ out) {
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  const uint32_t mask = 0xfu << base;
  const int comp = blockIdx.x * 8 + (lane >> 2);
  if (comp >= ncomp) return;
  fq home = quad_home(xyzz_load(parts + (size_t)(nranks - 1) * ncomp + comp), role);
#pragma unroll 1
  for (int r = nranks - 2; r >= 0; r--) {
    if (!fq_is_zero_raw(fq_shfl(mask, home, base + 2))) {
#pragma unroll 1
      for (int k = 0; k < shifts.s[r]; k++) quad_dbl(home, role, base, mask);
    }
    quad_add(home, xyzz_load(parts + (size_t)r * ncomp + comp), role, base, mask);
  }
  const xyzz res = quad_gather(home, base, mask);
  if (role == 0) xyzz_store(out + comp, res);
}
cudaError_t msm_fold_ranges(const xyzz* d_parts, int nranks, int ncomp, const int* shifts, xyzz* d_out, cudaStream_t stream) {
  if (nranks < 1 || nranks > 64) return cudaErrorInvalidValue;
  FoldShifts fs;
  for (int r = 0; r < 64; r++) fs.s[r] = r < nranks ? shifts[r] : 0;
  k_fold_ranges<<<(unsigned)((ncomp + 7) / 8), 32, 0, stream>>>(d_parts, nranks, ncomp, fs, d_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------

cudaError_t msm_run(MsmWorkspace* ws, const uint32_t* d_scalars, uint64_t n_scalars,
                    const affine* d_points, int ncomp, const MsmJob* h_jobs, int njobs, int c,
                    xyzz* d_out, cudaStream_t stream, int w_begin, int w_count, uint32_t table_nb) {
  NvtxRange nvtx(table_nb ? "msm_run (fixed-base table)" : "msm_run");
  ws->launches = 0;
  if (njobs <= 0) return cudaSuccess;
  if (ncomp != 1 && ncomp != 2) return cudaErrorInvalidValue;
  if (c < 2 || c > 20) return cudaErrorInvalidValue;
  const int W_all = (kScalarBits + c - 1) / c;
  if (w_count < 0) w_count = W_all - w_begin;
  if (w_begin < 0 || w_count < 1 || w_begin + w_count > W_all) return cudaErrorInvalidValue;
  const int W = w_count;  // windows handled by this call: [w_begin, w_begin + W)
  // fixed-base table mode (table_nb = bases per window of a table built by msm_build_table with
  // the same c): every window's digits index ONE bucket set per job and the entries point at the
  // pre-shifted bases, so there is one bucket reduction per job and no fold doublings at all
  if (table_nb && (w_begin != 0 || W != W_all)) return cudaErrorInvalidValue;
  const int Wb = table_nb ? 1 : W;
  const uint32_t B = 1u << (c - 1);
  // buckets per reduce_seg thread.  A thread's running-sum sweep is a serial chain of 2L additions and
  // the kernel holds 3 blocks per SM, so a launch costs (number of block waves) x L: among the segment
  // lengths that still give the chip enough threads, take the one with the cheapest whole number of
  // waves (2315 jobs x 37 windows x 2 components at L = 64 is 3.02 waves -- a fourth, almost empty wave
  // would cost a quarter of the kernel); longer segments win ties (fewer segment sums to combine).
  uint32_t L = std::min<uint32_t>(kSegLen, B);
  {
    static const uint64_t seg_threads = [] { const char* e = getenv("MP_SEG_THREADS"); return e ? strtoull(e, nullptr, 10) : 150000ull; }();
    static const uint64_t wave = [] {  // resident blocks of k_reduce_seg (launch bounds: 128 threads, 3 per SM)
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      return (uint64_t)sms * 3;
    }();
    auto cost = [&](uint32_t len) {
      const uint64_t blocks = ((uint64_t)njobs * Wb * ncomp * (B / len) + 127) / 128;
      return ((blocks + wave - 1) / wave) * len;
    };
    uint64_t best = cost(L);
    for (uint32_t cand = 2 * L; cand <= B && cand <= 256; cand *= 2) {
      if ((uint64_t)njobs * Wb * ncomp * (B / cand) < seg_threads) break;  // keep the chip full
      if (cost(cand) <= best) { best = cost(cand); L = cand; }
    }
  }
  const uint32_t nseg = B / L;
  const uint64_t nwin = (uint64_t)njobs * Wb;
  const uint64_t nbuckets = nwin * B;
  uint64_t total_terms = 0;
  uint32_t max_len = 0;
  for (int j = 0; j < njobs; j++) {
    total_terms += h_jobs[j].len;
    max_len = std::max(max_len, h_jobs[j].len);
    if ((uint64_t)h_jobs[j].scalar_off + h_jobs[j].len > n_scalars) return cudaErrorInvalidValue;
  }
  const uint64_t max_entries = total_terms * W;
  if (max_entries >= (1ull << 32) || nbuckets >= (1ull << 32)) return cudaErrorInvalidValue;
  // chunk length: longer chunks mean fewer partial sums to stitch; keep >= ~600k threads in flight
  uint32_t kChunk = kChunkMin;
  static const uint32_t chunk_cap_env = [] { const char* e = getenv("MP_ACC_CHUNK_CAP"); return e ? (uint32_t)atoi(e) : 0u; }();
  uint32_t chunk_cap = ncomp == 1 ? kChunkMax / 2 : kChunkMax;  // a warp's staged tile stays <= 8 KB of smem
  if (chunk_cap_env) chunk_cap = std::min(chunk_cap, std::max(kChunk, chunk_cap_env));
  while (kChunk < chunk_cap && max_entries * ncomp / (2 * kChunk) >= 600000) kChunk *= 2;
  // small launches are latency-bound (a thread adds its chunk serially): shorter chunks put more
  // threads to work as long as fewer than ~4 warps per SM would be busy otherwise
  while (kChunk > 8 && max_entries * ncomp / kChunk < 20000) kChunk /= 2;
  const uint64_t max_chunks = (max_entries + kChunk - 1) / kChunk;
  const uint64_t ntiles = (nbuckets + 1 + kScanTile - 1) / kScanTile;

  uint32_t *digits, *counts, *offsets, *cursor, *sorted, *tile_sums;
  MsmJob* d_jobs;
  xyzz *bucket_sums, *part, *segS, *segT, *win_out;
  uint32_t* chunk_bucket;
  MP_CK(ws->get(0, n_scalars * W, &digits));
  MP_CK(ws->get(1, nbuckets + 1, &counts));
  MP_CK(ws->get(2, nbuckets + 1, &offsets));
  MP_CK(ws->get(3, nbuckets + 1, &cursor));
  MP_CK(ws->get(4, max_entries + (size_t)kAccThreads * kChunkMax, &sorted));  // padded: tiles are copied whole
  MP_CK(ws->get(5, ntiles + 1, &tile_sums));
  MP_CK(ws->get(6, (size_t)njobs, &d_jobs));
  MP_CK(ws->get(7, nbuckets * ncomp, &bucket_sums));
  MP_CK(ws->get(8, max_chunks * ncomp, &part));
  MP_CK(ws->get(12, max_chunks + 1, &chunk_bucket));
  MP_CK(ws->get(9, nwin * nseg * ncomp, &segS));
  MP_CK(ws->get(10, nwin * nseg * ncomp, &segT));
  MP_CK(ws->get(11, nwin * ncomp, &win_out));

  MP_CK(cudaMemcpyAsync(d_jobs, h_jobs, sizeof(MsmJob) * njobs, cudaMemcpyHostToDevice, stream));
  MP_CK(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (nbuckets + 1), stream));

  if (n_scalars > 0) {
    k_digits<<<(unsigned)((n_scalars + 255) / 256), 256, 0, stream>>>(d_scalars, digits, n_scalars, c, w_begin, W);
    ws->launches++;
  }
  if (max_len > 0) {
    dim3 grid((unsigned)std::min<uint64_t>((max_len + 255) / 256, 65535), (unsigned)njobs);
    k_count<<<grid, 256, 0, stream>>>(digits, d_jobs, counts, n_scalars, W, B, Wb);
    ws->launches++;
  }
  k_scan_tile_sums<<<(unsigned)ntiles, 256, 0, stream>>>(counts, tile_sums, nbuckets + 1);
  k_scan_top<<<1, 256, 0, stream>>>(tile_sums, (uint32_t)ntiles, tile_sums + ntiles);
  k_scan_apply<<<(unsigned)ntiles, 256, 0, stream>>>(counts, tile_sums, offsets, cursor, nbuckets + 1);
  ws->launches += 3;
  if (max_len > 0) {
    dim3 grid((unsigned)std::min<uint64_t>((max_len + 255) / 256, 65535), (unsigned)njobs);
    k_scatter<<<grid, 256, 0, stream>>>(digits, d_jobs, cursor, sorted, n_scalars, W, B, Wb, table_nb);
    ws->launches++;
  }
  if (max_chunks > 0) {
    MsmWorkspace::Timed tm;
    const bool timed = ws->profile && (int)ws->timed.size() < MsmWorkspace::kMaxTimed;
    if (timed) {
      if (!ws->pinned_counts) MP_CK(cudaMallocHost(&ws->pinned_counts, sizeof(uint32_t) * MsmWorkspace::kMaxTimed));
      MP_CK(cudaEventCreate(&tm.e0));
      MP_CK(cudaEventCreate(&tm.e1));
      tm.ncomp = ncomp;
      tm.entries = ws->pinned_counts + ws->timed.size();
      // exact number of bucket additions = length of the sorted list (zero digits are skipped)
      MP_CK(cudaMemcpyAsync(tm.entries, offsets + nbuckets, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      MP_CK(cudaEventRecord(tm.e0, stream));
    }
    uint64_t threads = max_chunks * ncomp;
    unsigned blocks = (unsigned)((threads + kAccThreads - 1) / kAccThreads);
    // occupancy knob: resident blocks per SM the kernel is compiled for (register cap).  The 8-limb field
    // runs best at 4 blocks (128 registers); so does the 12-limb one now that its multiplication is a call
    // (the caller keeps the accumulator and the prefetched point in callee-saved registers or on the stack).
    constexpr int kMB0 = kFqLimbs > 8 ? 3 : 4;  // lowest compiled variant (kMB0, kMB0 + 1, kMB0 + 2)
    constexpr int kMBDefault = 4;              // measured, 12 limbs: 1.59 / 1.73 / 1.99 G adds/s at 2 / 3 / 4 blocks
    static const int acc_mb = [] { const char* e = getenv("MP_ACC_MINBLOCKS"); return e ? atoi(e) : kMBDefault; }();
    const size_t smem = (size_t)(kAccThreads / ncomp) * kChunk * sizeof(uint32_t);  // one staged tile per warp
#define MP_LAUNCH_ACC(NC, MB) \
  k_accumulate<NC, MB><<<blocks, kAccThreads, smem, stream>>>(sorted, offsets, nbuckets, d_points, bucket_sums, part, chunk_bucket, kChunk)
    if (ncomp == 1) {
      if (acc_mb >= kMB0 + 2) MP_LAUNCH_ACC(1, kMB0 + 2); else if (acc_mb == kMB0 + 1) MP_LAUNCH_ACC(1, kMB0 + 1); else MP_LAUNCH_ACC(1, kMB0);
    } else {
      if (acc_mb >= kMB0 + 2) MP_LAUNCH_ACC(2, kMB0 + 2); else if (acc_mb == kMB0 + 1) MP_LAUNCH_ACC(2, kMB0 + 1); else MP_LAUNCH_ACC(2, kMB0);
    }
#undef MP_LAUNCH_ACC
    ws->launches++;
    if (timed) {
      MP_CK(cudaEventRecord(tm.e1, stream));
      ws->timed.push_back(tm);
    }
  }
  if (max_chunks > 0) {
    k_stitch<<<(unsigned)((max_chunks * ncomp + 127) / 128), 128, 0, stream>>>(offsets, nbuckets, chunk_bucket, ncomp, bucket_sums, part, kChunk);
    ws->launches++;
  }
  {
    static const int seg_mb = [] { const char* e = getenv("MP_SEG_MINBLOCKS"); return e ? atoi(e) : 3; }();
    const unsigned sb = (unsigned)((nwin * nseg * ncomp + 127) / 128);
    if (seg_mb >= 4) k_reduce_seg<4><<<sb, 128, 0, stream>>>(offsets, bucket_sums, nwin, B, L, ncomp, segS, segT);
    else if (seg_mb <= 2) k_reduce_seg<2><<<sb, 128, 0, stream>>>(offsets, bucket_sums, nwin, B, L, ncomp, segS, segT);
    else k_reduce_seg<3><<<sb, 128, 0, stream>>>(offsets, bucket_sums, nwin, B, L, ncomp, segS, segT);
  }
  // Per-window combine of the segment sums: fold groups of 4 segments, level after level, until one segment per
  // window is left -- its T' is the window sum (ONE kernel family for both curves; the block-wide shuffle-scan
  // kernel of round 1 is gone).  A level with few groups runs the quad-cooperative kernel.
  {
    xyzz *ping, *pong;
    const size_t lvl = nwin * ((nseg + 3) / 4) * ncomp;  // outputs of the first (largest) level
    MP_CK(ws->get(14, 2 * lvl, &ping));                  // [S | T] of the even levels
    MP_CK(ws->get(15, 2 * lvl, &pong));                  // [S | T] of the odd levels
    static const uint64_t quad_below = [] { const char* e = getenv("MP_QUAD_BELOW"); return e ? strtoull(e, nullptr, 10) : 40000ull; }();
    const xyzz *curS = segS, *curT = segT;
    uint32_t cur_nseg = nseg, cur_L = L;
    int level = 0;
    do {
      const uint32_t G = std::min<uint32_t>(4, cur_nseg);
      xyzz* buf = (level & 1) ? pong : ping;
      const bool last = cur_nseg == G;
      const uint64_t groups = nwin * (cur_nseg / G) * ncomp;
      if (groups < quad_below)
        k_reduce_group_quad<<<(unsigned)((groups * 4 + 127) / 128), 128, 0, stream>>>(curS, curT, nwin, cur_nseg, G, cur_L, ncomp, buf,
                                                                                   last ? win_out : buf + lvl);
      else
        k_reduce_group<<<(unsigned)((groups + 63) / 64), 64, 0, stream>>>(curS, curT, nwin, cur_nseg, G, cur_L, ncomp, buf,
                                                                        last ? win_out : buf + lvl);
      ws->launches++;
      curS = buf; curT = buf + lvl;
      cur_nseg /= G; cur_L *= G;
      level++;
    } while (cur_nseg > 1);
    ws->launches -= 1;  // the tally below counts one combine launch
  }
  k_fold<<<(unsigned)((njobs * ncomp + 7) / 8), 32, 0, stream>>>(win_out, njobs, Wb, c, ncomp, d_out);
  ws->launches += 3;
  return cudaGetLastError();
}

}  // namespace mp
