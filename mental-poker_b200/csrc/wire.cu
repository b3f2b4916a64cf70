// Wire format, device half: batched decompression of ark-serialize 0.3
// compressed points -- y = sqrt(x^3 + x + b) in F_p, sign chosen by the "larger" flag.  The Stark prime
// has p - 1 = 2^192 * (2^59 + 17), so the square root is Tonelli-Shanks with a 192-bit two-adic part:
// the reference's CPU path (ark-ff 0.3 `SquareRootField::sqrt` behind `CanonicalDeserialize`, reached
// from every type bound at reference src/lib.rs:45-71) spends ~10^4 field multiplications per point.
//
// k_decompress: one thread per point, square root by the windowed form of fq_sqrt.cuh: the 192-bit
// discrete log of a^t is found eight bits at a time from ONE chain of 184 squarings plus table
// multiplications -- 242 squarings + 304 multiplications with the same control flow in every lane.  The
// first version ran the textbook loop (order exponent of b by repeated squaring, per round): ~4 600
// data-dependent squarings per point, lanes diverging on every round -- 16.2 ms for the 131 072 points of
// a deck against 0.79 ms now (ncu, profiles/).
#include "fq_sqrt.cuh"
#include "shuffle_internal.cuh"
#include "wire_host.hpp"

namespace mp {

// tables of fq_sqrt.cuh, built once per context: T | Tinv by one thread, then one thread per entry
__global__ void k_sqrt_chain(fq* __restrict__ T, fq* __restrict__ Tinv) {
  fq_sqrt_table(T);
  fq_sqrt_inverse_table(T, Tinv);
}
__global__ void __launch_bounds__(128) k_sqrt_entries(const fq* __restrict__ T, const fq* __restrict__ Tinv, fq* __restrict__ U,
                                                      fq* __restrict__ V, uint8_t* __restrict__ lut) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < kSqrtUCount + kSqrtVCount + kSqrtRadix) fq_sqrt_fill_entry(T, Tinv, g, U, V, lut);
}

// status: 0 ok, 1 malformed (x not canonical / stray flag bits), 2 x is not the abscissa of a curve point
__global__ void __launch_bounds__(128) k_decompress(const uint32_t* __restrict__ in, uint64_t n, const SqrtTables tb,
                                                    uint32_t* __restrict__ out, uint8_t* __restrict__ status, int* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int L = kFqLimbs;   // 8 (Stark) / 12 (BLS12-377): x is L words, flags in the top two bits of the last one
  uint32_t w[L];
  {
    const uint4* p = reinterpret_cast<const uint4*>(in + i * L);
#pragma unroll
    for (int q = 0; q < L / 4; q++) {
      const uint4 v = __ldg(p + q);
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
  }
  const uint32_t flags = w[L - 1] >> 30;  // bit 1 = larger, bit 0 = infinity
  w[L - 1] &= 0x3fffffffu;
  uint32_t res[2 * L];
#pragma unroll
  for (int k = 0; k < 2 * L; k++) res[k] = 0;
  uint8_t st = 0;
  fq xc;
#pragma unroll
  for (int k = 0; k < L; k++) xc.v[k] = w[k];
  if (flags & 1u) {
    if (!fq_is_zero_raw(xc) || (flags & 2u)) st = 1;
  } else {
    uint32_t borrow;
    fq_sub_raw(xc, fq_kp(1), &borrow);
    if (!borrow) {
      st = 1;  // x >= p
    } else {
      const fq xm = fq_reduce_full(fq_to_mont(xc));
#ifdef MP_CURVE_BLS12_377
      const fq rhs = fq_add(fq_mul(fq_sqr(xm), xm), fq_curve_b());               // y^2 = x^3 + 1
#else
      const fq rhs = fq_add(fq_add(fq_mul(fq_sqr(xm), xm), xm), fq_curve_b());  // y^2 = x^3 + x + b   [2] + [1] + [1]
#endif
      bool ok;
      const fq y = fq_sqrt_win(fq_reduce_weak(rhs), tb, &ok);
      if (!ok) {
        st = 2;
      } else {
        fq yc = fq_from_mont(y);  // canonical
        // larger of (y, p - y)  <=>  y > (p - 1) / 2
        fq half;
#pragma unroll
        for (int k = 0; k < L; k++) half.v[k] = fq_half_limb(k);
        uint32_t le;  // borrow of half - y: set iff y > half
        fq_sub_raw(half, yc, &le);
        if ((le != 0) != ((flags & 2u) != 0)) {
          uint32_t bw;
          yc = fq_is_zero_raw(yc) ? yc : fq_sub_raw(fq_kp(1), yc, &bw);
        }
#pragma unroll
        for (int k = 0; k < L; k++) { res[k] = w[k]; res[L + k] = yc.v[k]; }
      }
    }
  }
  if (st) atomicExch(bad, 1);
  if (status) status[i] = st;
  uint4* o = reinterpret_cast<uint4*>(out + i * 2 * L);
#pragma unroll
  for (int q = 0; q < 2 * L / 4; q++) o[q] = make_uint4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
enum WireSlot { sWireIn = 200, sWireOut, sWireStatus, sWireTable };

// device-resident decompression: d_in n*kWireFe bytes -> d_out n*kWirePt bytes (+ per-item status), asynchronous
static int32_t decompress_device(mp_ctx* ctx, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint8_t* d_status, int* d_bad) {
  // one slot: T | Tinv | U | V | lut
  const size_t fq_count = 2 * (size_t)kTwoAdicity + kSqrtUCount + kSqrtVCount;
  fq* base = (fq*)ctx->scratch(sWireTable, sizeof(fq) * fq_count + 65536 + 64);
  NEED(base);
  fq *T = base, *Tinv = base + kTwoAdicity, *U = Tinv + kTwoAdicity, *V = U + kSqrtUCount;
  uint8_t* lut = reinterpret_cast<uint8_t*>(V + kSqrtVCount);
  if (!ctx->wire_table_ready) {
    const size_t entries = kSqrtUCount + kSqrtVCount + kSqrtRadix;
    CK(cudaMemsetAsync(lut, 0, 65536, ctx->stream));
    k_sqrt_chain<<<1, 1, 0, ctx->stream>>>(T, Tinv);
    k_sqrt_entries<<<(unsigned)((entries + 127) / 128), 128, 0, ctx->stream>>>(T, Tinv, U, V, lut);
    ctx->wire_table_ready = true;
    ctx->launches += 2;
  }
  const SqrtTables tb{U, V, lut};
  k_decompress<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint32_t*)d_in, n, tb, (uint32_t*)d_out, d_status, d_bad);
  CK(cudaGetLastError());
  ctx->launches += 1;
  return MP_OK;
}

int32_t wire_points_decompress(mp_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, int32_t* statuses) {
  if (!ctx || (n && (!in || !out))) return MP_ERR_INVALID_ARG;
  if (n >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "too many points");
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  if (n == 0) return MP_OK;
  uint8_t* d_in = (uint8_t*)ctx->scratch(sWireIn, n * kWireFe);
  uint8_t* d_out = (uint8_t*)ctx->scratch(sWireOut, n * kWirePt);
  uint8_t* d_status = (uint8_t*)ctx->scratch(sWireStatus, n + 64);
  int* d_bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  NEED(d_in); NEED(d_out); NEED(d_status); NEED(d_bad);
  CK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  CK(cudaMemcpyAsync(d_in, in, n * kWireFe, cudaMemcpyHostToDevice, ctx->stream));
  int32_t rc = decompress_device(ctx, d_in, n, d_out, d_status, d_bad);
  if (rc != MP_OK) return rc;
  CK(cudaMemcpyAsync(out, d_out, n * kWirePt, cudaMemcpyDeviceToHost, ctx->stream));
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<uint8_t> st;
  if (statuses) {
    st.resize(n);
    CK(cudaMemcpyAsync(st.data(), d_status, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CK(stream_wait(ctx, ctx->stream));
  if (statuses)
    for (uint64_t i = 0; i < n; i++) statuses[i] = st[i];
  if (bad) return ctx->fail(MP_ERR_NOT_ON_CURVE, "a compressed point is malformed or not on the curve");
  return MP_OK;
}

int32_t wire_deck_deserialize(mp_ctx* ctx, const uint8_t* in, uint64_t in_len, uint8_t* out_deck, uint64_t* n_cards) {
  if (!ctx || !in || !n_cards) return MP_ERR_INVALID_ARG;
  if (in_len < 8) return ctx->fail(MP_ERR_INVALID_ARG, "serialized deck shorter than its length prefix");
  uint64_t n;
  memcpy(&n, in, 8);
  if (n >= (1ull << 28) || in_len != wire_deck_len(n)) return ctx->fail(MP_ERR_INVALID_ARG, "length prefix %llu does not match the buffer", (unsigned long long)n);
  if (*n_cards < n) { *n_cards = n; return ctx->fail(MP_ERR_INVALID_ARG, "output deck too small for %llu cards", (unsigned long long)n); }
  *n_cards = n;
  if (n && !out_deck) return MP_ERR_INVALID_ARG;
  return wire_points_decompress(ctx, in + 8, 2 * n, out_deck, nullptr);
}

int32_t wire_proof_deserialize(mp_ctx* ctx, int32_t m, int32_t n, const uint8_t* in, uint8_t* out_proof) {
  if (!ctx || !in || !out_proof || m < 1 || n < 1) return MP_ERR_INVALID_ARG;
  // gather the compressed points, decompress them in one launch, scatter into the flat layout
  const size_t npts = 11 * (size_t)m + 8;
  std::vector<uint8_t> comp(npts * kWireFe), pts(npts * kWirePt);
  {
    const uint8_t* p = in;
    uint8_t* c = comp.data();
    for (const WireRun& r : wire_proof_runs(m, n)) {
      const size_t len = (r.points ? kWireFe : 32) * r.count;
      if (r.points) { memcpy(c, p, len); c += len; }
      p += len;
    }
  }
  int32_t rc = wire_points_decompress(ctx, comp.data(), npts, pts.data(), nullptr);
  if (rc != MP_OK) return rc;
  const uint8_t *p = in, *q = pts.data();
  for (const WireRun& r : wire_proof_runs(m, n)) {
    if (r.points) {
      memcpy(out_proof, q, kWirePt * r.count);
      q += kWirePt * r.count;
      out_proof += kWirePt * r.count;
    } else {
      // field elements: ark-serialize rejects non-canonical encodings (value >= the group order) at deserialisation
      for (size_t k = 0; k < r.count; k++)
        if (!h_fr_is_canonical(p + 32 * k)) return ctx->fail(MP_ERR_NOT_CANONICAL, "a scalar of the proof is not below the group order");
      memcpy(out_proof, p, 32 * r.count);
      out_proof += 32 * r.count;
    }
    p += (r.points ? kWireFe : 32) * r.count;
  }
  return MP_OK;
}

}  // namespace mp
