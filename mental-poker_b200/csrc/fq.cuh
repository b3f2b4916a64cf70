// Base field F_p of the Stark curve, p = 2^251 + 17*2^192 + 1, 8 x 32-bit limbs, Montgomery
// form with R = 2^256 (the representation ark-ff 0.3 `Fp256` uses on the reference's CPU path,
// reference crate Cargo.toml:12).
//
// Two properties of this prime shape everything here:
//   * p == 1 (mod 2^192)  =>  -p^-1 mod 2^256 = -1 + 17*2^192 + 2^251, so the Montgomery
//     quotient M = T_lo * (-p^-1) mod 2^256 needs ONE 32x32 multiply (17*t0) instead of 64,
//     and M*p = M + (17 + 2^59) * M * 2^192 needs 8 (17*M) instead of 64.
//     A field multiplication is 64 + ~10 IMAD.WIDE instead of 128.
//   * p < 2^252 leaves 4 spare bits in 256, so values are kept LAZILY reduced: every element
//     is an integer < 2^256 congruent to the value, with a statically tracked bound k*p
//     (written [k] in comments).  fq_mul accepts [x]*[y] with x*y <= 30 and returns [2].
//     Additions do not reduce; subtractions add a multiple of p; `fq_reduce_weak` brings any
//     256-bit value back to [2] with ~12 ALU-pipe instructions.
//
// Everything is __host__ __device__: on the device the carry chains are inline PTX
// (mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32(.X)); on the host the
// same word-level algorithm runs on uint64 arithmetic so the logic is unit-testable without
// a GPU (tests/host_field_test.cpp is NOT part of the product path).
#pragma once
#ifdef MP_CURVE_BLS12_377
// second curve: same function names over a 12-limb field; the translation units
// of that build are compiled with -Dmp=mp_bls12_377 so both instantiations link into one library
#include "fq_bls12_377.cuh"
#else
#include <stdint.h>

#ifdef __CUDACC__
#define MP_HD __host__ __device__ __forceinline__
#define MP_D __device__ __forceinline__
#else
#define MP_HD inline
#define MP_D inline
#endif

#define MP_CURVE_NAME "Stark curve"
#define MP_CURVE_A_IS_ZERO 0

namespace mp {

static constexpr int kFqLimbs = 8;
// bits the signed-digit recoding must cover: bit length of the group order (252) + 1 for the carry
static constexpr int kScalarBits = 253;

struct fq {
  uint32_t v[8];
};

// p, little-endian u32 limbs
#define MP_P0 0x00000001u
#define MP_P6 0x00000011u
#define MP_P7 0x08000000u

MP_HD fq fq_zero() {
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}

// k * p for small k (k <= 31 so it fits 256 bits): limbs {k, 0,0,0,0,0, 17k, k<<27}
MP_HD fq fq_kp(uint32_t k) {
  fq r = fq_zero();
  r.v[0] = k;
  r.v[6] = 17u * k;
  r.v[7] = k << 27;
  return r;
}

// R mod p  (Montgomery one) and R^2 mod p (to-Montgomery multiplier)   [SURVEY.md A7]
MP_HD fq fq_one() {
  fq r;
  r.v[0] = 0xffffffe1u; r.v[1] = 0xffffffffu; r.v[2] = 0xffffffffu; r.v[3] = 0xffffffffu;
  r.v[4] = 0xffffffffu; r.v[5] = 0xffffffffu; r.v[6] = 0xfffffdf0u; r.v[7] = 0x07ffffffu;
  return r;
}
MP_HD fq fq_r2() {
  fq r;
  r.v[0] = 0x7e000401u; r.v[1] = 0xfffffd73u; r.v[2] = 0x330fffffu; r.v[3] = 0x00000001u;
  r.v[4] = 0xff6f8000u; r.v[5] = 0xffffffffu; r.v[6] = 0x5e008810u; r.v[7] = 0x07ffd4abu;
  return r;
}

// curve coefficient b in Montgomery form (b * R mod p)
MP_HD fq fq_curve_b() {
  fq r;
  r.v[0] = 0xb59a21cau; r.v[1] = 0x359ddd67u; r.v[2] = 0x7aab9006u; r.v[3] = 0x6725f223u;
  r.v[4] = 0x2a41f947u; r.v[5] = 0xab8a1e00u; r.v[6] = 0x1774247fu; r.v[7] = 0x01393165u;
  return r;
}

// ------------------------------------------------------------------------------------------
// word-level primitives
// ------------------------------------------------------------------------------------------

// r = a + b  (256-bit, carry out dropped: callers guarantee the bound fits)
MP_HD fq fq_add(const fq& a, const fq& b) {
  fq r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
        "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return r;
}

// r = a - b (256-bit wrap-around); *borrow = 1 if a < b
MP_HD fq fq_sub_raw(const fq& a, const fq& b, uint32_t* borrow) {
  fq r;
#ifdef __CUDA_ARCH__
  uint32_t bw;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7]), "=r"(bw)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]),
        "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  *borrow = bw & 1u;
#else
  int64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (int64_t)a.v[i] - (int64_t)b.v[i];
    r.v[i] = (uint32_t)c;
    c >>= 32;  // arithmetic shift: 0 or -1
  }
  *borrow = (uint32_t)(c & 1);
#endif
  return r;
}

// r = a - k*p, exploiting the three non-zero limbs of p (k*17 and k<<27 must fit: k <= 31)
MP_HD fq fq_sub_kp(const fq& a, uint32_t k) {
  fq r;
  uint32_t k17 = 17u * k, k27 = k << 27;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, 0;\n\t"
      "subc.cc.u32 %2, %10, 0;\n\t"
      "subc.cc.u32 %3, %11, 0;\n\t"
      "subc.cc.u32 %4, %12, 0;\n\t"
      "subc.cc.u32 %5, %13, 0;\n\t"
      "subc.cc.u32 %6, %14, %17;\n\t"
      "subc.u32 %7, %15, %18;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(k), "r"(k17), "r"(k27));
#else
  fq kp = fq_zero();
  kp.v[0] = k; kp.v[6] = k17; kp.v[7] = k27;
  uint32_t bw;
  r = fq_sub_raw(a, kp, &bw);
#endif
  return r;
}

// Any 256-bit value -> [2] (precisely: < 2^252 < 2p), same residue.  q = v >> 251; for
// q >= 1 subtract (q-1)*p: the result is 2^251 + (v mod 2^251) - (q-1)(17*2^192+1) > 0.
MP_HD fq fq_reduce_weak(const fq& a) {
  uint32_t q = a.v[7] >> 27;
  uint32_t k = q - (q != 0 ? 1u : 0u);
  return fq_sub_kp(a, k);
}

// a in [k] (k given) minus b in [kb]: returns a + kb*p - b   -> [k + kb]
MP_HD fq fq_sub(const fq& a, const fq& b, uint32_t kb) {
  uint32_t bw;
  fq t = fq_add(a, fq_kp(kb));
  return fq_sub_raw(t, b, &bw);
}

// canonical representative in [0, p) of any 256-bit value
MP_HD fq fq_reduce_full(const fq& a) {
  fq t = fq_reduce_weak(a);  // < 2^252 < 2p + ...; in fact < 2p
  uint32_t bw;
  fq u = fq_sub_raw(t, fq_kp(1), &bw);
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = bw ? t.v[i] : u.v[i];
  return r;
}

MP_HD bool fq_is_zero_raw(const fq& a) {
  return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}

// a in [2]: is a == 0 (mod p)?  (a is 0 or p)
MP_HD bool fq_is_zero_mod_p_2(const fq& a) {
  uint32_t mid = a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5];
  uint32_t z = mid | a.v[0] | a.v[6] | a.v[7];
  uint32_t e = mid | (a.v[0] ^ MP_P0) | (a.v[6] ^ MP_P6) | (a.v[7] ^ MP_P7);
  return (z == 0) | (e == 0);
}

MP_HD bool fq_eq_raw(const fq& a, const fq& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// ------------------------------------------------------------------------------------------
// 8x8 schoolbook product, T[16] = a * b
// ------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
// One row chain: acc[0..7] += {a0,a1,a2,a3} * b laid out as four adjacent 64-bit products;
// carry out is added into acc8.  lo/hi pairs fuse into IMAD.WIDE.U32 with carry predicates.
#define MP_ROW_CHAIN(c0, c1, c2, c3, c4, c5, c6, c7, c8, a0, a1, a2, a3, b)                  \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                   \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                  \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                 \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                 \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                 \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                 \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                 \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                 \
      "addc.u32 %8, %8, 0;"                                                                  \
      : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6), "+r"(c7),      \
        "+r"(c8)                                                                             \
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b))

__device__ __forceinline__ void fq_mul_wide(uint32_t* __restrict__ T, const fq& a, const fq& b) {
  // ev[k] = word k of the sum of products at even word offsets; od[k] = word k+1 of the
  // sum of products at odd word offsets.
  uint32_t ev[17], od[17];
#pragma unroll
  for (int i = 0; i < 17; i++) { ev[i] = 0; od[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    // row i (even): even-j products at offset i+j -> ev[i..i+7]; odd-j -> od[i..i+7]
    MP_ROW_CHAIN(ev[i], ev[i + 1], ev[i + 2], ev[i + 3], ev[i + 4], ev[i + 5], ev[i + 6],
                 ev[i + 7], ev[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
    MP_ROW_CHAIN(od[i], od[i + 1], od[i + 2], od[i + 3], od[i + 4], od[i + 5], od[i + 6],
                 od[i + 7], od[i + 8], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    // row i+1 (odd): even-j products at offset i+1+j (odd) -> od[i..i+7];
    //                odd-j products at offset i+1+j (even) -> ev[i+2..i+9]
    MP_ROW_CHAIN(od[i], od[i + 1], od[i + 2], od[i + 3], od[i + 4], od[i + 5], od[i + 6],
                 od[i + 7], od[i + 8], a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1]);
    MP_ROW_CHAIN(ev[i + 2], ev[i + 3], ev[i + 4], ev[i + 5], ev[i + 6], ev[i + 7], ev[i + 8],
                 ev[i + 9], ev[i + 10], a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
  }
  // T = ev + (od << 32)
  T[0] = ev[0];
  asm("add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;"
      : "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]),
        "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]),
        "=r"(T[14]), "=r"(T[15])
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]),
        "r"(ev[14]), "r"(ev[15]), "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]),
        "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]),
        "r"(od[12]), "r"(od[13]), "r"(od[14]));
}
#else
inline void fq_mul_wide(uint32_t* T, const fq& a, const fq& b) {
  for (int i = 0; i < 16; i++) T[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 8; j++) {
      uint64_t cur = (uint64_t)T[i + j] + (uint64_t)a.v[j] * b.v[i] + carry;
      T[i + j] = (uint32_t)cur;
      carry = cur >> 32;
    }
    T[i + 8] = (uint32_t)carry;
  }
}
#endif

// ------------------------------------------------------------------------------------------
// Dedicated squaring: T[16] = a^2 with 28 off-diagonal + 8 diagonal products (36 IMAD.WIDE
// instead of 64); the doubling and the diagonal add run on the ALU pipe, which has slack.
// ------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define MP_CHAIN1(c0, c1, c2, a0, b)                      \
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"                 \
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"                \
      "addc.u32 %2, %2, 0;"                               \
      : "+r"(c0), "+r"(c1), "+r"(c2)                      \
      : "r"(a0), "r"(b))
#define MP_CHAIN2(c0, c1, c2, c3, c4, a0, a1, b)          \
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"                 \
      "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"                \
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"                \
      "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"                \
      "addc.u32 %4, %4, 0;"                               \
      : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4)  \
      : "r"(a0), "r"(a1), "r"(b))
#define MP_CHAIN3(c0, c1, c2, c3, c4, c5, c6, a0, a1, a2, b)                  \
  asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"                                    \
      "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"                                   \
      "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"                                   \
      "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"                                   \
      "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"                                   \
      "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"                                   \
      "addc.u32 %6, %6, 0;"                                                   \
      : "+r"(c0), "+r"(c1), "+r"(c2), "+r"(c3), "+r"(c4), "+r"(c5), "+r"(c6)  \
      : "r"(a0), "r"(a1), "r"(a2), "r"(b))

__device__ __forceinline__ void fq_sqr_wide(uint32_t* __restrict__ T, const fq& a) {
  // ev[k] = word k of the off-diagonal products at even word offsets i + j;
  // od[k] = word k + 1 of those at odd offsets (same convention as fq_mul_wide).
  uint32_t ev[16], od[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
  const uint32_t* v = a.v;
  // row 0: odd offsets j = 1,3,5,7 -> od[0..7]; even offsets j = 2,4,6 -> ev[2..7]
  MP_ROW_CHAIN(od[0], od[1], od[2], od[3], od[4], od[5], od[6], od[7], od[8], v[1], v[3], v[5], v[7], v[0]);
  MP_CHAIN3(ev[2], ev[3], ev[4], ev[5], ev[6], ev[7], ev[8], v[2], v[4], v[6], v[0]);
  // row 1: odd offsets j = 2,4,6 -> od[2..7]; even offsets j = 3,5,7 -> ev[4..9]
  MP_CHAIN3(od[2], od[3], od[4], od[5], od[6], od[7], od[8], v[2], v[4], v[6], v[1]);
  MP_CHAIN3(ev[4], ev[5], ev[6], ev[7], ev[8], ev[9], ev[10], v[3], v[5], v[7], v[1]);
  // row 2: odd j = 3,5,7 -> od[4..9]; even j = 4,6 -> ev[6..9]
  MP_CHAIN3(od[4], od[5], od[6], od[7], od[8], od[9], od[10], v[3], v[5], v[7], v[2]);
  MP_CHAIN2(ev[6], ev[7], ev[8], ev[9], ev[10], v[4], v[6], v[2]);
  // row 3: odd j = 4,6 -> od[6..9]; even j = 5,7 -> ev[8..11]
  MP_CHAIN2(od[6], od[7], od[8], od[9], od[10], v[4], v[6], v[3]);
  MP_CHAIN2(ev[8], ev[9], ev[10], ev[11], ev[12], v[5], v[7], v[3]);
  // row 4: odd j = 5,7 -> od[8..11]; even j = 6 -> ev[10..11]
  MP_CHAIN2(od[8], od[9], od[10], od[11], od[12], v[5], v[7], v[4]);
  MP_CHAIN1(ev[10], ev[11], ev[12], v[6], v[4]);
  // row 5: odd j = 6 -> od[10..11]; even j = 7 -> ev[12..13]
  MP_CHAIN1(od[10], od[11], od[12], v[6], v[5]);
  MP_CHAIN1(ev[12], ev[13], ev[14], v[7], v[5]);
  // row 6: odd j = 7 -> od[12..13]
  MP_CHAIN1(od[12], od[13], od[14], v[7], v[6]);
  // off = ev + (od << 32): words 1..15 (word 0 is zero: the lowest off-diagonal offset is 1)
  uint32_t off[16];
  off[0] = 0;
  asm("add.cc.u32 %0, %15, %30;\n\t"
      "addc.cc.u32 %1, %16, %31;\n\t"
      "addc.cc.u32 %2, %17, %32;\n\t"
      "addc.cc.u32 %3, %18, %33;\n\t"
      "addc.cc.u32 %4, %19, %34;\n\t"
      "addc.cc.u32 %5, %20, %35;\n\t"
      "addc.cc.u32 %6, %21, %36;\n\t"
      "addc.cc.u32 %7, %22, %37;\n\t"
      "addc.cc.u32 %8, %23, %38;\n\t"
      "addc.cc.u32 %9, %24, %39;\n\t"
      "addc.cc.u32 %10, %25, %40;\n\t"
      "addc.cc.u32 %11, %26, %41;\n\t"
      "addc.cc.u32 %12, %27, %42;\n\t"
      "addc.cc.u32 %13, %28, %43;\n\t"
      "addc.u32 %14, %29, %44;"
      : "=r"(off[1]), "=r"(off[2]), "=r"(off[3]), "=r"(off[4]), "=r"(off[5]), "=r"(off[6]), "=r"(off[7]),
        "=r"(off[8]), "=r"(off[9]), "=r"(off[10]), "=r"(off[11]), "=r"(off[12]), "=r"(off[13]),
        "=r"(off[14]), "=r"(off[15])
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]),
        "r"(ev[14]), "r"(ev[15]), "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]),
        "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]),
        "r"(od[12]), "r"(od[13]), "r"(od[14]));
  // dbl = 2 * off (funnel shifts), then T = dbl + sum a_i^2 * 2^(64 i)
  uint32_t dbl[16];
  dbl[0] = 0;
#pragma unroll
  for (int i = 1; i < 16; i++) dbl[i] = __funnelshift_l(off[i - 1], off[i], 1);
  uint32_t dl[8], dh[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    dl[i] = v[i] * v[i];
    dh[i] = __umulhi(v[i], v[i]);
  }
  asm("add.cc.u32 %0, %16, %32;\n\t"
      "addc.cc.u32 %1, %17, %33;\n\t"
      "addc.cc.u32 %2, %18, %34;\n\t"
      "addc.cc.u32 %3, %19, %35;\n\t"
      "addc.cc.u32 %4, %20, %36;\n\t"
      "addc.cc.u32 %5, %21, %37;\n\t"
      "addc.cc.u32 %6, %22, %38;\n\t"
      "addc.cc.u32 %7, %23, %39;\n\t"
      "addc.cc.u32 %8, %24, %40;\n\t"
      "addc.cc.u32 %9, %25, %41;\n\t"
      "addc.cc.u32 %10, %26, %42;\n\t"
      "addc.cc.u32 %11, %27, %43;\n\t"
      "addc.cc.u32 %12, %28, %44;\n\t"
      "addc.cc.u32 %13, %29, %45;\n\t"
      "addc.cc.u32 %14, %30, %46;\n\t"
      "addc.u32 %15, %31, %47;"
      : "=r"(T[0]), "=r"(T[1]), "=r"(T[2]), "=r"(T[3]), "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]),
        "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
      : "r"(dbl[0]), "r"(dbl[1]), "r"(dbl[2]), "r"(dbl[3]), "r"(dbl[4]), "r"(dbl[5]), "r"(dbl[6]), "r"(dbl[7]),
        "r"(dbl[8]), "r"(dbl[9]), "r"(dbl[10]), "r"(dbl[11]), "r"(dbl[12]), "r"(dbl[13]), "r"(dbl[14]), "r"(dbl[15]),
        "r"(dl[0]), "r"(dh[0]), "r"(dl[1]), "r"(dh[1]), "r"(dl[2]), "r"(dh[2]), "r"(dl[3]), "r"(dh[3]),
        "r"(dl[4]), "r"(dh[4]), "r"(dl[5]), "r"(dh[5]), "r"(dl[6]), "r"(dh[6]), "r"(dl[7]), "r"(dh[7]));
}
#else
inline void fq_sqr_wide(uint32_t* T, const fq& a) { fq_mul_wide(T, a, a); }
#endif

// ------------------------------------------------------------------------------------------
// Sparse-prime Montgomery reduction:  r = (T + M*p) / 2^256,  M = T_lo * (-p^-1) mod 2^256.
//   k  = (17 + 2^59) * T_lo  mod 2^64                       (2 words)
//   M  = k*2^192 - T_lo      mod 2^256,  e = borrow of that subtraction
//   r  = T_hi + e + ((k + 17*M + (M << 59)) >> 64)          (the low 64 bits cancel exactly)
// Output < T/2^256 + p.
// ------------------------------------------------------------------------------------------
MP_HD fq fq_mont_reduce(const uint32_t* T) {
  // k
  uint64_t t0_17 = (uint64_t)T[0] * 17u;
  uint32_t k0 = (uint32_t)t0_17;
  uint32_t k1 = (uint32_t)(t0_17 >> 32) + 17u * T[1] + (T[0] << 27);
  // M = (k << 192) - T_lo
  fq kk = fq_zero();
  kk.v[6] = k0;
  kk.v[7] = k1;
  fq tlo;
#pragma unroll
  for (int i = 0; i < 8; i++) tlo.v[i] = T[i];
  uint32_t e;
  fq M = fq_sub_raw(kk, tlo, &e);
  // V = k + 17*M + (M << 59); words 0..9.  (M<<59) = (M<<27) moved up one word.
  uint32_t sh[10];
  sh[0] = 0;
  sh[1] = M.v[0] << 27;
#pragma unroll
  for (int i = 1; i < 8; i++) sh[i + 1] = (M.v[i] << 27) | (M.v[i - 1] >> 5);
  sh[9] = M.v[7] >> 5;
  fq r;
#ifdef __CUDA_ARCH__
  // X = sh + k  (k in words 0,1)
  uint32_t X[10];
  asm("add.cc.u32 %0, %10, %20;\n\t"
      "addc.cc.u32 %1, %11, %21;\n\t"
      "addc.cc.u32 %2, %12, 0;\n\t"
      "addc.cc.u32 %3, %13, 0;\n\t"
      "addc.cc.u32 %4, %14, 0;\n\t"
      "addc.cc.u32 %5, %15, 0;\n\t"
      "addc.cc.u32 %6, %16, 0;\n\t"
      "addc.cc.u32 %7, %17, 0;\n\t"
      "addc.cc.u32 %8, %18, 0;\n\t"
      "addc.u32 %9, %19, 0;"
      : "=r"(X[0]), "=r"(X[1]), "=r"(X[2]), "=r"(X[3]), "=r"(X[4]), "=r"(X[5]), "=r"(X[6]),
        "=r"(X[7]), "=r"(X[8]), "=r"(X[9])
      : "r"(sh[0]), "r"(sh[1]), "r"(sh[2]), "r"(sh[3]), "r"(sh[4]), "r"(sh[5]), "r"(sh[6]),
        "r"(sh[7]), "r"(sh[8]), "r"(sh[9]), "r"(k0), "r"(k1));
  // V = X + 17*M, even/odd split: ev gets 17*M[0,2,4,6] (+X), od gets 17*M[1,3,5,7]
  uint32_t ev[10], od[9];
#pragma unroll
  for (int i = 0; i < 10; i++) ev[i] = X[i];
#pragma unroll
  for (int i = 0; i < 9; i++) od[i] = 0;
  const uint32_t c17 = 17u;
  asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\t"
      "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
      "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
      "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
      "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
      "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
      "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
      "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
      "addc.cc.u32 %8, %8, 0;\n\t"
      "addc.u32 %9, %9, 0;"   // V < 2^316: word 9 cannot overflow
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]),
        "+r"(ev[6]), "+r"(ev[7]), "+r"(ev[8]), "+r"(ev[9])
      : "r"(M.v[0]), "r"(M.v[2]), "r"(M.v[4]), "r"(M.v[6]), "r"(c17));
  MP_ROW_CHAIN(od[0], od[1], od[2], od[3], od[4], od[5], od[6], od[7], od[8], M.v[1], M.v[3],
               M.v[5], M.v[7], c17);
  // V words 1.. = ev[1..9] + od[0..8]; we need V >> 64 = words 2..9 plus the carry from word 1
  uint32_t V[8];
  uint32_t w1;
  asm("add.cc.u32 %0, %9, %18;\n\t"
      "addc.cc.u32 %1, %10, %19;\n\t"
      "addc.cc.u32 %2, %11, %20;\n\t"
      "addc.cc.u32 %3, %12, %21;\n\t"
      "addc.cc.u32 %4, %13, %22;\n\t"
      "addc.cc.u32 %5, %14, %23;\n\t"
      "addc.cc.u32 %6, %15, %24;\n\t"
      "addc.cc.u32 %7, %16, %25;\n\t"
      "addc.u32 %8, %17, %26;"
      : "=r"(w1), "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]),
        "=r"(V[6]), "=r"(V[7])
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(ev[8]), "r"(ev[9]), "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]),
        "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]));
  (void)w1;  // word 1 of V only feeds the carry into word 2
  // r = T_hi + V + e
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7])
      : "r"(T[8]), "r"(T[9]), "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]),
        "r"(T[15]), "r"(V[0]), "r"(V[1]), "r"(V[2]), "r"(V[3]), "r"(V[4]), "r"(V[5]),
        "r"(V[6]), "r"(V[7]));
  // + e (0/1): fold into limb 0 with a carry chain
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.cc.u32 %4, %4, 0;\n\t"
      "addc.cc.u32 %5, %5, 0;\n\t"
      "addc.cc.u32 %6, %6, 0;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]),
        "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(e));
#else
  // portable: V = k + 17*M + (M<<59) over 10 words, then r = T_hi + e + (V >> 64)
  uint32_t V[10];
  {
    unsigned __int128 acc = 0;
    uint32_t m17lo[10] = {0}, m17hi[10] = {0};
    for (int i = 0; i < 8; i++) {
      uint64_t pr = (uint64_t)M.v[i] * 17u;
      m17lo[i] = (uint32_t)pr;
      m17hi[i + 1] = (uint32_t)(pr >> 32);
    }
    for (int i = 0; i < 10; i++) {
      acc += (unsigned __int128)m17lo[i] + m17hi[i] + sh[i] + ((i == 0) ? k0 : (i == 1 ? k1 : 0));
      V[i] = (uint32_t)acc;
      acc >>= 32;
    }
  }
  uint64_t cc = e;
  for (int i = 0; i < 8; i++) {
    cc += (uint64_t)T[8 + i] + V[2 + i];
    r.v[i] = (uint32_t)cc;
    cc >>= 32;
  }
#endif
  return r;
}

// a in [x], b in [y], x*y <= 30  ->  a*b/R in [2]
MP_HD fq fq_mul(const fq& a, const fq& b) {
  uint32_t T[16];
  fq_mul_wide(T, a, b);
  return fq_mont_reduce(T);
}
MP_HD fq fq_sqr(const fq& a) {
  uint32_t T[16];
  fq_sqr_wide(T, a);
  return fq_mont_reduce(T);
}

// canonical integer (< p, non-Montgomery)  ->  Montgomery [2]
MP_HD fq fq_to_mont(const fq& a) { return fq_mul(a, fq_r2()); }
// Montgomery (any [k], k <= 30) -> canonical integer in [0, p)
MP_HD fq fq_from_mont(const fq& a) {
  uint32_t T[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { T[i] = a.v[i]; T[8 + i] = 0; }
  return fq_reduce_full(fq_mont_reduce(T));
}

MP_HD fq fq_neg2(const fq& a) {  // a in [2] -> 2p - a in [2] (nonzero stays nonzero mod p)
  uint32_t bw;
  return fq_sub_raw(fq_kp(2), a, &bw);
}

// a^(p-2) by 4-bit fixed windows; a in [2] (Montgomery), result [2].  a == 0 -> 0.
MP_HD fq fq_inv(const fq& a) {
  fq tab[16];
  tab[0] = fq_one();
  tab[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tab[i] = fq_mul(tab[i - 1], a);
  // p - 2 = 0x0800000000000010 ffffffffffffffff ffffffffffffffff ffffffffffffffff
  const uint32_t e[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu,
                         0xffffffffu, 0xffffffffu, 0x00000010u, 0x08000000u};
  fq r = fq_one();
#pragma unroll 1
  for (int w = 62; w >= 0; w--) {  // 252 bits = 63 nibbles, top nibble index 62
    r = fq_sqr(r); r = fq_sqr(r); r = fq_sqr(r); r = fq_sqr(r);
    uint32_t nib = (e[w >> 3] >> ((w & 7) * 4)) & 0xfu;
    r = fq_mul(r, tab[nib]);
  }
  return r;
}

}  // namespace mp
#endif  // MP_CURVE_BLS12_377
