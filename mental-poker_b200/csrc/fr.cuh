// Scalar field F_n of the curve fq.cuh selects (Stark: n = group order, 252 bits; under MP_CURVE_BLS12_377
// the 253-bit ark_bls12_377::Fr -- same limb count, only the constants differ), 8 x 32-bit limbs,
// Montgomery form with R = 2^256, always fully reduced (< n).  This is the field the
// reference's protocol scalars live in (`C::ScalarField`, ark-ff 0.3 `Fp256`; reference
// barnett-smart-card-protocol/src/lib.rs:43, Cargo.toml:12).  n has no
// special shape, so this is generic CIOS Montgomery multiplication; it runs both in device
// kernels (the O(N) and O(m^2 n) scalar-vector work of the shuffle argument) and on the host
// (the O(m + n) glue around the Fiat-Shamir transcript).
//
// The Montgomery representation is bit-identical to ark-ff's (same R, same limbs viewed as
// 4 x u64), which matters once: a Fiat-Shamir challenge is the raw ChaCha20 output
// *interpreted as the Montgomery representation* (SURVEY.md A1), so it is loaded with
// fr_from_raw_mont and never multiplied by R^2.
#pragma once
#include <stdint.h>

#include "fq.cuh"

namespace mp {

struct fr {
  uint32_t v[8];
};

#ifdef MP_CURVE_BLS12_377
// r = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001 (253 bits), ark_bls12_377::Fr
#define MP_FR_NINV 0xffffffffu
#define MP_FR_NINV64 0x0a117fffffffffffull
static constexpr int kFrShaveBits = 3;  // ark-ff REPR_SHAVE_BITS = 256 - 253
MP_HD uint32_t fr_modulus_limb(int i) {
  switch (i) {
    case 0: return 0x00000001u;
    case 1: return 0x0a118000u;
    case 2: return 0xd0000001u;
    case 3: return 0x59aa76feu;
    case 4: return 0x5c37b001u;
    case 5: return 0x60b44d1eu;
    case 6: return 0x9a2ca556u;
    default: return 0x12ab655eu;
  }
}
#else
#define MP_FR_NINV 0xe8bde631u
#define MP_FR_NINV64 0xbb6b3c4ce8bde631ull
static constexpr int kFrShaveBits = 4;  // ark-ff REPR_SHAVE_BITS = 256 - 252

MP_HD uint32_t fr_modulus_limb(int i) {
  switch (i) {
    case 0: return 0xadc64d2fu;
    case 1: return 0x1e66a241u;
    case 2: return 0xcae7b232u;
    case 3: return 0xb781126du;
    case 4: return 0xffffffffu;
    case 5: return 0xffffffffu;
    case 6: return 0x00000010u;
    default: return 0x08000000u;
  }
}
#endif

MP_HD fr fr_zero() {
  fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = 0;
  return r;
}
#ifdef MP_CURVE_BLS12_377
MP_HD fr fr_one() {  // R mod r
  fr r;
  r.v[0] = 0xfffffff3u; r.v[1] = 0x7d1c7fffu; r.v[2] = 0x6ffffff2u; r.v[3] = 0x7257f50fu;
  r.v[4] = 0x512c0feeu; r.v[5] = 0x16d81575u; r.v[6] = 0x2bbb9a9du; r.v[7] = 0x0d4bda32u;
  return r;
}
MP_HD fr fr_r2() {  // R^2 mod r
  fr r;
  r.v[0] = 0xb861857bu; r.v[1] = 0x25d577bau; r.v[2] = 0x8860591fu; r.v[3] = 0xcc2c27b5u;
  r.v[4] = 0xe5dc8593u; r.v[5] = 0xa7cc008fu; r.v[6] = 0xeff1c939u; r.v[7] = 0x011fdae7u;
  return r;
}
#else
MP_HD fr fr_one() {  // R mod n
  fr r;
  r.v[0] = 0xf4fca74fu; r.v[1] = 0x51925a0bu; r.v[2] = 0x6df16beeu; r.v[3] = 0xc75ec4b4u;
  r.v[4] = 0x00000008u; r.v[5] = 0x00000000u; r.v[6] = 0xfffffdf1u; r.v[7] = 0x07ffffffu;
  return r;
}
MP_HD fr fr_r2() {  // R^2 mod n
  fr r;
  r.v[0] = 0xea1c688du; r.v[1] = 0x6021b3f1u; r.v[2] = 0x14ce60b9u; r.v[3] = 0x509cf64du;
  r.v[4] = 0xf78bbabbu; r.v[5] = 0xbaf0ab4cu; r.v[6] = 0x2333766eu; r.v[7] = 0x07d9e57cu;
  return r;
}
#endif

MP_HD bool fr_is_zero(const fr& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}
MP_HD bool fr_eq(const fr& a, const fr& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

// r = a - n if a >= n else a      (a < 2n)
MP_HD fr fr_cond_sub(const fr& a, uint32_t extra_carry) {
  fr d;
  int64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (int64_t)a.v[i] - (int64_t)fr_modulus_limb(i);
    d.v[i] = (uint32_t)c;
    c >>= 32;
  }
  // borrow (c == -1) means a < n, unless the value carried a 257th bit
  bool keep = (c != 0) && (extra_carry == 0);
  fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = keep ? a.v[i] : d.v[i];
  return r;
}

MP_HD fr fr_add(const fr& a, const fr& b) {
  fr s;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    s.v[i] = (uint32_t)c;
    c >>= 32;
  }
  return fr_cond_sub(s, (uint32_t)c);  // n < 2^253: the sum never carries, c == 0
}
MP_HD fr fr_sub(const fr& a, const fr& b) {
  fr d;
  int64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (int64_t)a.v[i] - (int64_t)b.v[i];
    d.v[i] = (uint32_t)c;
    c >>= 32;
  }
  if (c != 0) {  // borrow: add n back
    uint64_t k = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      k += (uint64_t)d.v[i] + fr_modulus_limb(i);
      d.v[i] = (uint32_t)k;
      k >>= 32;
    }
  }
  return d;
}
MP_HD fr fr_neg(const fr& a) { return fr_sub(fr_zero(), a); }

#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
// HOST form of the product below: the same CIOS over 4 x 64-bit limbs (unsigned __int128 products, mulx/adx on x86).
// The transcripts' scalar algebra, the lockstep provers and the verifier plans run on host threads; with the 32-bit
// form a 52-card proof cost 0.17 ms of host arithmetic, which is what bounded the batched provers once several GPUs
// shared one host.  Same result: both compute a*b/2^256 mod n, fully reduced.
// -n^-1 mod 2^64 as a per-curve literal (MP_FR_NINV64 beside MP_FR_NINV): a function-local static here would be one
// symbol shared by every library in the process that includes this header, and the two curves' values differ.
static_assert((uint32_t)MP_FR_NINV64 == MP_FR_NINV, "64-bit and 32-bit Montgomery constants disagree");
inline uint64_t fr_ninv64() { return MP_FR_NINV64; }
#endif

// CIOS Montgomery product, output fully reduced
MP_HD fr fr_mul(const fr& a, const fr& b) {
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
  typedef unsigned __int128 u128;
  uint64_t A[4], B[4], M[4], t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    A[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
    B[i] = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
    M[i] = (uint64_t)fr_modulus_limb(2 * i) | ((uint64_t)fr_modulus_limb(2 * i + 1) << 32);
  }
  const uint64_t ninv = fr_ninv64();
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)t[j] + (u128)A[j] * B[i];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    const uint64_t q = t[0] * ninv;
    c = ((u128)t[0] + (u128)q * M[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)t[j] + (u128)q * M[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fr r64;
  for (int i = 0; i < 4; i++) { r64.v[2 * i] = (uint32_t)t[i]; r64.v[2 * i + 1] = (uint32_t)(t[i] >> 32); }
  return fr_cond_sub(r64, (uint32_t)t[4]);
#else
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      c += (uint64_t)t[j] + (uint64_t)a.v[j] * b.v[i];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[8] = (uint32_t)c;
    t[9] = (uint32_t)(c >> 32);
    uint32_t q = t[0] * MP_FR_NINV;
    c = ((uint64_t)t[0] + (uint64_t)q * fr_modulus_limb(0)) >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      c += (uint64_t)t[j] + (uint64_t)q * fr_modulus_limb(j);
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[7] = (uint32_t)c;
    t[8] = t[9] + (uint32_t)(c >> 32);
  }
  fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  return fr_cond_sub(r, t[8]);
#endif
}
MP_HD fr fr_sqr(const fr& a) { return fr_mul(a, a); }

// canonical integer (8 words LE, any 256-bit value) -> Montgomery.  fr_mul tolerates
// unreduced inputs below 2^256 when the other operand is < n: the result is < 2n before the
// final conditional subtraction.
MP_HD fr fr_from_canonical(const uint32_t* w) {
  fr a;
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = w[i];
  return fr_mul(a, fr_r2());
}
MP_HD void fr_to_canonical(const fr& a, uint32_t* w) {
  fr one_raw = fr_zero();
  one_raw.v[0] = 1;
  fr c = fr_mul(a, one_raw);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = c.v[i];
}
MP_HD fr fr_from_u64(uint64_t x) {
  uint32_t w[8] = {(uint32_t)x, (uint32_t)(x >> 32), 0, 0, 0, 0, 0, 0};
  return fr_from_canonical(w);
}

// a^e for a small exponent (square-and-multiply, MSB first)
MP_HD fr fr_pow_u64(const fr& a, uint64_t e) {
  fr r = fr_one();
  for (int i = 63; i >= 0; i--) {
    r = fr_sqr(r);
    if ((e >> i) & 1) r = fr_mul(r, a);
  }
  return r;
}

}  // namespace mp
