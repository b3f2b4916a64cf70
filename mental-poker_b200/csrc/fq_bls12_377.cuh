// Base field F_q of BLS12-377 (377 bits), 12 x 32-bit limbs, Montgomery form with R = 2^384 -- the
// representation ark-ff 0.3 `Fp384` uses for `ark_bls12_377::Fq`, the curve the reference's own
// benchmark harness instantiates the protocol over (reference
// barnett-smart-card-protocol/examples/parameter_selection.rs:25-26; the trait is generic over
// `C: ProjectiveCurve`, src/discrete_log_cards/mod.rs:86).  Included INSTEAD of the Stark field by
// fq.cuh when MP_CURVE_BLS12_377 is defined, with the same function names and the same lazy-bound
// conventions, so ec.cuh / msm.cu compile unchanged against either field.
//
// What this prime offers:
//   * q == 1 (mod 2^46)  =>  -q^-1 == -1 (mod 2^32): the Montgomery quotient digit is a negation
//     and limb 0 of m*q costs nothing -- a reduction is 12 x 11 multiplies.
//   * q < 2^377 leaves 7 spare bits in 384 (2^384 / q > 152): values are kept LAZILY reduced as in
//     the Stark field: [k] = any 384-bit integer < k*q of the right residue.  fq_mul accepts
//     [x]*[y] with x*y <= 30 and returns [2] (< q * (1 + 30/152)).
//
// __host__ __device__ throughout: device code uses PTX carry chains (mad.lo.cc / madc.hi.cc pairs
// that ptxas fuses into IMAD.WIDE.U32.X), the host runs the same word-level algorithm on uint64
// arithmetic so that tests/test_host_bls12_377.py checks it with g++ before any GPU is involved.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MP_HD __host__ __device__ __forceinline__
#define MP_D __device__ __forceinline__
#else
#define MP_HD inline
#define MP_D inline
#endif

#define MP_CURVE_NAME "BLS12-377 G1"
#define MP_CURVE_A_IS_ZERO 1

namespace mp {

static constexpr int kFqLimbs = 12;
// bits the signed-digit recoding must cover: bit length of the group order (253) + 1 for the carry
static constexpr int kScalarBits = 254;

struct fq {
  uint32_t v[12];
};

MP_HD uint32_t fq_modulus_limb(int i) {
  switch (i) {
    case 0: return 0x00000001u;
    case 1: return 0x8508c000u;
    case 2: return 0x30000000u;
    case 3: return 0x170b5d44u;
    case 4: return 0xba094800u;
    case 5: return 0x1ef3622fu;
    case 6: return 0x00f5138fu;
    case 7: return 0x1a22d9f3u;
    case 8: return 0x6ca1493bu;
    case 9: return 0xc63b05c0u;
    case 10: return 0x17c510eau;
    default: return 0x01ae3a46u;
  }
}

MP_HD fq fq_zero() {
  fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = 0;
  return r;
}

// k * q for k <= 152 (fits 384 bits); folds to constants when k is a literal
MP_HD fq fq_kp(uint32_t k) {
  fq r;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    c += (uint64_t)k * fq_modulus_limb(i);
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
  return r;
}

// R mod q (Montgomery one) and R^2 mod q
MP_HD fq fq_one() {
  fq r;
  r.v[0] = 0xffffff68u; r.v[1] = 0x02cdffffu; r.v[2] = 0x7fffffb1u; r.v[3] = 0x51409f83u;
  r.v[4] = 0x8a7d3ff2u; r.v[5] = 0x9f7db3a9u; r.v[6] = 0x6e7c6305u; r.v[7] = 0x7b4e97b7u;
  r.v[8] = 0x803c84e8u; r.v[9] = 0x4cf495bfu; r.v[10] = 0xe2fdf49au; r.v[11] = 0x008d6661u;
  return r;
}
MP_HD fq fq_r2() {
  fq r;
  r.v[0] = 0x9400cd22u; r.v[1] = 0xb786686cu; r.v[2] = 0xb00431b1u; r.v[3] = 0x0329fcaau;
  r.v[4] = 0x62d6b46du; r.v[5] = 0x22a5f111u; r.v[6] = 0x827dc3acu; r.v[7] = 0xbfdf7d03u;
  r.v[8] = 0x41790bf9u; r.v[9] = 0x837e92f0u; r.v[10] = 0x1e914b88u; r.v[11] = 0x006dfccbu;
  return r;
}
// curve coefficient b = 1 in Montgomery form
MP_HD fq fq_curve_b() { return fq_one(); }

// ------------------------------------------------------------------------------------------
// word-level primitives
// ------------------------------------------------------------------------------------------
MP_HD fq fq_add(const fq& a, const fq& b) {
  fq r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %12, %24;\n\t"
      "addc.cc.u32 %1, %13, %25;\n\t"
      "addc.cc.u32 %2, %14, %26;\n\t"
      "addc.cc.u32 %3, %15, %27;\n\t"
      "addc.cc.u32 %4, %16, %28;\n\t"
      "addc.cc.u32 %5, %17, %29;\n\t"
      "addc.cc.u32 %6, %18, %30;\n\t"
      "addc.cc.u32 %7, %19, %31;\n\t"
      "addc.cc.u32 %8, %20, %32;\n\t"
      "addc.cc.u32 %9, %21, %33;\n\t"
      "addc.cc.u32 %10, %22, %34;\n\t"
      "addc.u32 %11, %23, %35;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7]), "=r"(r.v[8]), "=r"(r.v[9]), "=r"(r.v[10]), "=r"(r.v[11])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 12; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
#endif
  return r;
}

// r = a - b (384-bit wrap-around); *borrow = 1 if a < b
MP_HD fq fq_sub_raw(const fq& a, const fq& b, uint32_t* borrow) {
  fq r;
#ifdef __CUDA_ARCH__
  uint32_t bw;
  asm("sub.cc.u32 %0, %13, %25;\n\t"
      "subc.cc.u32 %1, %14, %26;\n\t"
      "subc.cc.u32 %2, %15, %27;\n\t"
      "subc.cc.u32 %3, %16, %28;\n\t"
      "subc.cc.u32 %4, %17, %29;\n\t"
      "subc.cc.u32 %5, %18, %30;\n\t"
      "subc.cc.u32 %6, %19, %31;\n\t"
      "subc.cc.u32 %7, %20, %32;\n\t"
      "subc.cc.u32 %8, %21, %33;\n\t"
      "subc.cc.u32 %9, %22, %34;\n\t"
      "subc.cc.u32 %10, %23, %35;\n\t"
      "subc.cc.u32 %11, %24, %36;\n\t"
      "subc.u32 %12, 0, 0;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
        "=r"(r.v[6]), "=r"(r.v[7]), "=r"(r.v[8]), "=r"(r.v[9]), "=r"(r.v[10]), "=r"(r.v[11]), "=r"(bw)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]),
        "r"(a.v[6]), "r"(a.v[7]), "r"(a.v[8]), "r"(a.v[9]), "r"(a.v[10]), "r"(a.v[11]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]),
        "r"(b.v[6]), "r"(b.v[7]), "r"(b.v[8]), "r"(b.v[9]), "r"(b.v[10]), "r"(b.v[11]));
  *borrow = bw & 1u;
#else
  int64_t c = 0;
  for (int i = 0; i < 12; i++) {
    c += (int64_t)a.v[i] - (int64_t)b.v[i];
    r.v[i] = (uint32_t)c;
    c >>= 32;  // arithmetic shift: 0 or -1
  }
  *borrow = (uint32_t)(c & 1);
#endif
  return r;
}

// r = a - k*q  (k <= 152; caller guarantees a >= k*q)
MP_HD fq fq_sub_kp(const fq& a, uint32_t k) {
  uint32_t bw;
  return fq_sub_raw(a, fq_kp(k), &bw);
}

// Any 384-bit value -> [2], same residue.  With t = the top word and 152 = floor(2^32 / (q_top + 1)),
// k = floor(152 * t / 2^32) satisfies k*q <= v and v - k*q < 1.34 q  (152 * q / 2^32 lies in
// (0.9978, 1) * 2^352, so k*q undershoots t * 2^352 by at most 0.0022 * 2^384 = 0.33 q).
MP_HD fq fq_reduce_weak(const fq& a) {
  uint32_t k = (uint32_t)(((uint64_t)a.v[11] * 152u) >> 32);
  return fq_sub_kp(a, k);
}

// a in [k] minus b in [kb]: returns a + kb*q - b  -> [k + kb]
MP_HD fq fq_sub(const fq& a, const fq& b, uint32_t kb) {
  uint32_t bw;
  fq t = fq_add(a, fq_kp(kb));
  return fq_sub_raw(t, b, &bw);
}

// canonical representative in [0, q) of any 384-bit value
MP_HD fq fq_reduce_full(const fq& a) {
  fq t = fq_reduce_weak(a);  // < 2q
  uint32_t bw;
  fq u = fq_sub_raw(t, fq_kp(1), &bw);
  fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.v[i] = bw ? t.v[i] : u.v[i];
  return r;
}

MP_HD bool fq_is_zero_raw(const fq& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) o |= a.v[i];
  return o == 0;
}

// a in [2]: is a == 0 (mod q)?  (a is 0 or q)
MP_HD bool fq_is_zero_mod_p_2(const fq& a) {
  uint32_t z = 0, e = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    z |= a.v[i];
    e |= a.v[i] ^ fq_modulus_limb(i);
  }
  return (z == 0) | (e == 0);
}

MP_HD bool fq_eq_raw(const fq& a, const fq& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// ------------------------------------------------------------------------------------------
// Carry-chain primitives.  Each has a PTX form (device) and a uint64 form (host) with identical
// semantics, so the index bookkeeping of the product and of the reduction below -- written once, in
// plain C++ over these primitives -- is what the host tests exercise.
// ------------------------------------------------------------------------------------------
// c[0..11] += {a0 .. a5} * b laid out as six adjacent 64-bit products; top += the carry out of c[11]
MP_HD void mp_mad6(uint32_t* c, uint32_t& top, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                   uint32_t a5, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %13, %19, %0;\n\t"
      "madc.hi.cc.u32 %1, %13, %19, %1;\n\t"
      "madc.lo.cc.u32 %2, %14, %19, %2;\n\t"
      "madc.hi.cc.u32 %3, %14, %19, %3;\n\t"
      "madc.lo.cc.u32 %4, %15, %19, %4;\n\t"
      "madc.hi.cc.u32 %5, %15, %19, %5;\n\t"
      "madc.lo.cc.u32 %6, %16, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %16, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %17, %19, %8;\n\t"
      "madc.hi.cc.u32 %9, %17, %19, %9;\n\t"
      "madc.lo.cc.u32 %10, %18, %19, %10;\n\t"
      "madc.hi.cc.u32 %11, %18, %19, %11;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]),
        "+r"(c[7]), "+r"(c[8]), "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(b));
#else
  const uint32_t a[6] = {a0, a1, a2, a3, a4, a5};
  uint64_t cy = 0;
  for (int k = 0; k < 6; k++) {
    uint64_t pr = (uint64_t)a[k] * b;
    cy += (uint64_t)c[2 * k] + (uint32_t)pr;
    c[2 * k] = (uint32_t)cy;
    cy >>= 32;
    cy += (uint64_t)c[2 * k + 1] + (uint32_t)(pr >> 32);
    c[2 * k + 1] = (uint32_t)cy;
    cy >>= 32;
  }
  top += (uint32_t)cy;
#endif
}
// c[0..9] += {a0 .. a4} * b as five adjacent 64-bit products; top += the carry out of c[9]
MP_HD void mp_mad5(uint32_t* c, uint32_t& top, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                   uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %11, %16, %0;\n\t"
      "madc.hi.cc.u32 %1, %11, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %12, %16, %2;\n\t"
      "madc.hi.cc.u32 %3, %12, %16, %3;\n\t"
      "madc.lo.cc.u32 %4, %13, %16, %4;\n\t"
      "madc.hi.cc.u32 %5, %13, %16, %5;\n\t"
      "madc.lo.cc.u32 %6, %14, %16, %6;\n\t"
      "madc.hi.cc.u32 %7, %14, %16, %7;\n\t"
      "madc.lo.cc.u32 %8, %15, %16, %8;\n\t"
      "madc.hi.cc.u32 %9, %15, %16, %9;\n\t"
      "addc.u32 %10, %10, 0;"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]),
        "+r"(c[7]), "+r"(c[8]), "+r"(c[9]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(b));
#else
  const uint32_t a[5] = {a0, a1, a2, a3, a4};
  uint64_t cy = 0;
  for (int k = 0; k < 5; k++) {
    uint64_t pr = (uint64_t)a[k] * b;
    cy += (uint64_t)c[2 * k] + (uint32_t)pr;
    c[2 * k] = (uint32_t)cy;
    cy >>= 32;
    cy += (uint64_t)c[2 * k + 1] + (uint32_t)(pr >> 32);
    c[2 * k + 1] = (uint32_t)cy;
    cy >>= 32;
  }
  top += (uint32_t)cy;
#endif
}

// c[0..9] += m * (q2 + q4*2^64 + ... + q10*2^256)  (even limbs of q above limb 0); top += carry out
MP_HD void mp_red_even(uint32_t* c, uint32_t& top, uint32_t m) {
  mp_mad5(c, top, fq_modulus_limb(2), fq_modulus_limb(4), fq_modulus_limb(6), fq_modulus_limb(8), fq_modulus_limb(10), m);
}
// c[0..11] += m * (q1 + q3*2^64 + ... + q11*2^320)  (the odd limbs of q); top += carry out
MP_HD void mp_red_odd(uint32_t* c, uint32_t& top, uint32_t m) {
  mp_mad6(c, top, fq_modulus_limb(1), fq_modulus_limb(3), fq_modulus_limb(5), fq_modulus_limb(7), fq_modulus_limb(9),
          fq_modulus_limb(11), m);
}

// One reduction round's word: w = t + o + p (the three accumulators' shares of word i).  Returns
// m = -w mod 2^32 and adds to `up` what moves into word i+1: the carries of the sum, plus one more when
// w != 0 mod 2^32, because w + m*q0 = w + m = 2^32 exactly (q0 = 1).
MP_HD uint32_t mp_fold(uint32_t t, uint32_t o, uint32_t p, uint32_t& up) {
  uint32_t m;
#ifdef __CUDA_ARCH__
  // volatile: the negation must stay hidden from the compiler -- seen as `0 - w` it is folded into the
  // multiplications as a negated operand, which the 64-bit IMAD.WIDE form does not have, and every product
  // of the reduction then costs an IMAD + an IMAD.HI (measured in SASS) instead of one IMAD.WIDE
  uint32_t w, hi;
  asm volatile("add.cc.u32 %0, %4, %5;\n\t"
               "addc.u32 %1, 0, 0;\n\t"
               "add.cc.u32 %0, %0, %6;\n\t"
               "addc.u32 %1, %1, 0;\n\t"
               "add.cc.u32 %2, %0, 0xffffffff;\n\t"  // carry out  <=>  w != 0
               "addc.u32 %1, %1, 0;\n\t"
               "sub.u32 %2, 0, %0;\n\t"
               "add.u32 %3, %3, %1;"
               : "=&r"(w), "=&r"(hi), "=&r"(m), "+r"(up)
               : "r"(t), "r"(o), "r"(p));
#else
  const uint64_t s = (uint64_t)t + o + p;
  const uint32_t w = (uint32_t)s;
  m = 0u - w;
  up += (uint32_t)(s >> 32) + (w != 0u ? 1u : 0u);
#endif
  return m;
}

// r[0..N) = x[0..N) + y[0..N), carry out dropped (callers know the bound)
template <int N>
MP_HD void mp_add_words(uint32_t* r, const uint32_t* x, const uint32_t* y) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    c += (uint64_t)x[i] + y[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
}
#ifdef __CUDA_ARCH__
template <>
__device__ __forceinline__ void mp_add_words<12>(uint32_t* r, const uint32_t* x, const uint32_t* y) {
  asm("add.cc.u32 %0, %12, %24;\n\t"
      "addc.cc.u32 %1, %13, %25;\n\t"
      "addc.cc.u32 %2, %14, %26;\n\t"
      "addc.cc.u32 %3, %15, %27;\n\t"
      "addc.cc.u32 %4, %16, %28;\n\t"
      "addc.cc.u32 %5, %17, %29;\n\t"
      "addc.cc.u32 %6, %18, %30;\n\t"
      "addc.cc.u32 %7, %19, %31;\n\t"
      "addc.cc.u32 %8, %20, %32;\n\t"
      "addc.cc.u32 %9, %21, %33;\n\t"
      "addc.cc.u32 %10, %22, %34;\n\t"
      "addc.u32 %11, %23, %35;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11])
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(x[8]),
        "r"(x[9]), "r"(x[10]), "r"(x[11]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]),
        "r"(y[6]), "r"(y[7]), "r"(y[8]), "r"(y[9]), "r"(y[10]), "r"(y[11]));
}
template <>
__device__ __forceinline__ void mp_add_words<23>(uint32_t* r, const uint32_t* x, const uint32_t* y) {
  asm("add.cc.u32 %0, %23, %46;\n\t"
      "addc.cc.u32 %1, %24, %47;\n\t"
      "addc.cc.u32 %2, %25, %48;\n\t"
      "addc.cc.u32 %3, %26, %49;\n\t"
      "addc.cc.u32 %4, %27, %50;\n\t"
      "addc.cc.u32 %5, %28, %51;\n\t"
      "addc.cc.u32 %6, %29, %52;\n\t"
      "addc.cc.u32 %7, %30, %53;\n\t"
      "addc.cc.u32 %8, %31, %54;\n\t"
      "addc.cc.u32 %9, %32, %55;\n\t"
      "addc.cc.u32 %10, %33, %56;\n\t"
      "addc.cc.u32 %11, %34, %57;\n\t"
      "addc.cc.u32 %12, %35, %58;\n\t"
      "addc.cc.u32 %13, %36, %59;\n\t"
      "addc.cc.u32 %14, %37, %60;\n\t"
      "addc.cc.u32 %15, %38, %61;\n\t"
      "addc.cc.u32 %16, %39, %62;\n\t"
      "addc.cc.u32 %17, %40, %63;\n\t"
      "addc.cc.u32 %18, %41, %64;\n\t"
      "addc.cc.u32 %19, %42, %65;\n\t"
      "addc.cc.u32 %20, %43, %66;\n\t"
      "addc.cc.u32 %21, %44, %67;\n\t"
      "addc.u32 %22, %45, %68;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22])
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(x[8]), "r"(x[9]), "r"(x[10]), "r"(x[11]), "r"(x[12]), "r"(x[13]), "r"(x[14]), "r"(x[15]),
        "r"(x[16]), "r"(x[17]), "r"(x[18]), "r"(x[19]), "r"(x[20]), "r"(x[21]), "r"(x[22]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]),
        "r"(y[8]), "r"(y[9]), "r"(y[10]), "r"(y[11]), "r"(y[12]), "r"(y[13]), "r"(y[14]), "r"(y[15]),
        "r"(y[16]), "r"(y[17]), "r"(y[18]), "r"(y[19]), "r"(y[20]), "r"(y[21]), "r"(y[22]));
}
#endif

// ------------------------------------------------------------------------------------------
// 12 x 12 schoolbook product, T[24] = a * b  (144 IMAD.WIDE on the device)
// ------------------------------------------------------------------------------------------
// ev[k] = word k of the sum of the products at even word offsets; od[k] = word k + 1 of the sum of the
// products at odd word offsets -- so the lo/hi halves of one product are adjacent in either array and
// a whole row is one carry chain.  The carry out of a chain lands in a word that so far holds only
// such carries (it is the top of the partial sum), so it is added without a further ripple.
MP_HD void fq_mul_wide(uint32_t* __restrict__ T, const fq& a, const fq& b) {
  uint32_t ev[26], od[24];
#pragma unroll
  for (int i = 0; i < 26; i++) ev[i] = 0;
#pragma unroll
  for (int i = 0; i < 24; i++) od[i] = 0;
#pragma unroll
  for (int i = 0; i < 12; i += 2) {
    // row i (even): even-j products at offset i + j (even) -> ev[i ..]; odd-j -> od[i ..]
    mp_mad6(ev + i, ev[i + 12], a.v[0], a.v[2], a.v[4], a.v[6], a.v[8], a.v[10], b.v[i]);
    mp_mad6(od + i, od[i + 12], a.v[1], a.v[3], a.v[5], a.v[7], a.v[9], a.v[11], b.v[i]);
    // row i + 1 (odd): even-j products at offset i + 1 + j (odd) -> od[i ..];
    //                  odd-j products at offset i + 1 + j (even) -> ev[i + 2 ..]
    mp_mad6(od + i, od[i + 12], a.v[0], a.v[2], a.v[4], a.v[6], a.v[8], a.v[10], b.v[i + 1]);
    mp_mad6(ev + i + 2, ev[i + 14], a.v[1], a.v[3], a.v[5], a.v[7], a.v[9], a.v[11], b.v[i + 1]);
  }
  // T = ev + (od << 32); the product is < 2^768, nothing is carried out of word 23
  T[0] = ev[0];
  mp_add_words<23>(T + 1, ev + 1, od);
}
MP_HD void fq_sqr_wide(uint32_t* T, const fq& a) { fq_mul_wide(T, a, a); }

// ------------------------------------------------------------------------------------------
// Montgomery reduction, word-serial:  r = (T + M*q) / 2^384  (12 x 11 IMAD.WIDE).  Round i clears word
// i with m = -word_i (mod 2^32) -- q == 1 (mod 2^32), so limb 0 of m*q only turns word i into a carry --
// and adds the other limbs of m*q at word offset i.  Every 64-bit product must sit on an even-aligned
// register pair for ptxas to keep mad.lo/mad.hi fused as IMAD.WIDE, so the accumulators are split by the
// PARITY OF THE WORD OFFSET, exactly as in fq_mul_wide:
//   T     products starting at an even word (pair T[2k], T[2k+1]),
//   od    od[k] = contribution to word k+1: products starting at an odd word,
//   pend  single carries waiting for word k (out of a fold or out of a chain).
// Even rounds send the even limbs of q to T and the odd limbs to od, odd rounds the other way round.
// The true word i is T[i] + od[i-1] + pend[i].  T is destroyed.  Output < T/2^384 + q.
// ------------------------------------------------------------------------------------------
MP_HD fq fq_mont_reduce(uint32_t* T) {
  uint32_t od[24], pend[26];
#pragma unroll
  for (int i = 0; i < 24; i++) od[i] = 0;
#pragma unroll
  for (int i = 0; i < 26; i++) pend[i] = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const uint32_t m = mp_fold(T[i], i > 0 ? od[i - 1] : 0u, pend[i], pend[i + 1]);
    if ((i & 1) == 0) {
      mp_red_even(T + i + 2, pend[i + 12], m);   // words i+2 .. i+11
      mp_red_odd(od + i, pend[i + 13], m);       // words i+1 .. i+12
    } else {
      mp_red_even(od + i + 1, pend[i + 12], m);  // words i+2 .. i+11 (odd-aligned pairs)
      mp_red_odd(T + i + 1, pend[i + 13], m);    // words i+1 .. i+12 (even-aligned pairs)
    }
  }
  // r = words 12 .. 23 of  T + (od << 32) + pend;  word 24 is empty by the bound
  fq r;
  uint32_t u[12];
  mp_add_words<12>(u, T + 12, od + 11);
  mp_add_words<12>(r.v, u, pend + 12);
  return r;
}

// a in [x], b in [y], x*y <= 30  ->  a*b/R in [2]
MP_HD fq fq_mul_inline(const fq& a, const fq& b) {
  uint32_t T[24];
  fq_mul_wide(T, a, b);
  return fq_mont_reduce(T);
}
#ifdef __CUDACC__
// ONE copy of the multiplication per kernel: fully unrolled it is ~1 300 instructions (20 KB); inlined ten
// times into a mixed addition the accumulate loop would be 200 KB of straight-line code and run out of the
// instruction cache (measured: the inlined form reached 0.86 G additions/s, a third of what its IMAD.WIDE
// count allows).  The call passes 24 + 12 words through the ABI -- under 5 % of the body.
// Operands and result travel BY VALUE.  The round-1 form took three pointers (fq* r, const fq* a, const fq* b);
// with it the NVVM optimiser merged the stack slot of a loop-carried accumulator with the slot of an unrelated
// product -- PTX of `z = fq_mul(p.ZZ, p.ZZZ); acc = fq_mul(acc, z);` read `call fq_mul_call(%rd27, %rd27, %rd27)`,
// i.e. acc = z * z -- which is what broke k_table_normalise and k_reduce_win on this build (DESIGN.md section 16,
// scripts/repro/).  Values have no address to merge.
static __device__ __noinline__ fq fq_mul_call(const fq a, const fq b) { return fq_mul_inline(a, b); }
#endif
MP_HD fq fq_mul(const fq& a, const fq& b) {
#ifdef __CUDA_ARCH__
  return fq_mul_call(a, b);
#else
  return fq_mul_inline(a, b);
#endif
}
MP_HD fq fq_sqr(const fq& a) { return fq_mul(a, a); }

MP_HD fq fq_to_mont(const fq& a) { return fq_mul(a, fq_r2()); }
MP_HD fq fq_from_mont(const fq& a) {
  uint32_t T[24];
#pragma unroll
  for (int i = 0; i < 12; i++) { T[i] = a.v[i]; T[12 + i] = 0; }
  return fq_reduce_full(fq_mont_reduce(T));
}

MP_HD fq fq_neg2(const fq& a) {  // a in [2] -> 2q - a in [2]
  uint32_t bw;
  return fq_sub_raw(fq_kp(2), a, &bw);
}

// a^(q-2) by 4-bit fixed windows; a in [2] (Montgomery), result [2].  a == 0 -> 0.
MP_HD fq fq_inv(const fq& a) {
  fq tab[16];
  tab[0] = fq_one();
  tab[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tab[i] = fq_mul(tab[i - 1], a);
  const uint32_t e[12] = {0xffffffffu, 0x8508bfffu, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                          0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
  fq r = fq_one();
#pragma unroll 1
  for (int w = 94; w >= 0; w--) {  // 377 bits = 95 nibbles
    r = fq_sqr(r); r = fq_sqr(r); r = fq_sqr(r); r = fq_sqr(r);
    uint32_t nib = (e[w >> 3] >> ((w & 7) * 4)) & 0xfu;
    r = fq_mul(r, tab[nib]);
  }
  return r;
}

}  // namespace mp
