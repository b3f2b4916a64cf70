// Square roots in the Stark base field (p - 1 = 2^192 * (2^59 + 17)): Tonelli-Shanks with the 192 powers
// root^(2^i) of a primitive 2^192-th root of unity precomputed.  Used by the wire-format decompression
// (wire.cu; reference path: ark-ff 0.3 `SquareRootField::sqrt` behind `CanonicalDeserialize`,
// reference src/lib.rs:45-71).  __host__ __device__ so that tests/host/host_shim.cpp can check the
// word-level algorithm against the oracle without a GPU.
#pragma once
#include "fq.cuh"

namespace mp {

#ifdef MP_CURVE_BLS12_377
// Second curve: q - 1 = 2^46 * t with t odd (331 bits).  The same windowed method with 2-bit windows (46 = 23 * 2):
// 23 digits, tables of a few hundred bytes; the cost is the exponentiation a^((t-1)/2) (330 squarings).
static constexpr int kTwoAdicity = 46;
static constexpr int kSqrtWin = 2;
// 5^t mod q (5 is the smallest non-residue): a primitive 2^46-th root of unity, Montgomery form (R = 2^384)
MP_HD fq fq_sqrt_root() {
  const uint32_t w[kFqLimbs] = {0x8bb191f2u, 0x68f876aau, 0xa6722e51u, 0x254e4780u, 0x1f8a0eafu, 0xa818ea19u,
                                0x1d8d5057u, 0x2c1a6dd3u, 0xa0df931bu, 0xcce5a0cbu, 0xc8cf8495u, 0x00ba7904u};
  fq r;
  for (int i = 0; i < kFqLimbs; i++) r.v[i] = w[i];
  return r;
}
// (q - 1) / 2, canonical: y is "the larger of (y, -y)" iff y > this
MP_HD uint32_t fq_half_limb(int i) {
  const uint32_t w[kFqLimbs] = {0x00000000u, 0x42846000u, 0x18000000u, 0x0b85aea2u, 0xdd04a400u, 0x8f79b117u,
                                0x807a89c7u, 0x8d116cf9u, 0x3650a49du, 0x631d82e0u, 0x0be28875u, 0x00d71d23u};
  return w[i];
}
// a^((t - 1) / 2): plain square-and-multiply over the 330-bit exponent (same control flow for every input)
MP_HD fq fq_sqrt_w(const fq& a) {
  const uint32_t e[11] = {0x00010a11u, 0xba886000u, 0x90002e16u, 0xc45f7412u, 0x271e3de6u, 0xb3e601eau,
                          0x92763445u, 0x0b80d942u, 0x21d58c76u, 0x748c2f8au, 0x0000035cu};
  fq w = a;                                   // bit 329 is set
  for (int i = 328; i >= 0; i--) {
    w = fq_sqr(w);
    if ((e[i >> 5] >> (i & 31)) & 1u) w = fq_mul(w, a);
  }
  return w;
}
#else
static constexpr int kTwoAdicity = 192;
static constexpr int kSqrtWin = 8;

// 3^t mod p, t = 2^59 + 17: a primitive 2^192-th root of unity, Montgomery form
MP_HD fq fq_sqrt_root() {
  fq r;
  r.v[0] = 0x64a2bdd8u; r.v[1] = 0x4106bccdu; r.v[2] = 0x31fe3be9u; r.v[3] = 0xaaada257u;
  r.v[4] = 0x60505574u; r.v[5] = 0x0a35c5beu; r.v[6] = 0xc47afc26u; r.v[7] = 0x07222e32u;
  return r;
}
// (p - 1) / 2 = 2^250 + 17 * 2^191, canonical: y is "the larger of (y, -y)" iff y > this
MP_HD uint32_t fq_half_limb(int i) {
  const uint32_t w[kFqLimbs] = {0, 0, 0, 0, 0, 0x80000000u, 0x00000008u, 0x04000000u};
  return w[i];
}
// a^((t - 1) / 2) = a^(2^58 + 8)
MP_HD fq fq_sqrt_w(const fq& a) {
  fq a8 = fq_sqr(fq_sqr(fq_sqr(a)));
  fq w = a8;
  for (int i = 3; i < 58; i++) w = fq_sqr(w);
  return fq_mul(w, a8);
}
#endif
// (the curve coefficient b * R mod p is fq_curve_b() of fq.cuh)

// T[i] = root^(2^i), fully reduced (191 dependent squarings, once per context)
MP_HD void fq_sqrt_table(fq* T) {
  fq cur = fq_reduce_full(fq_sqrt_root());
  for (int i = 0; i < kTwoAdicity; i++) {
    T[i] = cur;
    cur = fq_reduce_full(fq_sqr(cur));
  }
}

MP_HD bool fq_is_one(const fq& a) { return fq_eq_raw(fq_reduce_full(a), fq_one()); }

// a in [2] -> *ok = a is a square; returns a root (Montgomery, [2])
MP_HD fq fq_sqrt(const fq& a, const fq* T, bool* ok) {
  *ok = true;
  if (fq_is_zero_raw(fq_reduce_full(a))) return fq_zero();
  const fq w = fq_sqrt_w(a);
  fq x = fq_mul(a, w);
  fq b = fq_mul(x, w);
  int v = kTwoAdicity;
  while (!fq_is_one(b)) {
    int k = 0;
    fq t2 = b;
    do {
      t2 = fq_sqr(t2);
      k++;
    } while (k < v && !fq_is_one(t2));
    if (k >= v) { *ok = false; return fq_zero(); }
    x = fq_mul(x, T[kTwoAdicity - 1 - k]);
    b = fq_mul(b, T[kTwoAdicity - k]);   // T[s - k] = T[s - 1 - k]^2; k >= 1 so the index is <= s - 1
    v = k;
  }
  return x;
}

// ------------------------------------------------------------------------------------------
// Windowed form (what the GPU runs).  With zeta the root of unity above, x0 = a^((t+1)/2) and
// b = a^t = zeta^e (e even iff a is a square):  sqrt(a) = x0 * zeta^(-e/2).  The discrete log e is found
// eight bits at a time, least significant window first, without re-squaring:
//   b_j = b^(2^(8j)), j < 24                               one chain of 184 squarings
//   c_i = b_(23-i) * prod_{j<i} U[i-j][e_j]                = (zeta^(2^184))^(e_i), a 256th root of unity
//   e_i = lut[low 16 bits of c_i]                          (the 256 roots have distinct low 16 bits)
//   sqrt = x0 * prod_i V[i][e_i]
// with U[d][v] = zeta^(-v 2^(184-8d)) and V[i][v] = zeta^(-v 2^(8i-1)) precomputed (376 KB with the
// 64 KB lut).  242 squarings + 304 multiplications with the SAME control flow for every input, against
// ~4 600 data-dependent squarings for the loop above: on a GPU the lanes of a warp stay together.
// The result is checked (x^2 == a), which is also how non-residues are recognised.
// ------------------------------------------------------------------------------------------
static constexpr int kSqrtDigits = kTwoAdicity / kSqrtWin, kSqrtRadix = 1 << kSqrtWin;
static_assert(kSqrtDigits * kSqrtWin == kTwoAdicity, "window width must divide the two-adicity");
struct SqrtTables {
  const fq* U;         // [(d - 1) * 256 + v], d = 1 .. 23
  const fq* V;         // [i * 256 + v], i = 0 .. 23   (i = 0: even v only)
  const uint8_t* lut;  // 65536 entries
};
static constexpr size_t kSqrtUCount = (size_t)(kSqrtDigits - 1) * kSqrtRadix, kSqrtVCount = (size_t)kSqrtDigits * kSqrtRadix;

MP_HD uint32_t fq_sqrt_key(const fq& canonical_mont) { return canonical_mont.v[0] & 0xffffu; }

// prod of Tinv[shift + k] over the set bits k of v  (= zeta^(-v 2^shift)); shift may be -1 with v even
MP_HD fq fq_sqrt_pow_entry(const fq* Tinv, int shift, uint32_t v) {
  fq acc = fq_one();
  for (int k = 0; k < kSqrtWin; k++)
    if ((v >> k) & 1u) {
      const int idx = shift + k;
      if (idx >= 0) acc = fq_mul(acc, Tinv[idx]);
    }
  return fq_reduce_full(acc);
}
// Tinv[i] = zeta^(-2^i) from T[i] = zeta^(2^i): zeta^-1 = zeta^(2^192 - 1) = prod of all T[i]
MP_HD void fq_sqrt_inverse_table(const fq* T, fq* Tinv) {
  fq zinv = fq_one();
  for (int i = 0; i < kTwoAdicity; i++) zinv = fq_mul(zinv, T[i]);
  fq cur = fq_reduce_full(zinv);
  for (int i = 0; i < kTwoAdicity; i++) {
    Tinv[i] = cur;
    cur = fq_reduce_full(fq_sqr(cur));
  }
}
// entry `g` of the concatenated tables U | V | (256 lut writes): one call per g < kSqrtUCount + kSqrtVCount + 256
MP_HD void fq_sqrt_fill_entry(const fq* T, const fq* Tinv, size_t g, fq* U, fq* V, uint8_t* lut) {
  if (g < kSqrtUCount) {
    const int d = (int)(g / kSqrtRadix) + 1;
    U[g] = fq_sqrt_pow_entry(Tinv, kTwoAdicity - kSqrtWin - kSqrtWin * d, (uint32_t)(g % kSqrtRadix));
  } else if (g < kSqrtUCount + kSqrtVCount) {
    const size_t h = g - kSqrtUCount;
    const int i = (int)(h / kSqrtRadix);
    V[h] = fq_sqrt_pow_entry(Tinv, kSqrtWin * i - 1, (uint32_t)(h % kSqrtRadix));
  } else {
    const uint32_t j = (uint32_t)(g - kSqrtUCount - kSqrtVCount);  // omega^j, omega = zeta^(2^184)
    fq acc = fq_one();
    for (int k = 0; k < kSqrtWin; k++)
      if ((j >> k) & 1u) acc = fq_mul(acc, T[kTwoAdicity - kSqrtWin + k]);
    lut[fq_sqrt_key(fq_reduce_full(acc))] = (uint8_t)j;
  }
}

MP_HD fq fq_sqrt_win(const fq& a, const SqrtTables& tb, bool* ok) {
  *ok = true;
  const fq ar = fq_reduce_full(a);
  if (fq_is_zero_raw(ar)) return fq_zero();
  const fq w = fq_sqrt_w(ar);
  fq x = fq_mul(ar, w);
  fq bj[kSqrtDigits];
  bj[0] = fq_mul(x, w);
  for (int j = 1; j < kSqrtDigits; j++) {
    fq c = bj[j - 1];
    for (int k = 0; k < kSqrtWin; k++) c = fq_sqr(c);
    bj[j] = c;
  }
  uint8_t e[kSqrtDigits];
  for (int i = 0; i < kSqrtDigits; i++) {
    fq c = bj[kSqrtDigits - 1 - i];
    for (int j = 0; j < i; j++) c = fq_mul(c, tb.U[(size_t)(i - j - 1) * kSqrtRadix + e[j]]);
    e[i] = tb.lut[fq_sqrt_key(fq_reduce_full(c))];
  }
  for (int i = 0; i < kSqrtDigits; i++) x = fq_mul(x, tb.V[(size_t)i * kSqrtRadix + e[i]]);
  *ok = fq_eq_raw(fq_reduce_full(fq_sqr(x)), ar);
  return x;
}

}  // namespace mp
