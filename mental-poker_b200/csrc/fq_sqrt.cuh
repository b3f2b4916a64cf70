// Square roots in the Stark base field (p - 1 = 2^192 * (2^59 + 17)): Tonelli-Shanks with the 192 powers
// root^(2^i) of a primitive 2^192-th root of unity precomputed.  Used by the wire-format decompression
// (wire.cu; reference path: ark-ff 0.3 `SquareRootField::sqrt` behind `CanonicalDeserialize`,
// reference src/lib.rs:45-71).  __host__ __device__ so that tests/host/host_shim.cpp can check the
// word-level algorithm against the oracle without a GPU.
#pragma once
#include "fq.cuh"

namespace mp {

static constexpr int kTwoAdicity = 192;

// 3^t mod p, t = 2^59 + 17: a primitive 2^192-th root of unity, Montgomery form
MP_HD fq fq_sqrt_root() {
  fq r;
  r.v[0] = 0x64a2bdd8u; r.v[1] = 0x4106bccdu; r.v[2] = 0x31fe3be9u; r.v[3] = 0xaaada257u;
  r.v[4] = 0x60505574u; r.v[5] = 0x0a35c5beu; r.v[6] = 0xc47afc26u; r.v[7] = 0x07222e32u;
  return r;
}
MP_HD fq fq_curve_b() {  // b * R mod p
  fq bm;
  bm.v[0] = 0xb59a21cau; bm.v[1] = 0x359ddd67u; bm.v[2] = 0x7aab9006u; bm.v[3] = 0x6725f223u;
  bm.v[4] = 0x2a41f947u; bm.v[5] = 0xab8a1e00u; bm.v[6] = 0x1774247fu; bm.v[7] = 0x01393165u;
  return bm;
}

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
  // w = a^(2^58 + 8)
  fq a8 = fq_sqr(fq_sqr(fq_sqr(a)));
  fq w = a8;
  for (int i = 3; i < 58; i++) w = fq_sqr(w);
  w = fq_mul(w, a8);
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
    b = fq_mul(b, T[kTwoAdicity - k]);   // T[192 - k] = T[191 - k]^2; k >= 1 so the index is <= 191
    v = k;
  }
  return x;
}

}  // namespace mp
