// Short-Weierstrass group arithmetic in extended Jacobian ("XYZZ") coordinates, for the curve whose
// base field fq.cuh selects (the Stark curve, a = 1; or BLS12-377 G1, a = 0, under MP_CURVE_BLS12_377):
//   x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity <=> ZZ == 0 (stored as exact zero words).
// Replaces, for the shuffle hot path, the ark-ec 0.3 short-Weierstrass group the reference
// reaches through `C: ProjectiveCurve` (reference src/discrete_log_cards/mod.rs:86;
// SURVEY.md A2/A7).  Curve: y^2 = x^3 + x + b (a = 1).
//
// Formulas: EFD shortw/xyzz  madd-2008-s (8M+2S), add-2008-s (12M+2S), dbl-2008-s-1 (6M+4S
// with the a*ZZ^2 term kept because a = 1, mdbl-2008-s-1 for affine input).
// Bounds ([k] = value < k*p, see fq.cuh): every stored coordinate is [2].
#pragma once
#include "fq.cuh"

namespace mp {

struct affine {  // Montgomery-form coordinates, canonical or [2]; inf encoded by x == y == 0
  fq x, y;
};

struct xyzz {
  fq X, Y, ZZ, ZZZ;
};

MP_HD xyzz xyzz_identity() {
  xyzz r;
  r.X = fq_zero(); r.Y = fq_zero(); r.ZZ = fq_zero(); r.ZZZ = fq_zero();
  return r;
}
MP_HD bool xyzz_is_identity(const xyzz& p) { return fq_is_zero_raw(p.ZZ); }
// (0,0) is not on the curve (b != 0): used as the affine encoding of the identity everywhere.
MP_HD bool affine_is_identity(const affine& p) { return fq_is_zero_raw(p.x) & fq_is_zero_raw(p.y); }

MP_HD xyzz xyzz_from_affine(const affine& p) {
  xyzz r;
  if (affine_is_identity(p)) return xyzz_identity();
  r.X = p.x; r.Y = p.y; r.ZZ = fq_one(); r.ZZZ = fq_one();
  return r;
}

// 2 * (affine point).  mdbl-2008-s-1.
MP_HD xyzz xyzz_dbl_affine(const affine& p) {
  if (affine_is_identity(p)) return xyzz_identity();
  fq U = fq_add(p.y, p.y);                      // [4]
  fq Ured = fq_reduce_full(U);
  if (fq_is_zero_raw(Ured)) return xyzz_identity();  // y == 0: order-2 point (none on this curve)
  fq V = fq_sqr(U);                             // [2]
  fq W = fq_mul(U, V);                          // [2]
  fq S = fq_mul(p.x, V);                        // [2]
  fq XX = fq_sqr(p.x);                          // [2]
#if MP_CURVE_A_IS_ZERO
  fq M = fq_add(fq_add(XX, XX), XX);            // 3*XX -> [6]
#else
  fq M = fq_add(fq_add(XX, XX), fq_add(XX, fq_one()));  // 3*XX + a, a = 1 -> [8]
#endif
  M = fq_reduce_weak(M);                        // [2]
  fq X3 = fq_sub(fq_sqr(M), fq_add(S, S), 4);   // [2] + 4p - [4] -> [6]
  X3 = fq_reduce_weak(X3);                      // [2]
  fq Y3 = fq_sub(fq_mul(M, fq_sub(S, X3, 2)), fq_mul(W, p.y), 2);  // [4]
  xyzz r;
  r.X = X3; r.Y = fq_reduce_weak(Y3); r.ZZ = V; r.ZZZ = W;
  return r;
}

// 2 * P.  dbl-2008-s-1 with a = 1.
MP_HD xyzz xyzz_dbl(const xyzz& p) {
  if (xyzz_is_identity(p)) return p;
  fq U = fq_add(p.Y, p.Y);                      // [4]
  fq V = fq_sqr(U);                             // [2]
  fq W = fq_mul(U, V);                          // [2]
  fq S = fq_mul(p.X, V);                        // [2]
  fq XX = fq_sqr(p.X);                          // [2]
#if MP_CURVE_A_IS_ZERO
  fq M = fq_add(fq_add(XX, XX), XX);            // [6]  (a = 0: no ZZ^2 term)
#else
  fq ZZ2 = fq_sqr(p.ZZ);                        // [2]  (a * ZZ^2, a = 1)
  fq M = fq_add(fq_add(XX, XX), fq_add(XX, ZZ2));  // [8]
#endif
  M = fq_reduce_weak(M);
  fq X3 = fq_reduce_weak(fq_sub(fq_sqr(M), fq_add(S, S), 4));
  fq Y3 = fq_sub(fq_mul(M, fq_sub(S, X3, 2)), fq_mul(W, p.Y), 2);
  xyzz r;
  r.X = X3; r.Y = fq_reduce_weak(Y3);
  r.ZZ = fq_mul(V, p.ZZ); r.ZZZ = fq_mul(W, p.ZZZ);
  // Y == 0 (mod p) would give ZZ3 == 0 (mod p) but not exact zero words: normalise.
  if (fq_is_zero_mod_p_2(r.ZZ)) return xyzz_identity();
  return r;
}

// acc += q (q affine, Montgomery form).  madd-2008-s; complete (handles O, P+P, P-P).
MP_HD void xyzz_madd(xyzz& acc, const affine& q) {
  if (affine_is_identity(q)) return;
  if (xyzz_is_identity(acc)) {
    acc.X = q.x; acc.Y = q.y; acc.ZZ = fq_one(); acc.ZZZ = fq_one();
    return;
  }
  fq U2 = fq_mul(q.x, acc.ZZ);                  // [2]
  fq S2 = fq_mul(q.y, acc.ZZZ);                 // [2]
  fq P = fq_sub(U2, acc.X, 2);                  // [4]
  fq R = fq_sub(S2, acc.Y, 2);                  // [4]
  fq PP = fq_sqr(P);                            // [2]
  fq ZZ3 = fq_mul(acc.ZZ, PP);                  // [2]
  if (fq_is_zero_mod_p_2(ZZ3)) {                // P == 0 (mod p): same x
    if (fq_is_zero_raw(fq_reduce_full(R))) acc = xyzz_dbl_affine(q);
    else acc = xyzz_identity();
    return;
  }
  fq PPP = fq_mul(P, PP);                       // [2]
  fq Q = fq_mul(acc.X, PP);                     // [2]
  fq X3 = fq_sub(fq_sqr(R), fq_add(PPP, fq_add(Q, Q)), 6);  // [2] + 6p - [6] -> [8]
  X3 = fq_reduce_weak(X3);                      // [2]
  fq Y3 = fq_sub(fq_mul(R, fq_sub(Q, X3, 2)), fq_mul(acc.Y, PPP), 2);  // [4]
  acc.X = X3;
  acc.Y = fq_reduce_weak(Y3);
  acc.ZZ = ZZ3;
  acc.ZZZ = fq_mul(acc.ZZZ, PPP);
}

// acc += q, both XYZZ.  add-2008-s; complete.
MP_HD void xyzz_add(xyzz& acc, const xyzz& q) {
  if (xyzz_is_identity(q)) return;
  if (xyzz_is_identity(acc)) { acc = q; return; }
  fq U1 = fq_mul(acc.X, q.ZZ);
  fq U2 = fq_mul(q.X, acc.ZZ);
  fq S1 = fq_mul(acc.Y, q.ZZZ);
  fq S2 = fq_mul(q.Y, acc.ZZZ);
  fq P = fq_sub(U2, U1, 2);                     // [4]
  fq R = fq_sub(S2, S1, 2);                     // [4]
  fq PP = fq_sqr(P);
  fq ZZ3 = fq_mul(fq_mul(acc.ZZ, q.ZZ), PP);
  if (fq_is_zero_mod_p_2(ZZ3)) {
    if (fq_is_zero_raw(fq_reduce_full(R))) acc = xyzz_dbl(acc);
    else acc = xyzz_identity();
    return;
  }
  fq PPP = fq_mul(P, PP);
  fq Q = fq_mul(U1, PP);
  fq X3 = fq_reduce_weak(fq_sub(fq_sqr(R), fq_add(PPP, fq_add(Q, Q)), 6));
  fq Y3 = fq_sub(fq_mul(R, fq_sub(Q, X3, 2)), fq_mul(S1, PPP), 2);
  acc.X = X3;
  acc.Y = fq_reduce_weak(Y3);
  acc.ZZ = ZZ3;
  acc.ZZZ = fq_mul(fq_mul(acc.ZZZ, q.ZZZ), PPP);
}

MP_HD affine affine_neg(const affine& p) {
  affine r;
  r.x = p.x;
  if (affine_is_identity(p)) { r.y = p.y; return r; }
  // y in [2] -> 2p - y; canonical nonzero y stays nonzero.  (y == 0 never on this curve.)
  r.y = fq_neg2(p.y);
  return r;
}

MP_HD xyzz xyzz_neg(const xyzz& p) {
  xyzz r = p;
  if (!xyzz_is_identity(p)) r.Y = fq_neg2(p.Y);
  return r;
}

// XYZZ -> affine Montgomery, canonical (fully reduced) coordinates; identity -> (0,0).
MP_HD affine xyzz_to_affine(const xyzz& p) {
  affine r;
  if (xyzz_is_identity(p)) { r.x = fq_zero(); r.y = fq_zero(); return r; }
  fq i = fq_inv(fq_mul(p.ZZ, p.ZZZ));
  fq izz = fq_mul(i, p.ZZZ);
  fq izzz = fq_mul(i, p.ZZ);
  r.x = fq_reduce_full(fq_mul(p.X, izz));
  r.y = fq_reduce_full(fq_mul(p.Y, izzz));
  return r;
}

// canonical little-endian x || y (non-Montgomery, 2 * kFqLimbs words) <-> affine Montgomery
MP_HD affine affine_from_canonical(const uint32_t* w) {
  affine r;
  fq x, y;
#pragma unroll
  for (int i = 0; i < kFqLimbs; i++) { x.v[i] = w[i]; y.v[i] = w[kFqLimbs + i]; }
  if (fq_is_zero_raw(x) & fq_is_zero_raw(y)) { r.x = x; r.y = y; return r; }
  r.x = fq_reduce_full(fq_to_mont(x));
  r.y = fq_reduce_full(fq_to_mont(y));
  return r;
}
MP_HD void affine_to_canonical(const affine& p, uint32_t* w) {
  fq x = fq_from_mont(p.x), y = fq_from_mont(p.y);
#pragma unroll
  for (int i = 0; i < kFqLimbs; i++) { w[i] = x.v[i]; w[kFqLimbs + i] = y.v[i]; }
}

// y^2 == x^3 + a*x + b ?  (Montgomery inputs)
MP_HD bool affine_on_curve(const affine& p) {
  if (affine_is_identity(p)) return true;
  const fq bm = fq_curve_b();  // b * R mod p
  fq lhs = fq_sqr(p.y);
#if MP_CURVE_A_IS_ZERO
  fq rhs = fq_add(fq_mul(fq_sqr(p.x), p.x), bm);               // [2]+[1]
#else
  fq rhs = fq_add(fq_add(fq_mul(fq_sqr(p.x), p.x), p.x), bm);  // [2]+[2]+[1]
#endif
  return fq_eq_raw(fq_reduce_full(lhs), fq_reduce_full(rhs));
}

}  // namespace mp
