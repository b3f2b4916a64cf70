/* ORACLE (test infrastructure only).
 *
 * Stark-curve group arithmetic the way the reference's CPU path performs it through
 * ark-ec 0.3 (Cargo.toml:11): Jacobian projective coordinates, `mul` =
 * MSB-first double-and-add with mixed addition, `VariableBaseMSM::multi_scalar_mul` =
 * unsigned c-bit windows with c = 3 (n < 32) else ln_without_floats(n) + 2.
 * Curve y^2 = x^3 + x + b over F_p.  Outputs are affine + canonical, so the
 * choice of projective formulas cannot influence any result.  PARITY UNPINNED vs upstream.
 */
#ifndef ORACLE_CURVE_H
#define ORACLE_CURVE_H
#include <stdlib.h>

#include "field.h"

typedef struct { fe x, y; int inf; } aff;   /* coordinates in Montgomery form */
typedef struct { fe X, Y, Z; } jac;         /* identity <=> Z == 0 */

extern field_t FQ, FR;
extern fe CURVE_B; /* Montgomery form; a = 1 */
void oracle_init(void);

static inline void jac_set_inf(jac* p) { p->X = FQ.one; p->Y = FQ.one; fe_set_zero(&p->Z); }
static inline int jac_is_inf(const jac* p) { return fe_is_zero(&p->Z); }
static inline void jac_from_aff(jac* r, const aff* p) {
  if (p->inf) { jac_set_inf(r); return; }
  r->X = p->x; r->Y = p->y; r->Z = FQ.one;
}

/* dbl-2007-bl (general a; a = 1 here) */
static inline void jac_dbl(jac* r, const jac* p) {
  if (jac_is_inf(p)) { *r = *p; return; }
  fe XX, YY, YYYY, ZZ, S, M, T, t;
  fe_sqr(&XX, &p->X, &FQ);
  fe_sqr(&YY, &p->Y, &FQ);
  fe_sqr(&YYYY, &YY, &FQ);
  fe_sqr(&ZZ, &p->Z, &FQ);
  fe_add(&S, &p->X, &YY, &FQ);
  fe_sqr(&S, &S, &FQ);
  fe_sub(&S, &S, &XX, &FQ);
  fe_sub(&S, &S, &YYYY, &FQ);
  fe_dbl(&S, &S, &FQ);
  fe_dbl(&M, &XX, &FQ);
  fe_add(&M, &M, &XX, &FQ);
  fe_sqr(&t, &ZZ, &FQ); /* a * ZZ^2, a = 1 */
  fe_add(&M, &M, &t, &FQ);
  fe_sqr(&T, &M, &FQ);
  fe_sub(&T, &T, &S, &FQ);
  fe_sub(&T, &T, &S, &FQ);
  fe Z3;
  fe_add(&Z3, &p->Y, &p->Z, &FQ);
  fe_sqr(&Z3, &Z3, &FQ);
  fe_sub(&Z3, &Z3, &YY, &FQ);
  fe_sub(&Z3, &Z3, &ZZ, &FQ);
  fe Y3;
  fe_sub(&Y3, &S, &T, &FQ);
  fe_mul(&Y3, &Y3, &M, &FQ);
  fe_dbl(&YYYY, &YYYY, &FQ);
  fe_dbl(&YYYY, &YYYY, &FQ);
  fe_dbl(&YYYY, &YYYY, &FQ);
  fe_sub(&Y3, &Y3, &YYYY, &FQ);
  r->X = T; r->Y = Y3; r->Z = Z3;
}

/* madd-2007-bl, complete */
static inline void jac_add_mixed(jac* r, const jac* p, const aff* q) {
  if (q->inf) { *r = *p; return; }
  if (jac_is_inf(p)) { jac_from_aff(r, q); return; }
  fe Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
  fe_sqr(&Z1Z1, &p->Z, &FQ);
  fe_mul(&U2, &q->x, &Z1Z1, &FQ);
  fe_mul(&S2, &q->y, &p->Z, &FQ);
  fe_mul(&S2, &S2, &Z1Z1, &FQ);
  if (fe_eq(&U2, &p->X)) {
    if (fe_eq(&S2, &p->Y)) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fe_sub(&H, &U2, &p->X, &FQ);
  fe_sqr(&HH, &H, &FQ);
  fe_dbl(&I, &HH, &FQ);
  fe_dbl(&I, &I, &FQ);
  fe_mul(&J, &H, &I, &FQ);
  fe_sub(&rr, &S2, &p->Y, &FQ);
  fe_dbl(&rr, &rr, &FQ);
  fe_mul(&V, &p->X, &I, &FQ);
  fe X3, Y3, Z3;
  fe_sqr(&X3, &rr, &FQ);
  fe_sub(&X3, &X3, &J, &FQ);
  fe_sub(&X3, &X3, &V, &FQ);
  fe_sub(&X3, &X3, &V, &FQ);
  fe_sub(&Y3, &V, &X3, &FQ);
  fe_mul(&Y3, &Y3, &rr, &FQ);
  fe_mul(&t, &p->Y, &J, &FQ);
  fe_dbl(&t, &t, &FQ);
  fe_sub(&Y3, &Y3, &t, &FQ);
  fe_add(&Z3, &p->Z, &H, &FQ);
  fe_sqr(&Z3, &Z3, &FQ);
  fe_sub(&Z3, &Z3, &Z1Z1, &FQ);
  fe_sub(&Z3, &Z3, &HH, &FQ);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}

/* add-2007-bl, complete */
static inline void jac_add(jac* r, const jac* p, const jac* q) {
  if (jac_is_inf(q)) { *r = *p; return; }
  if (jac_is_inf(p)) { *r = *q; return; }
  fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
  fe_sqr(&Z1Z1, &p->Z, &FQ);
  fe_sqr(&Z2Z2, &q->Z, &FQ);
  fe_mul(&U1, &p->X, &Z2Z2, &FQ);
  fe_mul(&U2, &q->X, &Z1Z1, &FQ);
  fe_mul(&S1, &p->Y, &q->Z, &FQ);
  fe_mul(&S1, &S1, &Z2Z2, &FQ);
  fe_mul(&S2, &q->Y, &p->Z, &FQ);
  fe_mul(&S2, &S2, &Z1Z1, &FQ);
  if (fe_eq(&U1, &U2)) {
    if (fe_eq(&S1, &S2)) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fe_sub(&H, &U2, &U1, &FQ);
  fe_dbl(&I, &H, &FQ);
  fe_sqr(&I, &I, &FQ);
  fe_mul(&J, &H, &I, &FQ);
  fe_sub(&rr, &S2, &S1, &FQ);
  fe_dbl(&rr, &rr, &FQ);
  fe_mul(&V, &U1, &I, &FQ);
  fe X3, Y3, Z3;
  fe_sqr(&X3, &rr, &FQ);
  fe_sub(&X3, &X3, &J, &FQ);
  fe_sub(&X3, &X3, &V, &FQ);
  fe_sub(&X3, &X3, &V, &FQ);
  fe_sub(&Y3, &V, &X3, &FQ);
  fe_mul(&Y3, &Y3, &rr, &FQ);
  fe_mul(&t, &S1, &J, &FQ);
  fe_dbl(&t, &t, &FQ);
  fe_sub(&Y3, &Y3, &t, &FQ);
  fe_add(&Z3, &p->Z, &q->Z, &FQ);
  fe_sqr(&Z3, &Z3, &FQ);
  fe_sub(&Z3, &Z3, &Z1Z1, &FQ);
  fe_sub(&Z3, &Z3, &Z2Z2, &FQ);
  fe_mul(&Z3, &Z3, &H, &FQ);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}

static inline void jac_to_aff(aff* r, const jac* p) {
  if (jac_is_inf(p)) { fe_set_zero(&r->x); r->y = FQ.one; r->inf = 1; return; }
  fe zi, zi2, zi3;
  fe_inv(&zi, &p->Z, &FQ);
  fe_sqr(&zi2, &zi, &FQ);
  fe_mul(&zi3, &zi2, &zi, &FQ);
  fe_mul(&r->x, &p->X, &zi2, &FQ);
  fe_mul(&r->y, &p->Y, &zi3, &FQ);
  r->inf = 0;
}
static inline void aff_neg(aff* r, const aff* p) {
  *r = *p;
  if (!p->inf) fe_neg(&r->y, &p->y, &FQ);
}
static inline int aff_eq(const aff* a, const aff* b) {
  if (a->inf || b->inf) return a->inf && b->inf;
  return fe_eq(&a->x, &b->x) && fe_eq(&a->y, &b->y);
}
static inline int aff_on_curve(const aff* p) {
  if (p->inf) return 1;
  fe l, r;
  fe_sqr(&l, &p->y, &FQ);
  fe_sqr(&r, &p->x, &FQ);
  fe_mul(&r, &r, &p->x, &FQ);
  fe_add(&r, &r, &p->x, &FQ);
  fe_add(&r, &r, &CURVE_B, &FQ);
  return fe_eq(&l, &r);
}

/* 64-byte C-ABI layout: x || y canonical LE; all-zero = identity */
static inline void aff_from_bytes64(aff* r, const uint8_t* b) {
  int z = 1;
  for (int i = 0; i < 64; i++) if (b[i]) { z = 0; break; }
  if (z) { fe_set_zero(&r->x); r->y = FQ.one; r->inf = 1; return; }
  fe_from_bytes(&r->x, b, &FQ);
  fe_from_bytes(&r->y, b + 32, &FQ);
  r->inf = 0;
}
static inline void aff_to_bytes64(uint8_t* b, const aff* p) {
  if (p->inf) { memset(b, 0, 64); return; }
  fe_to_bytes(b, &p->x, &FQ);
  fe_to_bytes(b + 32, &p->y, &FQ);
}
/* ark-ec 0.3 `GroupAffine::write`: x || y || infinity flag, identity = (0, 1, true) */
static inline void aff_to_bytes65(uint8_t* b, const aff* p) {
  if (p->inf) { memset(b, 0, 65); b[32] = 1; b[64] = 1; return; }
  fe_to_bytes(b, &p->x, &FQ);
  fe_to_bytes(b + 32, &p->y, &FQ);
  b[64] = 0;
}

/* ark-ec 0.3 `AffineCurve::mul`: MSB-first double-and-add over the canonical scalar bits,
 * leading zeros skipped, mixed addition. */
static inline void aff_mul_raw(jac* r, const aff* p, const uint64_t* k) {
  jac acc;
  jac_set_inf(&acc);
  int top = 255;
  while (top >= 0 && !((k[top >> 6] >> (top & 63)) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    jac_dbl(&acc, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) jac_add_mixed(&acc, &acc, p);
  }
  *r = acc;
}
static inline void aff_mul(jac* r, const aff* p, const fe* k_mont) {
  uint64_t k[4];
  fe_to_raw(k, k_mont, &FR);
  aff_mul_raw(r, p, k);
}
#endif
