/* ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * Generic 256-bit prime-field arithmetic in Montgomery form on 4 x 64-bit limbs, restating
 * the algorithm of ark-ff 0.3 `Fp256<P>` (reference crate dependency,
 * barnett-smart-card-protocol/Cargo.toml:12): CIOS Montgomery multiplication
 * with R = 2^256, little-endian limbs, canonical 32-byte little-endian I/O (`ToBytes`).
 * Instantiated for the Stark base field F_p and scalar field F_n (constants SURVEY.md A7,
 * derived at start-up from the moduli and cross-checked in tests/test_oracle_c.py).
 *
 * PARITY UNPINNED against the upstream Rust crates (absent, unbuildable here); pinned against
 * the Python big-int oracle (oracle/py) and the committed golden fixtures.
 */
#ifndef ORACLE_FIELD_H
#define ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
  uint64_t m[4];   /* modulus */
  uint64_t inv;    /* -m^-1 mod 2^64 */
  fe one;          /* R mod m */
  fe r2;           /* R^2 mod m */
  uint64_t m2[4];  /* m - 2 (Fermat inversion exponent) */
} field_t;

static inline int limbs_geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline uint64_t limbs_sub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
static inline uint64_t limbs_add(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t carry = 0;
  for (int i = 0; i < 4; i++) {
    u128 s = (u128)a[i] + b[i] + carry;
    r[i] = (uint64_t)s;
    carry = (uint64_t)(s >> 64);
  }
  return carry;
}

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return memcmp(a, b, sizeof(fe)) == 0; }
static inline void fe_set_zero(fe* a) { memset(a, 0, sizeof(fe)); }

static inline void fe_add(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[4];
  uint64_t c = limbs_add(t, a->l, b->l);
  if (c || limbs_geq(t, F->m)) limbs_sub(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[4];
  if (limbs_sub(t, a->l, b->l)) limbs_add(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_neg(fe* r, const fe* a, const field_t* F) {
  if (fe_is_zero(a)) { fe_set_zero(r); return; }
  limbs_sub(r->l, F->m, a->l);
}
static inline void fe_dbl(fe* r, const fe* a, const field_t* F) { fe_add(r, a, a, F); }

/* CIOS Montgomery product (both moduli are < 2^252 so the running value fits 5 limbs) */
static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* F) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)t[j] + (u128)a->l[j] * b->l[i];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t q = t[0] * F->inv;
    c = ((u128)t[0] + (u128)q * F->m[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)t[j] + (u128)q * F->m[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || limbs_geq(t, F->m)) limbs_sub(t, t, F->m);
  memcpy(r->l, t, 32);
}
static inline void fe_sqr(fe* r, const fe* a, const field_t* F) { fe_mul(r, a, a, F); }

/* a^e, e = 4 little-endian limbs, MSB-first square-and-multiply */
static inline void fe_pow(fe* r, const fe* a, const uint64_t* e, const field_t* F) {
  fe acc = F->one;
  for (int i = 255; i >= 0; i--) {
    fe_sqr(&acc, &acc, F);
    if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a, F);
  }
  *r = acc;
}
static inline void fe_inv(fe* r, const fe* a, const field_t* F) { fe_pow(r, a, F->m2, F); }

/* canonical little-endian bytes <-> Montgomery */
static inline void fe_from_raw(fe* r, const uint64_t* canon, const field_t* F) {
  fe t;
  memcpy(t.l, canon, 32);
  fe_mul(r, &t, &F->r2, F);
}
static inline void fe_to_raw(uint64_t* canon, const fe* a, const field_t* F) {
  fe one_raw = {{1, 0, 0, 0}}, t;
  fe_mul(&t, a, &one_raw, F);
  memcpy(canon, t.l, 32);
}
static inline void fe_from_bytes(fe* r, const uint8_t* b, const field_t* F) {
  uint64_t c[4];
  memcpy(c, b, 32); /* little-endian host */
  /* reduce inputs >= m (API robustness; canonical inputs never need it) */
  while (limbs_geq(c, F->m)) limbs_sub(c, c, F->m);
  fe_from_raw(r, c, F);
}
static inline void fe_to_bytes(uint8_t* b, const fe* a, const field_t* F) {
  uint64_t c[4];
  fe_to_raw(c, a, F);
  memcpy(b, c, 32);
}
static inline void fe_from_u64(fe* r, uint64_t v, const field_t* F) {
  uint64_t c[4] = {v, 0, 0, 0};
  fe_from_raw(r, c, F);
}

/* derive inv, one, r2, m2 from the modulus */
static inline void field_init(field_t* F, const uint64_t* modulus) {
  memcpy(F->m, modulus, 32);
  uint64_t inv = 1;
  for (int i = 0; i < 63; i++) { inv *= inv; inv *= modulus[0]; } /* m^(2^63-1) = m^-1 mod 2^64 */
  F->inv = (uint64_t)0 - inv;
  /* R mod m by 256 modular doublings of 1, R^2 by 256 more */
  uint64_t t[4] = {1, 0, 0, 0};
  for (int i = 0; i < 512; i++) {
    uint64_t c = limbs_add(t, t, t);
    if (c || limbs_geq(t, F->m)) limbs_sub(t, t, F->m);
    if (i == 255) memcpy(F->one.l, t, 32);
  }
  memcpy(F->r2.l, t, 32);
  uint64_t two[4] = {2, 0, 0, 0};
  limbs_sub(F->m2, F->m, two);
}
#endif
