/* ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * Plain-C restatement of the CPU group arithmetic the reference reaches for its SECOND curve
 * instantiation, `DLCards<ark_bls12_377::G1Projective>` (reference
 * barnett-smart-card-protocol/examples/parameter_selection.rs:25-29): ark-ff 0.3 `Fp384` (6 x u64 CIOS
 * Montgomery, R = 2^384), ark-ec 0.3 short-Weierstrass Jacobian arithmetic (a = 0), per-term
 * double-and-add (`AffineCurve::mul`) and `VariableBaseMSM::multi_scalar_mul` (SURVEY.md A1/A2,
 * restated from recall: PARITY UNPINNED against the upstream crates, pinned against the Python big-int
 * oracle oracle/py/bls12_377.py and tests/golden/bls12_377_vectors.json in tests/test_oracle_bls12_377.py).
 * Doubles as the CPU baseline ("port") of bench.py's BLS12-377 section.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fe;
typedef struct { fe x, y; int inf; } aff;
typedef struct { fe X, Y, Z; } jac;

static const uint64_t QM[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull,
                               0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x01ae3a4617c510eaull};
static const uint64_t RM[4] = {0x0a11800000000001ull, 0x59aa76fed0000001ull, 0x60b44d1e5c37b001ull, 0x12ab655e9a2ca556ull};
static const uint64_t QINV = 0x8508bfffffffffffull; /* -q^-1 mod 2^64 (q = 1 mod 2^46) */
static fe ONE, R2;
static int g_threads = 1, g_init = 0;

static int geq6(const uint64_t* a, const uint64_t* b) {
  for (int i = 5; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static uint64_t sub6(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  return borrow;
}
static uint64_t add6(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t carry = 0;
  for (int i = 0; i < 6; i++) {
    u128 s = (u128)a[i] + b[i] + carry;
    r[i] = (uint64_t)s;
    carry = (uint64_t)(s >> 64);
  }
  return carry;
}
static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3] | a->l[4] | a->l[5]) == 0; }
static void fe_add(fe* r, const fe* a, const fe* b) {
  uint64_t t[6];
  add6(t, a->l, b->l); /* q < 2^377: no carry out */
  if (geq6(t, QM)) sub6(t, t, QM);
  memcpy(r->l, t, 48);
}
static void fe_sub(fe* r, const fe* a, const fe* b) {
  uint64_t t[6];
  if (sub6(t, a->l, b->l)) add6(t, t, QM);
  memcpy(r->l, t, 48);
}
static void fe_dbl(fe* r, const fe* a) { fe_add(r, a, a); }
/* CIOS Montgomery product (ark-ff 0.3 `mul_assign` without the no-carry shortcut) */
static void fe_mul(fe* r, const fe* a, const fe* b) {
  uint64_t t[8] = {0};
  for (int i = 0; i < 6; i++) {
    u128 c = 0;
    for (int j = 0; j < 6; j++) {
      c += (u128)t[j] + (u128)a->l[j] * b->l[i];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[6] = (uint64_t)c;
    t[7] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * QINV;
    c = ((u128)t[0] + (u128)m * QM[0]) >> 64;
    for (int j = 1; j < 6; j++) {
      c += (u128)t[j] + (u128)m * QM[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[5] = (uint64_t)c;
    t[6] = t[7] + (uint64_t)(c >> 64);
  }
  if (t[6] || geq6(t, QM)) sub6(t, t, QM);
  memcpy(r->l, t, 48);
}
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static void fe_inv(fe* r, const fe* a) { /* a^(q-2), square-and-multiply */
  uint64_t e[6];
  uint64_t two[6] = {2, 0, 0, 0, 0, 0};
  sub6(e, QM, two);
  fe acc = ONE;
  for (int i = 376; i >= 0; i--) {
    fe_sqr(&acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(&acc, &acc, a);
  }
  *r = acc;
}
static void fe_from_bytes(fe* r, const uint8_t* b) {
  fe t;
  memcpy(t.l, b, 48);
  fe_mul(r, &t, &R2);
}
static void fe_to_bytes(uint8_t* b, const fe* a) {
  fe one_raw = {{1, 0, 0, 0, 0, 0}}, t;
  fe_mul(&t, a, &one_raw);
  memcpy(b, t.l, 48);
}

static void init(void) {
  if (g_init) return;
  /* R mod q and R^2 mod q by repeated doubling of 1 (768 doublings) */
  fe x = {{1, 0, 0, 0, 0, 0}};
  for (int i = 0; i < 384; i++) fe_dbl(&x, &x);
  ONE = x;
  for (int i = 0; i < 384; i++) fe_dbl(&x, &x);
  R2 = x;
  g_init = 1;
}

/* ---- ark-ec 0.3 short-Weierstrass Jacobian formulas, a = 0 (dbl-2009-l, madd-2007-bl, add-2007-bl) */
static void jac_set_inf(jac* p) { p->X = ONE; p->Y = ONE; memset(&p->Z, 0, sizeof(fe)); }
static int jac_is_inf(const jac* p) { return fe_is_zero(&p->Z); }
static void jac_dbl(jac* r, const jac* p) {
  if (jac_is_inf(p)) { *r = *p; return; }
  fe A, B, C, D, E, F, t;
  fe_sqr(&A, &p->X);
  fe_sqr(&B, &p->Y);
  fe_sqr(&C, &B);
  fe_add(&t, &p->X, &B); fe_sqr(&t, &t); fe_sub(&t, &t, &A); fe_sub(&t, &t, &C); fe_dbl(&D, &t);
  fe_dbl(&E, &A); fe_add(&E, &E, &A);
  fe_sqr(&F, &E);
  fe Z3;
  fe_mul(&Z3, &p->Y, &p->Z); fe_dbl(&Z3, &Z3);
  fe X3;
  fe_sub(&X3, &F, &D); fe_sub(&X3, &X3, &D);
  fe Y3, c8;
  fe_dbl(&c8, &C); fe_dbl(&c8, &c8); fe_dbl(&c8, &c8);
  fe_sub(&t, &D, &X3); fe_mul(&Y3, &E, &t); fe_sub(&Y3, &Y3, &c8);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void jac_add_mixed(jac* r, const jac* p, const aff* q) {
  if (q->inf) { *r = *p; return; }
  if (jac_is_inf(p)) { r->X = q->x; r->Y = q->y; r->Z = ONE; return; }
  fe Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
  fe_sqr(&Z1Z1, &p->Z);
  fe_mul(&U2, &q->x, &Z1Z1);
  fe_mul(&S2, &q->y, &p->Z); fe_mul(&S2, &S2, &Z1Z1);
  if (memcmp(&U2, &p->X, sizeof(fe)) == 0) {
    if (memcmp(&S2, &p->Y, sizeof(fe)) == 0) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fe_sub(&H, &U2, &p->X);
  fe_sqr(&HH, &H);
  fe_dbl(&I, &HH); fe_dbl(&I, &I);
  fe_mul(&J, &H, &I);
  fe_sub(&rr, &S2, &p->Y); fe_dbl(&rr, &rr);
  fe_mul(&V, &p->X, &I);
  fe X3, Y3, Z3;
  fe_sqr(&X3, &rr); fe_sub(&X3, &X3, &J); fe_sub(&X3, &X3, &V); fe_sub(&X3, &X3, &V);
  fe_sub(&t, &V, &X3); fe_mul(&Y3, &rr, &t);
  fe_mul(&t, &p->Y, &J); fe_dbl(&t, &t); fe_sub(&Y3, &Y3, &t);
  fe_add(&Z3, &p->Z, &H); fe_sqr(&Z3, &Z3); fe_sub(&Z3, &Z3, &Z1Z1); fe_sub(&Z3, &Z3, &HH);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void jac_add(jac* r, const jac* p, const jac* q) {
  if (jac_is_inf(p)) { *r = *q; return; }
  if (jac_is_inf(q)) { *r = *p; return; }
  fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
  fe_sqr(&Z1Z1, &p->Z); fe_sqr(&Z2Z2, &q->Z);
  fe_mul(&U1, &p->X, &Z2Z2); fe_mul(&U2, &q->X, &Z1Z1);
  fe_mul(&S1, &p->Y, &q->Z); fe_mul(&S1, &S1, &Z2Z2);
  fe_mul(&S2, &q->Y, &p->Z); fe_mul(&S2, &S2, &Z1Z1);
  if (memcmp(&U1, &U2, sizeof(fe)) == 0) {
    if (memcmp(&S1, &S2, sizeof(fe)) == 0) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  fe_sub(&H, &U2, &U1);
  fe_dbl(&I, &H); fe_sqr(&I, &I);
  fe_mul(&J, &H, &I);
  fe_sub(&rr, &S2, &S1); fe_dbl(&rr, &rr);
  fe_mul(&V, &U1, &I);
  fe X3, Y3, Z3;
  fe_sqr(&X3, &rr); fe_sub(&X3, &X3, &J); fe_sub(&X3, &X3, &V); fe_sub(&X3, &X3, &V);
  fe_sub(&t, &V, &X3); fe_mul(&Y3, &rr, &t);
  fe_mul(&t, &S1, &J); fe_dbl(&t, &t); fe_sub(&Y3, &Y3, &t);
  fe_add(&Z3, &p->Z, &q->Z); fe_sqr(&Z3, &Z3); fe_sub(&Z3, &Z3, &Z1Z1); fe_sub(&Z3, &Z3, &Z2Z2); fe_mul(&Z3, &Z3, &H);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void jac_to_aff(aff* r, const jac* p) {
  if (jac_is_inf(p)) { memset(r, 0, sizeof(aff)); r->inf = 1; return; }
  fe zi, zi2, zi3;
  fe_inv(&zi, &p->Z);
  fe_sqr(&zi2, &zi);
  fe_mul(&zi3, &zi2, &zi);
  fe_mul(&r->x, &p->X, &zi2);
  fe_mul(&r->y, &p->Y, &zi3);
  r->inf = 0;
}
static void aff_from_bytes(aff* r, const uint8_t* b) {
  static const uint8_t zero[96] = {0};
  if (memcmp(b, zero, 96) == 0) { memset(r, 0, sizeof(aff)); r->inf = 1; return; }
  fe_from_bytes(&r->x, b);
  fe_from_bytes(&r->y, b + 48);
  r->inf = 0;
}
static void aff_to_bytes(uint8_t* b, const aff* p) {
  if (p->inf) { memset(b, 0, 96); return; }
  fe_to_bytes(b, &p->x);
  fe_to_bytes(b + 48, &p->y);
}
/* k * P, MSB-first double-and-add over the 256 bits of the canonical scalar (ark-ec 0.3 `mul_bits`) */
static void aff_mul(jac* r, const aff* p, const uint64_t* k) {
  jac acc;
  jac_set_inf(&acc);
  for (int i = 255; i >= 0; i--) {
    jac_dbl(&acc, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) jac_add_mixed(&acc, &acc, p);
  }
  *r = acc;
}

static int ark_window(size_t n) {
  if (n < 32) return 3;
  int lg = 0;
  while (((size_t)1 << lg) < n) lg++;
  return lg * 69 / 100 + 2;
}
/* ark-ec 0.3 VariableBaseMSM::multi_scalar_mul; num_bits = 253 (ark_bls12_377::Fr) */
static void msm_pippenger(jac* out, const aff* bases, const uint64_t (*k)[4], size_t n) {
  const int c = ark_window(n), num_bits = 253;
  const int nwin = (num_bits + c - 1) / c;
  const size_t nb = ((size_t)1 << c) - 1;
  jac* wsum = (jac*)malloc(sizeof(jac) * nwin);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1) if (g_threads > 1)
  for (int w = 0; w < nwin; w++) {
    const int start = w * c;
    jac res;
    jac_set_inf(&res);
    jac* buckets = (jac*)malloc(sizeof(jac) * nb);
    for (size_t b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
    for (size_t i = 0; i < n; i++) {
      const uint64_t* s = k[i];
      if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
      if (s[0] == 1 && (s[1] | s[2] | s[3]) == 0) {
        if (start == 0) jac_add_mixed(&res, &res, &bases[i]);
        continue;
      }
      const int word = start >> 6, sh = start & 63;
      uint64_t d = s[word] >> sh;
      if (sh && word < 3) d |= s[word + 1] << (64 - sh);
      d &= ((uint64_t)1 << c) - 1;
      if (d) jac_add_mixed(&buckets[d - 1], &buckets[d - 1], &bases[i]);
    }
    jac run;
    jac_set_inf(&run);
    for (size_t b = nb; b-- > 0;) {
      jac_add(&run, &run, &buckets[b]);
      jac_add(&res, &res, &run);
    }
    free(buckets);
    wsum[w] = res;
  }
  jac total;
  jac_set_inf(&total);
  for (int w = nwin - 1; w >= 1; w--) {
    jac_add(&total, &total, &wsum[w]);
    for (int j = 0; j < c; j++) jac_dbl(&total, &total);
  }
  jac_add(out, &total, &wsum[0]);
  free(wsum);
}
static void msm_naive(jac* out, const aff* bases, const uint64_t (*k)[4], size_t n) {
  jac total;
  jac_set_inf(&total);
#pragma omp parallel num_threads(g_threads) if (g_threads > 1)
  {
    jac part;
    jac_set_inf(&part);
#pragma omp for schedule(static) nowait
    for (size_t i = 0; i < n; i++) {
      jac t;
      aff_mul(&t, &bases[i], k[i]);
      jac_add(&part, &part, &t);
    }
#pragma omp critical
    jac_add(&total, &total, &part);
  }
  *out = total;
}

/* ---- exports ------------------------------------------------------------------------------ */
void oc377_set_threads(int t) { g_threads = t < 1 ? 1 : t; }
/* out = sum k_i * P_i per component.  mode 0: one double-and-add per term (how proof-essentials'
 * ciphertext dot products run); mode 1: ark-ec 0.3 VariableBaseMSM (Pedersen commitments).
 * points: n * ncomp * 96 bytes, scalars: n * 32 bytes (reduced mod r here), out: ncomp * 96 bytes. */
int oc377_msm(const uint8_t* points, const uint8_t* scalars, uint64_t n, int ncomp, int mode, uint8_t* out) {
  init();
  aff* p = (aff*)malloc(sizeof(aff) * (n ? n : 1));
  uint64_t(*k)[4] = (uint64_t(*)[4])malloc(32 * (n ? n : 1));
  for (uint64_t i = 0; i < n; i++) {
    memcpy(k[i], scalars + 32 * i, 32);
    for (;;) { /* canonical representative below r */
      int ge = 1;
      for (int j = 3; j >= 0; j--) {
        if (k[i][j] > RM[j]) break;
        if (k[i][j] < RM[j]) { ge = 0; break; }
      }
      if (!ge) break;
      uint64_t borrow = 0;
      for (int j = 0; j < 4; j++) {
        u128 d = (u128)k[i][j] - RM[j] - borrow;
        k[i][j] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
      }
    }
  }
  for (int comp = 0; comp < ncomp; comp++) {
    for (uint64_t i = 0; i < n; i++) aff_from_bytes(&p[i], points + 96 * (i * ncomp + comp));
    jac j;
    if (mode == 0) msm_naive(&j, p, (const uint64_t(*)[4])k, n);
    else msm_pippenger(&j, p, (const uint64_t(*)[4])k, n);
    aff a;
    jac_to_aff(&a, &j);
    aff_to_bytes(out + 96 * comp, &a);
  }
  free(p);
  free(k);
  return 0;
}
void oc377_fq_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) { /* canonical in, canonical out */
  init();
  fe x, y, z;
  fe_from_bytes(&x, a);
  fe_from_bytes(&y, b);
  fe_mul(&z, &x, &y);
  fe_to_bytes(out, &z);
}
