/* ORACLE (test infrastructure only -- only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does).
 *
 * CPU restatement, in plain C, of the reference's shuffle hot path:
 *   DLCards::shuffle_and_remask  barnett-smart-card-protocol/src/discrete_log_cards/mod.rs:380-418
 *   DLCards::verify_shuffle      .../mod.rs:420-443
 *   MaskedCard::remask           .../remasking.rs:9-22  -> Card::mask .../masking.rs:10-20
 * and of what those calls execute inside the un-vendored dependencies (proof-essentials @
 * unpinned git, arkworks 0.3.0; Cargo.toml:10-20), restated from SURVEY.md Appendix A/B:
 * Bayer-Groth shuffle argument (B.1-B.5'), Pedersen commitments via the ark-ec 0.3
 * VariableBaseMSM algorithm, ElGamal ciphertext algebra via per-term double-and-add
 * (`msm_mode` 0, the faithful CPU cost model) or Pippenger (`msm_mode` 1, best-effort CPU),
 * FiatShamirRng<Blake2s> transcript with the byte/draw order of Appendix B.6.
 *
 * PARITY UNPINNED against upstream bytes (no Rust toolchain, deps absent, reference ships no
 * vectors).  Pinned against: oracle/py (big-int Python), tests/golden/oracle_vectors.json, RFC
 * vectors for Blake2s/ChaCha20, and the reference's behavioural test (tests.rs:175-227).
 * The flat proof layout is the one include/mpshuffle.h documents.
 */
#include <stdio.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "curve.h"
#include "hash.h"

field_t FQ, FR;
fe CURVE_B;
static int g_threads = 1;
static int g_msm_mode = 0;
static int g_inited = 0;

void oracle_init(void) {
  if (g_inited) return;
  static const uint64_t P[4] = {0x0000000000000001ull, 0, 0, 0x0800000000000011ull};
  static const uint64_t N[4] = {0x1e66a241adc64d2full, 0xb781126dcae7b232ull, 0xffffffffffffffffull, 0x0800000000000010ull};
  static const uint64_t B[4] = {0xf4cdfcb99cee9e89ull, 0x609ad26c15c915c1ull, 0x150e596d72f7a8c5ull, 0x06f21413efbe40deull};
  field_init(&FQ, P);
  field_init(&FR, N);
  fe_from_raw(&CURVE_B, B, &FQ);
  g_inited = 1;
}
void oracle_set_threads(int t) { g_threads = t < 1 ? 1 : t; }
void oracle_set_msm_mode(int mode) { g_msm_mode = mode; }
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * MSM
 * ---------------------------------------------------------------------------------------- */
static int ark_window(size_t n) {
  if (n < 32) return 3;
  int lg = 0; /* ark_std::log2 = ceil(log2 n) */
  while (((size_t)1 << lg) < n) lg++;
  return lg * 69 / 100 + 2;
}

/* ark-ec 0.3 VariableBaseMSM::multi_scalar_mul on canonical scalars k[i][4];
 * window c (0 = arkworks' choice).  Zero scalars skipped, scalars == 1 added directly. */
static void msm_pippenger(jac* out, const aff* bases, const uint64_t (*k)[4], size_t n, int c) {
  if (c <= 0) c = ark_window(n);
  const int num_bits = 252;
  int nwin = (num_bits + c - 1) / c;
  size_t nb = ((size_t)1 << c) - 1;
  jac* wsum = (jac*)malloc(sizeof(jac) * nwin);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1) if (g_threads > 1)
  for (int w = 0; w < nwin; w++) {
    int start = w * c;
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
      int word = start >> 6, sh = start & 63;
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

/* sum_i k_i * P_i with one double-and-add per term (ark-ec 0.3 `AffineCurve::mul`), the cost
 * model of the reference's ciphertext dot products [UPSTREAM-RECALL, SURVEY.md section 2b K1] */
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
      aff_mul_raw(&t, &bases[i], k[i]);
      jac_add(&part, &part, &t);
    }
#pragma omp critical
    jac_add(&total, &total, &part);
  }
  *out = total;
}

static void msm_dispatch(jac* out, const aff* bases, const fe* scal, size_t n, int faithful_naive) {
  uint64_t(*k)[4] = (uint64_t(*)[4])malloc(32 * (n ? n : 1));
  for (size_t i = 0; i < n; i++) fe_to_raw(k[i], &scal[i], &FR);
  if (faithful_naive && g_msm_mode == 0) msm_naive(out, bases, (const uint64_t(*)[4])k, n);
  else msm_pippenger(out, bases, (const uint64_t(*)[4])k, n, 0);
  free(k);
}

/* ------------------------------------------------------------------------------------------
 * protocol building blocks
 * ---------------------------------------------------------------------------------------- */
typedef struct { aff c1, c2; } ct_t;
typedef struct {
  int m, n;
  aff enc_g, ck_h, ghat, pk;
  aff* ck; /* [h, g_1 .. g_n]: Pedersen bases in MSM order */
} params_t;

/* com(v; r) = r*h + sum v_j*g_j through VariableBaseMSM over [h, g...] x [r, v...] */
static void commit(aff* out, const params_t* pp, const fe* v, int len, const fe* r) {
  fe* s = (fe*)malloc(sizeof(fe) * (len + 1));
  s[0] = *r;
  memcpy(s + 1, v, sizeof(fe) * len);
  jac j;
  msm_dispatch(&j, pp->ck, s, (size_t)len + 1, 0);
  jac_to_aff(out, &j);
  free(s);
}
static void commit_const(aff* out, const params_t* pp, const fe* value, int len) {
  fe* v = (fe*)malloc(sizeof(fe) * len);
  for (int i = 0; i < len; i++) v[i] = *value;
  fe zero;
  fe_set_zero(&zero);
  commit(out, pp, v, len, &zero);
  free(v);
}
static void pt_add(aff* out, const aff* a, const aff* b) {
  jac j;
  jac_from_aff(&j, a);
  jac_add_mixed(&j, &j, b);
  jac_to_aff(out, &j);
}
static void pt_mul(aff* out, const aff* p, const fe* k) {
  jac j;
  aff_mul(&j, p, k);
  jac_to_aff(out, &j);
}
/* sum k_i * P_i over commitments (small) */
static void pt_lincomb(aff* out, const aff* pts, const fe* k, int n) {
  jac j;
  msm_dispatch(&j, pts, k, (size_t)n, 1);
  jac_to_aff(out, &j);
}
/* ElGamal::encrypt(msg; r) = (r*g, msg + r*pk)   masking.rs:17 */
static void encrypt(ct_t* out, const params_t* pp, const aff* msg, const fe* r) {
  jac a, b;
  aff_mul(&a, &pp->enc_g, r);
  aff_mul(&b, &pp->pk, r);
  jac_add_mixed(&b, &b, msg);
  jac_to_aff(&out->c1, &a);
  jac_to_aff(&out->c2, &b);
}
static void ct_add(ct_t* out, const ct_t* a, const ct_t* b) {
  pt_add(&out->c1, &a->c1, &b->c1);
  pt_add(&out->c2, &a->c2, &b->c2);
}
static int ct_eq(const ct_t* a, const ct_t* b) { return aff_eq(&a->c1, &b->c1) && aff_eq(&a->c2, &b->c2); }
/* ciphertext dot product sum k_i * C_i (component-wise) */
static void ct_msm(ct_t* out, const ct_t* cts, const fe* k, size_t n) {
  aff* p = (aff*)malloc(sizeof(aff) * (n ? n : 1));
  jac j;
  for (int comp = 0; comp < 2; comp++) {
    for (size_t i = 0; i < n; i++) p[i] = comp ? cts[i].c2 : cts[i].c1;
    msm_dispatch(&j, p, k, n, 1);
    jac_to_aff(comp ? &out->c2 : &out->c1, &j);
  }
  free(p);
}

/* u * v = sum_j u_j v_j y^j (j = 1..n) */
static void bilinear(fe* out, const fe* u, const fe* v, int n, const fe* y) {
  fe acc, yp = FR.one, t;
  fe_set_zero(&acc);
  for (int j = 0; j < n; j++) {
    fe_mul(&yp, &yp, y, &FR);
    fe_mul(&t, &u[j], &v[j], &FR);
    fe_mul(&t, &t, &yp, &FR);
    fe_add(&acc, &acc, &t, &FR);
  }
  *out = acc;
}
static void powers(fe* xp, const fe* x, int count) { /* xp[k] = x^k, k = 0..count-1 */
  if (count > 0) xp[0] = FR.one;
  for (int k = 1; k < count; k++) fe_mul(&xp[k], &xp[k - 1], x, &FR);
}
/* out[j] = sum_i coeff[i] * vecs[i][j] */
static void lincomb(fe* out, const fe* coeff, fe* const* vecs, int count, int n) {
  for (int j = 0; j < n; j++) {
    fe acc, t;
    fe_set_zero(&acc);
    for (int i = 0; i < count; i++) {
      fe_mul(&t, &coeff[i], &vecs[i][j], &FR);
      fe_add(&acc, &acc, &t, &FR);
    }
    out[j] = acc;
  }
}
static void dot(fe* out, const fe* a, const fe* b, int n) {
  fe acc, t;
  fe_set_zero(&acc);
  for (int i = 0; i < n; i++) {
    fe_mul(&t, &a[i], &b[i], &FR);
    fe_add(&acc, &acc, &t, &FR);
  }
  *out = acc;
}

/* transcript helpers */
static void feed_label(fsrng_t* fs, const char* label) { fs_absorb_feed(fs, (const uint8_t*)label, strlen(label)); }
static void feed_pts(fsrng_t* fs, const aff* p, int n) {
  uint8_t b[65];
  for (int i = 0; i < n; i++) { aff_to_bytes65(b, &p[i]); fs_absorb_feed(fs, b, 65); }
}
static void feed_cts(fsrng_t* fs, const ct_t* c, size_t n) {
  uint8_t b[130];
  for (size_t i = 0; i < n; i++) {
    aff_to_bytes65(b, &c[i].c1);
    aff_to_bytes65(b + 65, &c[i].c2);
    fs_absorb_feed(fs, b, 130);
  }
}

/* flat proof cursor (layout of include/mpshuffle.h / oracle/py proof_to_bytes) */
typedef struct { uint8_t* p; } wr_t;
typedef struct { const uint8_t* p; } rd_t;
static void wr_pt(wr_t* w, const aff* a) { aff_to_bytes64(w->p, a); w->p += 64; }
static void wr_pts(wr_t* w, const aff* a, int n) { for (int i = 0; i < n; i++) wr_pt(w, &a[i]); }
static void wr_fr(wr_t* w, const fe* a) { fe_to_bytes(w->p, a, &FR); w->p += 32; }
static void wr_frs(wr_t* w, const fe* a, int n) { for (int i = 0; i < n; i++) wr_fr(w, &a[i]); }
static void rd_pt(rd_t* r, aff* a) { aff_from_bytes64(a, r->p); r->p += 64; }
static void rd_pts(rd_t* r, aff* a, int n) { for (int i = 0; i < n; i++) rd_pt(r, &a[i]); }
static void rd_fr(rd_t* r, fe* a) { fe_from_bytes(a, r->p, &FR); r->p += 32; }
static void rd_frs(rd_t* r, fe* a, int n) { for (int i = 0; i < n; i++) rd_fr(r, &a[i]); }

typedef struct { const fe* s; size_t i; } rand_t; /* flat prover randomness, Appendix B.6 order */
static fe rnd1(rand_t* r) { return r->s[r->i++]; }
static void rndv(rand_t* r, fe* out, int k) { for (int i = 0; i < k; i++) out[i] = rnd1(r); }

static inline void* new_array(size_t count, size_t size) { return calloc(count ? count : 1, size); }
#define NEW(T, count) ((T*)new_array((size_t)(count), sizeof(T)))

enum { OK = 0, ERR_HADAMARD = 1, ERR_ZERO = 2, ERR_SVP = 3, ERR_MULTIEXP = 4 };

/* ------------------------------------------------------------------------------------------
 * B.4 zero argument.  Statement: cA[1..m], cB[1..m], y.  Witness rows A[m][n], r[m], B[m][n], s[m].
 * Proof: c_A0, c_Bm1, c_D[2m+1] | a[n], b[n], r, s, t
 * ---------------------------------------------------------------------------------------- */
static void zero_prove(const params_t* pp, fsrng_t* fs, rand_t* rd, wr_t* pts_out, wr_t* frs_out, int m,
                       const fe* y, fe* const* A, const fe* r, fe* const* B, const fe* s) {
  int n = pp->n;
  fe* a0 = NEW(fe, n); fe* bm1 = NEW(fe, n);
  rndv(rd, a0, n); rndv(rd, bm1, n);
  fe r0 = rnd1(rd), sm1 = rnd1(rd);
  fe* t = NEW(fe, 2 * m + 1);
  for (int k = 0; k <= 2 * m; k++) { if (k != m + 1) t[k] = rnd1(rd); else fe_set_zero(&t[k]); }
  const fe** Ae = NEW(const fe*, m + 1); const fe** Be = NEW(const fe*, m + 1);
  fe* re = NEW(fe, m + 1); fe* se = NEW(fe, m + 1);
  Ae[0] = a0; re[0] = r0;
  for (int i = 0; i < m; i++) { Ae[i + 1] = A[i]; re[i + 1] = r[i]; Be[i] = B[i]; se[i] = s[i]; }
  Be[m] = bm1; se[m] = sm1;
  aff* cpts = NEW(aff, 2 * m + 3); /* c_A0, c_Bm1, c_D... */
  commit(&cpts[0], pp, a0, n, &r0);
  commit(&cpts[1], pp, bm1, n, &sm1);
  fe* d = NEW(fe, 2 * m + 1);
  for (int i = 0; i <= m; i++)
    for (int j = 1; j <= m + 1; j++) {
      fe v;
      bilinear(&v, Ae[i], Be[j - 1], n, y);
      int k = i + m + 1 - j;
      fe_add(&d[k], &d[k], &v, &FR);
    }
  for (int k = 0; k <= 2 * m; k++) commit(&cpts[2 + k], pp, &d[k], 1, &t[k]);
  fs_absorb_begin(fs); feed_label(fs, "zero_argument"); feed_pts(fs, cpts, 2 * m + 3); fs_absorb_end(fs);
  fe x;
  fs_challenge(fs, &x, &FR);
  fe* xp = NEW(fe, 2 * m + 1);
  powers(xp, &x, 2 * m + 1);
  fe* av = NEW(fe, n); fe* bv = NEW(fe, n);
  lincomb(av, xp, (fe* const*)Ae, m + 1, n);
  fe* xr = NEW(fe, m + 1); /* x^{m+1-j}, j = 1..m+1 */
  for (int j = 1; j <= m + 1; j++) xr[j - 1] = xp[m + 1 - j];
  lincomb(bv, xr, (fe* const*)Be, m + 1, n);
  fe rr, ss, tt;
  dot(&rr, xp, re, m + 1);
  dot(&ss, xr, se, m + 1);
  dot(&tt, xp, t, 2 * m + 1);
  wr_pts(pts_out, cpts, 2 * m + 3);
  wr_frs(frs_out, av, n); wr_frs(frs_out, bv, n);
  wr_fr(frs_out, &rr); wr_fr(frs_out, &ss); wr_fr(frs_out, &tt);
  free(a0); free(bm1); free(t); free(Ae); free(Be); free(re); free(se); free(cpts); free(d); free(xp);
  free(av); free(bv); free(xr);
}

static int zero_verify(const params_t* pp, fsrng_t* fs, rd_t* in, int m, const aff* cA, const aff* cB, const fe* y) {
  int n = pp->n, st = OK;
  aff* cpts = NEW(aff, 2 * m + 3);
  rd_pts(in, cpts, 2 * m + 3);
  fe* av = NEW(fe, n); fe* bv = NEW(fe, n);
  rd_frs(in, av, n); rd_frs(in, bv, n);
  fe rr, ss, tt;
  rd_fr(in, &rr); rd_fr(in, &ss); rd_fr(in, &tt);
  fs_absorb_begin(fs); feed_label(fs, "zero_argument"); feed_pts(fs, cpts, 2 * m + 3); fs_absorb_end(fs);
  fe x;
  fs_challenge(fs, &x, &FR);
  fe* xp = NEW(fe, 2 * m + 1);
  powers(xp, &x, 2 * m + 1);
  aff* pts = NEW(aff, m + 1); fe* xr = NEW(fe, m + 1);
  aff lhs, rhs;
  if (!cpts[2 + m + 1].inf) st = ERR_ZERO;
  if (st == OK) {
    pts[0] = cpts[0];
    for (int i = 0; i < m; i++) pts[i + 1] = cA[i];
    pt_lincomb(&lhs, pts, xp, m + 1);
    commit(&rhs, pp, av, n, &rr);
    if (!aff_eq(&lhs, &rhs)) st = ERR_ZERO;
  }
  if (st == OK) {
    for (int j = 1; j <= m + 1; j++) xr[j - 1] = xp[m + 1 - j];
    for (int i = 0; i < m; i++) pts[i] = cB[i];
    pts[m] = cpts[1];
    pt_lincomb(&lhs, pts, xr, m + 1);
    commit(&rhs, pp, bv, n, &ss);
    if (!aff_eq(&lhs, &rhs)) st = ERR_ZERO;
  }
  if (st == OK) {
    fe v;
    bilinear(&v, av, bv, n, y);
    pt_lincomb(&lhs, cpts + 2, xp, 2 * m + 1);
    commit(&rhs, pp, &v, 1, &tt);
    if (!aff_eq(&lhs, &rhs)) st = ERR_ZERO;
  }
  free(cpts); free(av); free(bv); free(xp); free(pts); free(xr);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * B.3 Hadamard argument.  Proof: c_B[m] | zero proof
 * ---------------------------------------------------------------------------------------- */
static void hadamard_prove(const params_t* pp, fsrng_t* fs, rand_t* rd, wr_t* out, int m, const aff* cA,
                           const aff* c_b, fe* const* A, const fe* r, const fe* s) {
  int n = pp->n;
  fe** Bv = NEW(fe*, m);
  for (int i = 0; i < m; i++) Bv[i] = NEW(fe, n);
  memcpy(Bv[0], A[0], sizeof(fe) * n);
  for (int i = 1; i < m; i++)
    for (int j = 0; j < n; j++) fe_mul(&Bv[i][j], &Bv[i - 1][j], &A[i][j], &FR);
  fe* sv = NEW(fe, m);
  sv[0] = r[0];
  for (int i = 1; i < m - 1; i++) sv[i] = rnd1(rd);
  if (m >= 2) sv[m - 1] = *s;
  aff* c_B = NEW(aff, m);
  c_B[0] = cA[0];
  for (int i = 1; i < m - 1; i++) commit(&c_B[i], pp, Bv[i], n, &sv[i]);
  c_B[m - 1] = *c_b;
  fs_absorb_begin(fs); feed_label(fs, "hadamard_argument"); feed_pts(fs, c_b, 1); feed_pts(fs, c_B, m); fs_absorb_end(fs);
  fe x, y;
  fs_challenge(fs, &x, &FR);
  fs_challenge(fs, &y, &FR);
  fe* xp = NEW(fe, m);
  powers(xp, &x, m);
  /* zero-argument instance */
  fe minus1;
  fe_neg(&minus1, &FR.one, &FR);
  fe* m1v = NEW(fe, n);
  for (int j = 0; j < n; j++) m1v[j] = minus1;
  fe** zA = NEW(fe*, m); fe** zB = NEW(fe*, m);
  fe* zr = NEW(fe, m); fe* zs = NEW(fe, m);
  aff* zcA = NEW(aff, m); aff* zcB = NEW(aff, m);
  for (int i = 0; i < m - 1; i++) { zA[i] = A[i + 1]; zr[i] = r[i + 1]; zcA[i] = cA[i + 1]; }
  zA[m - 1] = m1v;
  fe_set_zero(&zr[m - 1]);
  commit_const(&zcA[m - 1], pp, &minus1, n);
  for (int i = 1; i < m; i++) { /* D_i = x^i * b_i (b_i = Bv[i-1]) */
    zB[i - 1] = NEW(fe, n);
    for (int j = 0; j < n; j++) fe_mul(&zB[i - 1][j], &xp[i], &Bv[i - 1][j], &FR);
    fe_mul(&zs[i - 1], &xp[i], &sv[i - 1], &FR);
    pt_mul(&zcB[i - 1], &c_B[i - 1], &xp[i]);
  }
  zB[m - 1] = NEW(fe, n);
  lincomb(zB[m - 1], xp + 1, Bv + 1, m - 1, n);
  dot(&zs[m - 1], xp + 1, sv + 1, m - 1);
  pt_lincomb(&zcB[m - 1], c_B + 1, xp + 1, m - 1);
  wr_pts(out, c_B, m);
  /* zero proof: points then scalars, contiguous */
  wr_t zpts = {out->p};
  wr_t zfrs = {out->p + 64 * (2 * m + 3)};
  zero_prove(pp, fs, rd, &zpts, &zfrs, m, &y, zA, zr, zB, zs);
  out->p = zfrs.p;
  for (int i = 0; i < m; i++) { free(Bv[i]); free(zB[i]); }
  free(Bv); free(sv); free(c_B); free(xp); free(m1v); free(zA); free(zB); free(zr); free(zs); free(zcA); free(zcB);
}

static int hadamard_verify(const params_t* pp, fsrng_t* fs, rd_t* in, int m, const aff* cA, const aff* c_b) {
  int n = pp->n;
  aff* c_B = NEW(aff, m);
  rd_pts(in, c_B, m);
  if (!aff_eq(&c_B[0], &cA[0]) || !aff_eq(&c_B[m - 1], c_b)) { free(c_B); return ERR_HADAMARD; }
  fs_absorb_begin(fs); feed_label(fs, "hadamard_argument"); feed_pts(fs, c_b, 1); feed_pts(fs, c_B, m); fs_absorb_end(fs);
  fe x, y;
  fs_challenge(fs, &x, &FR);
  fs_challenge(fs, &y, &FR);
  fe* xp = NEW(fe, m);
  powers(xp, &x, m);
  fe minus1;
  fe_neg(&minus1, &FR.one, &FR);
  aff* zcA = NEW(aff, m); aff* zcB = NEW(aff, m);
  for (int i = 0; i < m - 1; i++) zcA[i] = cA[i + 1];
  commit_const(&zcA[m - 1], pp, &minus1, n);
  for (int i = 1; i < m; i++) pt_mul(&zcB[i - 1], &c_B[i - 1], &xp[i]);
  pt_lincomb(&zcB[m - 1], c_B + 1, xp + 1, m - 1);
  int st = zero_verify(pp, fs, in, m, zcA, zcB, &y);
  free(c_B); free(xp); free(zcA); free(zcB);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * B.5 single-value product.  Proof: c_d, c_delta, c_Delta | a~[n], b~[n], r~, s~
 * ---------------------------------------------------------------------------------------- */
static void svp_prove(const params_t* pp, fsrng_t* fs, rand_t* rd, wr_t* out, const fe* a, const fe* r) {
  int n = pp->n;
  fe* bk = NEW(fe, n);
  bk[0] = a[0];
  for (int i = 1; i < n; i++) fe_mul(&bk[i], &bk[i - 1], &a[i], &FR);
  fe* d = NEW(fe, n);
  rndv(rd, d, n);
  fe r_d = rnd1(rd);
  fe* delta = NEW(fe, n);
  delta[0] = d[0];
  for (int i = 1; i < n - 1; i++) delta[i] = rnd1(rd);
  fe_set_zero(&delta[n - 1]);
  fe s1 = rnd1(rd), sx = rnd1(rd);
  aff c[3];
  commit(&c[0], pp, d, n, &r_d);
  fe* v = NEW(fe, n);
  fe t, u;
  for (int i = 0; i < n - 1; i++) { fe_mul(&t, &delta[i], &d[i + 1], &FR); fe_neg(&v[i], &t, &FR); }
  commit(&c[1], pp, v, n - 1, &s1);
  for (int i = 0; i < n - 1; i++) {
    fe_mul(&t, &a[i + 1], &delta[i], &FR);
    fe_sub(&u, &delta[i + 1], &t, &FR);
    fe_mul(&t, &bk[i], &d[i + 1], &FR);
    fe_sub(&v[i], &u, &t, &FR);
  }
  commit(&c[2], pp, v, n - 1, &sx);
  fs_absorb_begin(fs); feed_label(fs, "single_value_product_argument"); feed_pts(fs, c, 3); fs_absorb_end(fs);
  fe x;
  fs_challenge(fs, &x, &FR);
  wr_pts(out, c, 3);
  for (int i = 0; i < n; i++) { fe_mul(&t, &x, &a[i], &FR); fe_add(&t, &t, &d[i], &FR); wr_fr(out, &t); }
  for (int i = 0; i < n; i++) { fe_mul(&t, &x, &bk[i], &FR); fe_add(&t, &t, &delta[i], &FR); wr_fr(out, &t); }
  fe_mul(&t, &x, r, &FR); fe_add(&t, &t, &r_d, &FR); wr_fr(out, &t);
  fe_mul(&t, &x, &sx, &FR); fe_add(&t, &t, &s1, &FR); wr_fr(out, &t);
  free(bk); free(d); free(delta); free(v);
}

static int svp_verify(const params_t* pp, fsrng_t* fs, rd_t* in, const aff* c_a, const fe* b) {
  int n = pp->n, st = OK;
  aff c[3];
  rd_pts(in, c, 3);
  fe* at = NEW(fe, n); fe* bt = NEW(fe, n);
  rd_frs(in, at, n); rd_frs(in, bt, n);
  fe rr, ss;
  rd_fr(in, &rr); rd_fr(in, &ss);
  fs_absorb_begin(fs); feed_label(fs, "single_value_product_argument"); feed_pts(fs, c, 3); fs_absorb_end(fs);
  fe x, t, u;
  fs_challenge(fs, &x, &FR);
  aff lhs, rhs;
  pt_mul(&lhs, c_a, &x);
  pt_add(&lhs, &lhs, &c[0]);
  commit(&rhs, pp, at, n, &rr);
  if (!aff_eq(&lhs, &rhs)) st = ERR_SVP;
  if (st == OK) {
    fe* e = NEW(fe, n);
    for (int i = 0; i < n - 1; i++) {
      fe_mul(&t, &x, &bt[i + 1], &FR);
      fe_mul(&u, &bt[i], &at[i + 1], &FR);
      fe_sub(&e[i], &t, &u, &FR);
    }
    pt_mul(&lhs, &c[2], &x);
    pt_add(&lhs, &lhs, &c[1]);
    commit(&rhs, pp, e, n - 1, &ss);
    if (!aff_eq(&lhs, &rhs)) st = ERR_SVP;
    free(e);
  }
  if (st == OK) {
    fe_mul(&t, &x, b, &FR);
    if (!fe_eq(&bt[0], &at[0]) || !fe_eq(&bt[n - 1], &t)) st = ERR_SVP;
  }
  free(at); free(bt);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * B.2 product argument.  Proof: c_b | hadamard | svp
 * ---------------------------------------------------------------------------------------- */
static void product_prove(const params_t* pp, fsrng_t* fs, rand_t* rd, wr_t* out, int m, const aff* cA,
                          fe* const* A, const fe* r) {
  int n = pp->n;
  fe s = rnd1(rd);
  fe* col = NEW(fe, n);
  memcpy(col, A[0], sizeof(fe) * n);
  for (int i = 1; i < m; i++)
    for (int j = 0; j < n; j++) fe_mul(&col[j], &col[j], &A[i][j], &FR);
  aff c_b;
  commit(&c_b, pp, col, n, &s);
  wr_pt(out, &c_b);
  hadamard_prove(pp, fs, rd, out, m, cA, &c_b, A, r, &s);
  svp_prove(pp, fs, rd, out, col, &s);
  free(col);
}
static int product_verify(const params_t* pp, fsrng_t* fs, rd_t* in, int m, const aff* cA, const fe* b) {
  aff c_b;
  rd_pt(in, &c_b);
  int st = hadamard_verify(pp, fs, in, m, cA, &c_b);
  if (st != OK) return st;
  return svp_verify(pp, fs, in, &c_b, b);
}

/* ------------------------------------------------------------------------------------------
 * B.5' multi-exponentiation.  Proof: c_A0, c_B[2m], E[2m] | a[n], r, b, s, tau
 * ---------------------------------------------------------------------------------------- */
static void feed_multiexp(fsrng_t* fs, const aff* cpts, const ct_t* E, int m) {
  fs_absorb_begin(fs);
  feed_label(fs, "multi_exponentiation_argument");
  feed_pts(fs, cpts, 2 * m + 1);
  feed_cts(fs, E, (size_t)2 * m);
  fs_absorb_end(fs);
}
static void multiexp_prove(const params_t* pp, fsrng_t* fs, rand_t* rd, wr_t* out, int m, const ct_t* deck2,
                           const ct_t* C, fe* const* A, const fe* r, const fe* rho) {
  int n = pp->n;
  fe* a0 = NEW(fe, n);
  rndv(rd, a0, n);
  fe r0 = rnd1(rd);
  fe* b = NEW(fe, 2 * m); fe* s = NEW(fe, 2 * m); fe* tau = NEW(fe, 2 * m);
  for (int k = 0; k < 2 * m; k++) {
    if (k == m) { fe_set_zero(&b[k]); fe_set_zero(&s[k]); tau[k] = *rho; }
    else { b[k] = rnd1(rd); s[k] = rnd1(rd); tau[k] = rnd1(rd); }
  }
  const fe** Ae = NEW(const fe*, m + 1);
  fe* re = NEW(fe, m + 1);
  Ae[0] = a0; re[0] = r0;
  for (int i = 0; i < m; i++) { Ae[i + 1] = A[i]; re[i + 1] = r[i]; }
  aff* cpts = NEW(aff, 2 * m + 1);
  commit(&cpts[0], pp, a0, n, &r0);
  for (int k = 0; k < 2 * m; k++) commit(&cpts[1 + k], pp, &b[k], 1, &s[k]);
  ct_t* E = NEW(ct_t, 2 * m);
  for (int k = 0; k < 2 * m; k++) {
    aff gb;
    pt_mul(&gb, &pp->ghat, &b[k]);
    ct_t acc, t;
    encrypt(&acc, pp, &gb, &tau[k]);
    for (int i = 1; i <= m; i++) {
      int j = k - m + i;
      if (j < 0 || j > m) continue;
      ct_msm(&t, deck2 + (size_t)(i - 1) * n, Ae[j], (size_t)n);
      ct_add(&acc, &acc, &t);
    }
    E[k] = acc;
  }
  if (!ct_eq(&E[m], C)) fprintf(stderr, "oracle: multi-exp witness does not open the statement\n");
  feed_multiexp(fs, cpts, E, m);
  fe x;
  fs_challenge(fs, &x, &FR);
  fe* xp = NEW(fe, 2 * m);
  powers(xp, &x, 2 * m);
  fe* av = NEW(fe, n);
  lincomb(av, xp, (fe* const*)Ae, m + 1, n);
  fe rr, bb, ss, tt;
  dot(&rr, xp, re, m + 1);
  dot(&bb, xp, b, 2 * m);
  dot(&ss, xp, s, 2 * m);
  dot(&tt, xp, tau, 2 * m);
  wr_pts(out, cpts, 2 * m + 1);
  for (int k = 0; k < 2 * m; k++) { wr_pt(out, &E[k].c1); wr_pt(out, &E[k].c2); }
  wr_frs(out, av, n);
  wr_fr(out, &rr); wr_fr(out, &bb); wr_fr(out, &ss); wr_fr(out, &tt);
  free(a0); free(b); free(s); free(tau); free(Ae); free(re); free(cpts); free(E); free(xp); free(av);
}

static int multiexp_verify(const params_t* pp, fsrng_t* fs, rd_t* in, int m, const ct_t* deck2, const ct_t* C,
                           const aff* cA) {
  int n = pp->n, st = OK;
  size_t Nc = (size_t)m * n;
  aff* cpts = NEW(aff, 2 * m + 1);
  rd_pts(in, cpts, 2 * m + 1);
  ct_t* E = NEW(ct_t, 2 * m);
  for (int k = 0; k < 2 * m; k++) { rd_pt(in, &E[k].c1); rd_pt(in, &E[k].c2); }
  fe* av = NEW(fe, n);
  rd_frs(in, av, n);
  fe rr, bb, ss, tt;
  rd_fr(in, &rr); rd_fr(in, &bb); rd_fr(in, &ss); rd_fr(in, &tt);
  feed_multiexp(fs, cpts, E, m);
  fe x;
  fs_challenge(fs, &x, &FR);
  fe* xp = NEW(fe, 2 * m);
  powers(xp, &x, 2 * m);
  aff lhs, rhs;
  aff* pts = NEW(aff, m + 1);
  if (!cpts[1 + m].inf || !ct_eq(&E[m], C)) st = ERR_MULTIEXP;
  if (st == OK) {
    pts[0] = cpts[0];
    for (int i = 0; i < m; i++) pts[i + 1] = cA[i];
    pt_lincomb(&lhs, pts, xp, m + 1);
    commit(&rhs, pp, av, n, &rr);
    if (!aff_eq(&lhs, &rhs)) st = ERR_MULTIEXP;
  }
  if (st == OK) {
    pt_lincomb(&lhs, cpts + 1, xp, 2 * m);
    commit(&rhs, pp, &bb, 1, &ss);
    if (!aff_eq(&lhs, &rhs)) st = ERR_MULTIEXP;
  }
  if (st == OK) {
    ct_t L, Rr, enc;
    ct_msm(&L, E, xp, (size_t)2 * m);
    fe* flat = NEW(fe, Nc);
    for (int i = 1; i <= m; i++)
      for (int j = 0; j < n; j++) fe_mul(&flat[(size_t)(i - 1) * n + j], &xp[m - i], &av[j], &FR);
    ct_msm(&Rr, deck2, flat, Nc);
    aff gb;
    pt_mul(&gb, &pp->ghat, &bb);
    encrypt(&enc, pp, &gb, &tt);
    ct_add(&Rr, &Rr, &enc);
    if (!ct_eq(&L, &Rr)) st = ERR_MULTIEXP;
    free(flat);
  }
  free(cpts); free(E); free(av); free(xp); free(pts);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * B.1 shuffle argument
 * ---------------------------------------------------------------------------------------- */
static void absorb_statement(fsrng_t* fs, const params_t* pp, const ct_t* deck, const ct_t* deck2, size_t Nc,
                             const aff* c_A) {
  fs_absorb_begin(fs);
  feed_label(fs, "shuffle_argument");
  feed_pts(fs, &pp->enc_g, 1);
  feed_pts(fs, &pp->pk, 1);
  feed_pts(fs, pp->ck + 1, pp->n);
  feed_pts(fs, &pp->ck_h, 1);
  feed_pts(fs, &pp->ghat, 1);
  feed_cts(fs, deck, Nc);
  feed_cts(fs, deck2, Nc);
  feed_pts(fs, c_A, pp->m);
  fs_absorb_end(fs);
}
/* c_D[k] = y*c_A[k] + c_B[k] + com(-z..-z; 0);  b* = prod_{i=1..N}(y*i + x^i - z) */
static void product_statement(const params_t* pp, const aff* c_A, const aff* c_B, const fe* x, const fe* y,
                              const fe* z, size_t Nc, aff* c_D, fe* bstar) {
  fe mz;
  fe_neg(&mz, z, &FR);
  aff c_mz, t;
  commit_const(&c_mz, pp, &mz, pp->n);
  for (int k = 0; k < pp->m; k++) {
    pt_mul(&t, &c_A[k], y);
    pt_add(&t, &t, &c_B[k]);
    pt_add(&c_D[k], &t, &c_mz);
  }
  fe acc = FR.one, xi = FR.one, yi, u;
  fe_set_zero(&yi);
  for (size_t i = 1; i <= Nc; i++) {
    fe_mul(&xi, &xi, x, &FR);
    fe_add(&yi, &yi, y, &FR);
    fe_add(&u, &yi, &xi, &FR);
    fe_sub(&u, &u, z, &FR);
    fe_mul(&acc, &acc, &u, &FR);
  }
  *bstar = acc;
}

static void load_params(params_t* pp, int m, int n, const uint8_t* enc_g, const uint8_t* ck_g, const uint8_t* ck_h,
                        const uint8_t* ghat, const uint8_t* pk) {
  pp->m = m; pp->n = n;
  aff_from_bytes64(&pp->enc_g, enc_g);
  aff_from_bytes64(&pp->ck_h, ck_h);
  aff_from_bytes64(&pp->ghat, ghat);
  aff_from_bytes64(&pp->pk, pk);
  pp->ck = NEW(aff, n + 1);
  pp->ck[0] = pp->ck_h;
  for (int i = 0; i < n; i++) aff_from_bytes64(&pp->ck[i + 1], ck_g + 64 * (size_t)i);
}
static ct_t* load_deck(const uint8_t* b, size_t Nc) {
  ct_t* d = NEW(ct_t, Nc);
  for (size_t i = 0; i < Nc; i++) { aff_from_bytes64(&d[i].c1, b + 128 * i); aff_from_bytes64(&d[i].c2, b + 128 * i + 64); }
  return d;
}

size_t oc_proof_len(int m, int n) { return (size_t)(11 * m + 8) * 64 + (size_t)(5 * n + 9) * 32; }
size_t oc_prover_randomness_len(int m, int n) { return (size_t)11 * m + 5 * n; }

/* mod.rs:388-395: out[i] = deck[perm[i]] + (rho_i*g, rho_i*pk) */
int oc_remask(const uint8_t* enc_g, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm, const uint8_t* rho,
              uint64_t Nc, uint8_t* out) {
  oracle_init();
  aff g, k;
  aff_from_bytes64(&g, enc_g);
  aff_from_bytes64(&k, pk);
#pragma omp parallel for num_threads(g_threads) if (g_threads > 1)
  for (uint64_t i = 0; i < Nc; i++) {
    fe r;
    fe_from_bytes(&r, rho + 32 * i, &FR);
    aff c1, c2;
    aff_from_bytes64(&c1, deck + 128 * (uint64_t)perm[i]);
    aff_from_bytes64(&c2, deck + 128 * (uint64_t)perm[i] + 64);
    jac a, b;
    aff_mul(&a, &g, &r);
    aff_mul(&b, &k, &r);
    jac_add_mixed(&a, &a, &c1);
    jac_add_mixed(&b, &b, &c2);
    jac_to_aff(&c1, &a);
    jac_to_aff(&c2, &b);
    aff_to_bytes64(out + 128 * i, &c1);
    aff_to_bytes64(out + 128 * i + 64, &c2);
  }
  return 0;
}

/* ShuffleArgument::prove (call site mod.rs:409-415) */
int oc_shuffle_prove(int m, int n, const uint8_t* enc_g, const uint8_t* ck_g, const uint8_t* ck_h, const uint8_t* ghat,
                     const uint8_t* pk, const uint8_t* deck_b, const uint8_t* deck2_b, const uint32_t* perm,
                     const uint8_t* rho_b, const uint8_t* rand_b, uint8_t* proof_out) {
  oracle_init();
  params_t pp;
  load_params(&pp, m, n, enc_g, ck_g, ck_h, ghat, pk);
  size_t Nc = (size_t)m * n, nr = oc_prover_randomness_len(m, n);
  ct_t* deck = load_deck(deck_b, Nc);
  ct_t* deck2 = load_deck(deck2_b, Nc);
  fe* rs = NEW(fe, nr);
  for (size_t i = 0; i < nr; i++) fe_from_bytes(&rs[i], rand_b + 32 * i, &FR);
  fe* rho = NEW(fe, Nc);
  for (size_t i = 0; i < Nc; i++) fe_from_bytes(&rho[i], rho_b + 32 * i, &FR);
  rand_t rd = {rs, 0};
  fsrng_t fs;
  fs_from_seed(&fs, (const uint8_t*)"Shuffle Proof", 13);
  wr_t out = {proof_out};

  fe* r = NEW(fe, m); fe* s = NEW(fe, m);
  rndv(&rd, r, m); rndv(&rd, s, m);
  fe* a = NEW(fe, Nc);
  for (size_t i = 0; i < Nc; i++) fe_from_u64(&a[i], (uint64_t)perm[i] + 1, &FR);
  aff* c_A = NEW(aff, m); aff* c_B = NEW(aff, m);
  for (int k = 0; k < m; k++) commit(&c_A[k], &pp, a + (size_t)k * n, n, &r[k]);
  absorb_statement(&fs, &pp, deck, deck2, Nc, c_A);
  fe x, y, z;
  fs_challenge(&fs, &x, &FR);
  fe* xp = NEW(fe, Nc + 1);
  powers(xp, &x, (int)Nc + 1);
  fe* b = NEW(fe, Nc);
  for (size_t i = 0; i < Nc; i++) b[i] = xp[perm[i] + 1];
  for (int k = 0; k < m; k++) commit(&c_B[k], &pp, b + (size_t)k * n, n, &s[k]);
  fs_absorb_begin(&fs); feed_label(&fs, "shuffle_argument_b"); feed_pts(&fs, c_B, m); fs_absorb_end(&fs);
  fs_challenge(&fs, &y, &FR);
  fs_challenge(&fs, &z, &FR);
  fe* d = NEW(fe, Nc); fe* t = NEW(fe, m);
  for (size_t i = 0; i < Nc; i++) { fe u; fe_mul(&u, &y, &a[i], &FR); fe_add(&u, &u, &b[i], &FR); fe_sub(&d[i], &u, &z, &FR); }
  for (int k = 0; k < m; k++) { fe u; fe_mul(&u, &y, &r[k], &FR); fe_add(&t[k], &u, &s[k], &FR); }
  aff* c_D = NEW(aff, m);
  fe bstar;
  product_statement(&pp, c_A, c_B, &x, &y, &z, Nc, c_D, &bstar);
  wr_pts(&out, c_A, m);
  wr_pts(&out, c_B, m);
  fe** drows = NEW(fe*, m); fe** brows = NEW(fe*, m);
  for (int k = 0; k < m; k++) { drows[k] = d + (size_t)k * n; brows[k] = b + (size_t)k * n; }
  product_prove(&pp, &fs, &rd, &out, m, c_D, drows, t);
  fe rho_star, u;
  fe_set_zero(&rho_star);
  for (size_t i = 0; i < Nc; i++) { fe_mul(&u, &rho[i], &b[i], &FR); fe_sub(&rho_star, &rho_star, &u, &FR); }
  ct_t Chat;
  ct_msm(&Chat, deck, xp + 1, Nc);
  multiexp_prove(&pp, &fs, &rd, &out, m, deck2, &Chat, brows, s, &rho_star);
  int ok = (rd.i == nr) && ((size_t)(out.p - proof_out) == oc_proof_len(m, n));
  free(deck); free(deck2); free(rs); free(rho); free(r); free(s); free(a); free(c_A); free(c_B); free(xp); free(b);
  free(d); free(t); free(c_D); free(drows); free(brows); free(pp.ck);
  return ok ? 0 : -1;
}

/* ShuffleArgument::verify (call site mod.rs:437-442); returns 0 or the failing check's code */
int oc_shuffle_verify(int m, int n, const uint8_t* enc_g, const uint8_t* ck_g, const uint8_t* ck_h, const uint8_t* ghat,
                      const uint8_t* pk, const uint8_t* deck_b, const uint8_t* deck2_b, const uint8_t* proof) {
  oracle_init();
  {
    /* `Proof: CanonicalDeserialize` (reference src/lib.rs:45-71): ark-serialize rejects a scalar >= the group order
     * before the verifier runs; -5 = MP_ERR_NOT_CANONICAL of include/mpshuffle.h.  Scalar runs of the flat layout:
     * 2n+3 after 5m+4 points, 2n+2 after 3 more points, n+4 after 6m+1 more points. */
    const size_t runs[3][2] = {{64 * (5 * (size_t)m + 4), 2 * (size_t)n + 3},
                               {64 * (5 * (size_t)m + 7) + 32 * (2 * (size_t)n + 3), 2 * (size_t)n + 2},
                               {64 * (11 * (size_t)m + 8) + 32 * (4 * (size_t)n + 5), (size_t)n + 4}};
    for (int r = 0; r < 3; r++)
      for (size_t k = 0; k < runs[r][1]; k++) {
        uint64_t c[4];
        memcpy(c, proof + runs[r][0] + 32 * k, 32);
        if (limbs_geq(c, FR.m)) return -5;
      }
  }
  params_t pp;
  load_params(&pp, m, n, enc_g, ck_g, ck_h, ghat, pk);
  size_t Nc = (size_t)m * n;
  ct_t* deck = load_deck(deck_b, Nc);
  ct_t* deck2 = load_deck(deck2_b, Nc);
  fsrng_t fs;
  fs_from_seed(&fs, (const uint8_t*)"Shuffle Proof", 13);
  rd_t in = {proof};
  aff* c_A = NEW(aff, m); aff* c_B = NEW(aff, m); aff* c_D = NEW(aff, m);
  rd_pts(&in, c_A, m);
  rd_pts(&in, c_B, m);
  absorb_statement(&fs, &pp, deck, deck2, Nc, c_A);
  fe x, y, z, bstar;
  fs_challenge(&fs, &x, &FR);
  fs_absorb_begin(&fs); feed_label(&fs, "shuffle_argument_b"); feed_pts(&fs, c_B, m); fs_absorb_end(&fs);
  fs_challenge(&fs, &y, &FR);
  fs_challenge(&fs, &z, &FR);
  product_statement(&pp, c_A, c_B, &x, &y, &z, Nc, c_D, &bstar);
  int st = product_verify(&pp, &fs, &in, m, c_D, &bstar);
  if (st == OK) {
    fe* xp = NEW(fe, Nc + 1);
    powers(xp, &x, (int)Nc + 1);
    ct_t Chat;
    ct_msm(&Chat, deck, xp + 1, Nc);
    st = multiexp_verify(&pp, &fs, &in, m, deck2, &Chat, c_B);
    free(xp);
  }
  free(deck); free(deck2); free(c_A); free(c_B); free(c_D); free(pp.ck);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * primitive exports (parity targets for the CUDA primitives + CPU baseline of the MSM bench)
 * ---------------------------------------------------------------------------------------- */
/* mode 0: per-term double-and-add; mode 1: ark-ec 0.3 Pippenger with its own window;
 * mode 2..: Pippenger with window = mode.  ncomp = 1 (G1) or 2 (ciphertexts). */
int oc_msm(const uint8_t* points, const uint8_t* scalars, uint64_t n, int ncomp, int mode, uint8_t* out) {
  oracle_init();
  aff* p = NEW(aff, n);
  uint64_t(*k)[4] = (uint64_t(*)[4])malloc(32 * (n ? n : 1));
  for (uint64_t i = 0; i < n; i++) {
    memcpy(k[i], scalars + 32 * i, 32);
    while (limbs_geq(k[i], FR.m)) limbs_sub(k[i], k[i], FR.m);
  }
  for (int comp = 0; comp < ncomp; comp++) {
    for (uint64_t i = 0; i < n; i++) aff_from_bytes64(&p[i], points + 64 * (i * ncomp + comp));
    jac j;
    if (mode == 0) msm_naive(&j, p, (const uint64_t(*)[4])k, n);
    else msm_pippenger(&j, p, (const uint64_t(*)[4])k, n, mode == 1 ? 0 : mode);
    aff a;
    jac_to_aff(&a, &j);
    aff_to_bytes64(out + 64 * comp, &a);
  }
  free(p); free(k);
  return 0;
}
/* out[i] = scalars[i] * base  (builds synthetic instances for the full-size CPU legs of bench.py and the tests) */
int oc_scalar_mul_batch(const uint8_t* base, const uint8_t* scalars, uint64_t n, uint8_t* out) {
  oracle_init();
  aff b;
  aff_from_bytes64(&b, base);
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1)
  for (uint64_t i = 0; i < n; i++) {
    fe r;
    fe_from_bytes(&r, scalars + 32 * i, &FR);
    jac j;
    aff_mul(&j, &b, &r);
    aff a;
    jac_to_aff(&a, &j);
    aff_to_bytes64(out + 64 * i, &a);
  }
  return 0;
}
int oc_on_curve(const uint8_t* point) { oracle_init(); aff a; aff_from_bytes64(&a, point); return aff_on_curve(&a); }
int oc_pedersen_commit(int n, const uint8_t* ck_g, const uint8_t* ck_h, const uint8_t* values, int len, const uint8_t* r,
                       uint8_t* out) {
  oracle_init();
  params_t pp;
  uint8_t z[64] = {0};
  load_params(&pp, 1, n, z, ck_g, ck_h, z, z);
  fe* v = NEW(fe, len);
  for (int i = 0; i < len; i++) fe_from_bytes(&v[i], values + 32 * i, &FR);
  fe rr;
  fe_from_bytes(&rr, r, &FR);
  aff c;
  commit(&c, &pp, v, len, &rr);
  aff_to_bytes64(out, &c);
  free(v); free(pp.ck);
  return 0;
}
void oc_fr_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  oracle_init();
  fe x, y;
  fe_from_bytes(&x, a, &FR); fe_from_bytes(&y, b, &FR);
  fe_mul(&x, &x, &y, &FR);
  fe_to_bytes(out, &x, &FR);
}
void oc_fq_mul(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  oracle_init();
  fe x, y;
  fe_from_bytes(&x, a, &FQ); fe_from_bytes(&y, b, &FQ);
  fe_mul(&x, &x, &y, &FQ);
  fe_to_bytes(out, &x, &FQ);
}
void oc_blake2s(const uint8_t* in, uint64_t len, uint8_t* out) {
  blake2s_t S;
  b2s_init(&S);
  b2s_update(&S, in, len);
  b2s_final(&S, out);
}
/* first `count` challenges after absorbing `data` into a fresh "Shuffle Proof" transcript */
void oc_fs_challenges(const uint8_t* data, uint64_t len, int count, uint8_t* out) {
  oracle_init();
  fsrng_t fs;
  fs_from_seed(&fs, (const uint8_t*)"Shuffle Proof", 13);
  if (len) { fs_absorb_begin(&fs); fs_absorb_feed(&fs, data, len); fs_absorb_end(&fs); }
  for (int i = 0; i < count; i++) { fe c; fs_challenge(&fs, &c, &FR); fe_to_bytes(out + 32 * i, &c, &FR); }
}

/* ------------------------------------------------------------------------------------------
 * Sigma protocols either side of the shuffle: Schnorr key ownership
 * (reference mod.rs:132-165) and Chaum-Pedersen DL equality behind mask / remask / reveal
 * (mod.rs:182-354; seeds mod.rs:80-83).  Bodies restated as in oracle/py/sigma.py
 * [UPSTREAM-RECALL]; PARITY UNPINNED against upstream bytes.  Batch entry points, one proof per
 * item, OpenMP over items: the CPU baseline of the batched GPU entry points.
 * ------------------------------------------------------------------------------------------ */
static void pt_sub(aff* out, const aff* a, const aff* b) {
  aff nb = *b;
  if (!nb.inf) fe_neg(&nb.y, &nb.y, &FQ);
  pt_add(out, a, &nb);
}
static void cp_challenge(fe* c, const char* seed, const aff* g, const aff* h, const aff* s0, const aff* s1, const aff* a,
                         const aff* b) {
  fsrng_t fs;
  fs_from_seed(&fs, (const uint8_t*)seed, strlen(seed));
  fs_absorb_begin(&fs);
  feed_label(&fs, "chaum_pedersen");
  feed_pts(&fs, g, 1); feed_pts(&fs, h, 1); feed_pts(&fs, s0, 1); feed_pts(&fs, s1, 1); feed_pts(&fs, a, 1); feed_pts(&fs, b, 1);
  fs_absorb_end(&fs);
  fs_challenge(&fs, c, &FR);
}
/* proof = a (64) | b (64) | r (32) */
static void cp_prove(uint8_t* proof, const char* seed, const aff* g, const aff* h, const aff* s0, const aff* s1, const fe* x,
                     const fe* omega) {
  aff a, b;
  fe c, r;
  pt_mul(&a, g, omega);
  pt_mul(&b, h, omega);
  cp_challenge(&c, seed, g, h, s0, s1, &a, &b);
  fe_mul(&r, &c, x, &FR);
  fe_add(&r, &r, omega, &FR);
  aff_to_bytes64(proof, &a);
  aff_to_bytes64(proof + 64, &b);
  fe_to_bytes(proof + 128, &r, &FR);
}
static int cp_verify(const uint8_t* proof, const char* seed, const aff* g, const aff* h, const aff* s0, const aff* s1) {
  aff a, b, l, t, rr;
  fe c, r;
  aff_from_bytes64(&a, proof);
  aff_from_bytes64(&b, proof + 64);
  fe_from_bytes(&r, proof + 128, &FR);
  cp_challenge(&c, seed, g, h, s0, s1, &a, &b);
  pt_mul(&l, g, &r); pt_mul(&t, s0, &c); pt_add(&rr, &a, &t);
  if (!aff_eq(&l, &rr)) return 5;
  pt_mul(&l, h, &r); pt_mul(&t, s1, &c); pt_add(&rr, &b, &t);
  return aff_eq(&l, &rr) ? 0 : 5;
}

#define SIGMA_FOR(i, n) _Pragma("omp parallel for schedule(dynamic, 8) num_threads(g_threads)") for (int64_t i = 0; i < (int64_t)(n); i++)

/* mask: masked_i = (r_i*g, card_i + r_i*pk); proof of mod.rs:182-214 */
int oc_mask_batch(const uint8_t* g64, const uint8_t* pk64, const uint8_t* cards, const uint8_t* rs, const uint8_t* omegas,
                  uint64_t n, uint8_t* out_masked, uint8_t* out_proofs) {
  oracle_init();
  aff g, pk;
  aff_from_bytes64(&g, g64); aff_from_bytes64(&pk, pk64);
  SIGMA_FOR(i, n) {
    aff card, c1, t, c2;
    fe r, om;
    aff_from_bytes64(&card, cards + 64 * i);
    fe_from_bytes(&r, rs + 32 * i, &FR); fe_from_bytes(&om, omegas + 32 * i, &FR);
    pt_mul(&c1, &g, &r); pt_mul(&t, &pk, &r); pt_add(&c2, &card, &t);
    aff_to_bytes64(out_masked + 128 * i, &c1); aff_to_bytes64(out_masked + 128 * i + 64, &c2);
    cp_prove(out_proofs + 160 * i, "Masking Proof", &g, &pk, &c1, &t, &r, &om);
  }
  return 0;
}
int oc_verify_mask_batch(const uint8_t* g64, const uint8_t* pk64, const uint8_t* cards, const uint8_t* masked,
                         const uint8_t* proofs, uint64_t n, int32_t* statuses) {
  oracle_init();
  aff g, pk;
  aff_from_bytes64(&g, g64); aff_from_bytes64(&pk, pk64);
  SIGMA_FOR(i, n) {
    aff card, c1, c2, s1;
    aff_from_bytes64(&card, cards + 64 * i);
    aff_from_bytes64(&c1, masked + 128 * i); aff_from_bytes64(&c2, masked + 128 * i + 64);
    pt_sub(&s1, &c2, &card);
    statuses[i] = cp_verify(proofs + 160 * i, "Masking Proof", &g, &pk, &c1, &s1);
  }
  return 0;
}
/* remask (no permutation: the reference's per-card remask, mod.rs:242-272) */
int oc_remask_prove_batch(const uint8_t* g64, const uint8_t* pk64, const uint8_t* deck, const uint8_t* alphas,
                          const uint8_t* omegas, uint64_t n, uint8_t* out_deck, uint8_t* out_proofs) {
  oracle_init();
  aff g, pk;
  aff_from_bytes64(&g, g64); aff_from_bytes64(&pk, pk64);
  SIGMA_FOR(i, n) {
    aff c1, c2, s0, s1, o1, o2;
    fe al, om;
    aff_from_bytes64(&c1, deck + 128 * i); aff_from_bytes64(&c2, deck + 128 * i + 64);
    fe_from_bytes(&al, alphas + 32 * i, &FR); fe_from_bytes(&om, omegas + 32 * i, &FR);
    pt_mul(&s0, &g, &al); pt_mul(&s1, &pk, &al);
    pt_add(&o1, &c1, &s0); pt_add(&o2, &c2, &s1);
    aff_to_bytes64(out_deck + 128 * i, &o1); aff_to_bytes64(out_deck + 128 * i + 64, &o2);
    cp_prove(out_proofs + 160 * i, "Remasking Proof", &g, &pk, &s0, &s1, &al, &om);
  }
  return 0;
}
int oc_verify_remask_batch(const uint8_t* g64, const uint8_t* pk64, const uint8_t* deck, const uint8_t* remasked,
                           const uint8_t* proofs, uint64_t n, int32_t* statuses) {
  oracle_init();
  aff g, pk;
  aff_from_bytes64(&g, g64); aff_from_bytes64(&pk, pk64);
  SIGMA_FOR(i, n) {
    aff c1, c2, o1, o2, s0, s1;
    aff_from_bytes64(&c1, deck + 128 * i); aff_from_bytes64(&c2, deck + 128 * i + 64);
    aff_from_bytes64(&o1, remasked + 128 * i); aff_from_bytes64(&o2, remasked + 128 * i + 64);
    pt_sub(&s0, &o1, &c1); pt_sub(&s1, &o2, &c2);
    statuses[i] = cp_verify(proofs + 160 * i, "Remasking Proof", &g, &pk, &s0, &s1);
  }
  return 0;
}
/* reveal tokens of ONE player for n masked cards (mod.rs:301-328) */
int oc_reveal_batch(const uint8_t* g64, const uint8_t* sk32, const uint8_t* pk64, const uint8_t* masked, const uint8_t* omegas,
                    uint64_t n, uint8_t* out_tokens, uint8_t* out_proofs) {
  oracle_init();
  aff g, pk;
  fe sk;
  aff_from_bytes64(&g, g64); aff_from_bytes64(&pk, pk64);
  fe_from_bytes(&sk, sk32, &FR);
  SIGMA_FOR(i, n) {
    aff c1, tok;
    fe om;
    aff_from_bytes64(&c1, masked + 128 * i);
    fe_from_bytes(&om, omegas + 32 * i, &FR);
    pt_mul(&tok, &c1, &sk);
    aff_to_bytes64(out_tokens + 64 * i, &tok);
    cp_prove(out_proofs + 160 * i, "Reveal Proof", &c1, &g, &tok, &pk, &sk, &om);
  }
  return 0;
}
int oc_verify_reveal_batch(const uint8_t* g64, const uint8_t* pk64, const uint8_t* tokens, const uint8_t* masked,
                           const uint8_t* proofs, uint64_t n, int32_t* statuses) {
  oracle_init();
  aff g, pk;
  aff_from_bytes64(&g, g64); aff_from_bytes64(&pk, pk64);
  SIGMA_FOR(i, n) {
    aff c1, tok;
    aff_from_bytes64(&c1, masked + 128 * i);
    aff_from_bytes64(&tok, tokens + 64 * i);
    statuses[i] = cp_verify(proofs + 160 * i, "Reveal Proof", &c1, &g, &tok, &pk);
  }
  return 0;
}
/* Schnorr key ownership (mod.rs:132-165): seed = "Key Ownership Proof" || info_i;
 * proof = commit (64) | opening (32), opening = omega - c*sk */
static void schnorr_challenge(fe* c, const uint8_t* info, size_t info_len, const aff* g, const aff* pk, const aff* commit) {
  static const char kSeed[] = "Key Ownership Proof";
  uint8_t* seed = (uint8_t*)malloc(sizeof(kSeed) - 1 + info_len + 1);
  memcpy(seed, kSeed, sizeof(kSeed) - 1);
  memcpy(seed + sizeof(kSeed) - 1, info, info_len);
  fsrng_t fs;
  fs_from_seed(&fs, seed, sizeof(kSeed) - 1 + info_len);
  free(seed);
  fs_absorb_begin(&fs);
  feed_label(&fs, "schnorr_identity");
  feed_pts(&fs, g, 1); feed_pts(&fs, pk, 1); feed_pts(&fs, commit, 1);
  fs_absorb_end(&fs);
  fs_challenge(&fs, c, &FR);
}
int oc_key_ownership_prove_batch(const uint8_t* g64, const uint8_t* pks, const uint8_t* sks, const uint8_t* infos,
                                 const uint64_t* info_off, const uint8_t* omegas, uint64_t n, uint8_t* out_proofs) {
  oracle_init();
  aff g;
  aff_from_bytes64(&g, g64);
  SIGMA_FOR(i, n) {
    aff pk, commit;
    fe sk, om, c, op;
    aff_from_bytes64(&pk, pks + 64 * i);
    fe_from_bytes(&sk, sks + 32 * i, &FR); fe_from_bytes(&om, omegas + 32 * i, &FR);
    pt_mul(&commit, &g, &om);
    schnorr_challenge(&c, infos + info_off[i], (size_t)(info_off[i + 1] - info_off[i]), &g, &pk, &commit);
    fe_mul(&op, &c, &sk, &FR);
    fe_sub(&op, &om, &op, &FR);
    aff_to_bytes64(out_proofs + 96 * i, &commit);
    fe_to_bytes(out_proofs + 96 * i + 64, &op, &FR);
  }
  return 0;
}
int oc_key_ownership_verify_batch(const uint8_t* g64, const uint8_t* pks, const uint8_t* infos, const uint64_t* info_off,
                                  const uint8_t* proofs, uint64_t n, int32_t* statuses) {
  oracle_init();
  aff g;
  aff_from_bytes64(&g, g64);
  SIGMA_FOR(i, n) {
    aff pk, commit, l, t, s;
    fe op, c;
    aff_from_bytes64(&pk, pks + 64 * i);
    aff_from_bytes64(&commit, proofs + 96 * i);
    fe_from_bytes(&op, proofs + 96 * i + 64, &FR);
    schnorr_challenge(&c, infos + info_off[i], (size_t)(info_off[i + 1] - info_off[i]), &g, &pk, &commit);
    pt_mul(&l, &g, &op); pt_mul(&t, &pk, &c); pt_add(&s, &l, &t);
    statuses[i] = aff_eq(&s, &commit) ? 0 : 6;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * ark-serialize 0.3 wire format [UPSTREAM-RECALL]:
 * compressed SW affine = x (32 B LE) | flags in the top bits of the last byte (bit 7: y is the larger
 * of (y, -y); bit 6: infinity).  Decompression = Tonelli-Shanks over F_p, p - 1 = 2^192 * (2^59 + 17),
 * the algorithm ark-ff 0.3 runs for `sqrt` [UPSTREAM-RECALL].  As in oracle/py/wire.py.
 * ------------------------------------------------------------------------------------------ */
static fe SQRT_ROOT; /* 3^t, a primitive 2^192-th root of unity (Montgomery form) */
static int g_sqrt_inited = 0;
static void sqrt_init(void) {
  if (g_sqrt_inited) return;
  static const uint64_t T[4] = {0x0800000000000011ull, 0, 0, 0}; /* t = (p - 1) / 2^192 */
  fe three;
  fe_from_u64(&three, 3, &FQ);
  fe_pow(&SQRT_ROOT, &three, T, &FQ);
  g_sqrt_inited = 1;
}
/* returns 1 and a root in *r, or 0 if a is a non-residue */
static int fe_sqrt(fe* r, const fe* a) {
  if (fe_is_zero(a)) { fe_set_zero(r); return 1; }
  static const uint64_t E[4] = {0x0400000000000008ull, 0, 0, 0}; /* (t - 1) / 2 */
  fe w, x, b, z = SQRT_ROOT;
  fe_pow(&w, a, E, &FQ);
  fe_mul(&x, a, &w, &FQ);
  fe_mul(&b, &x, &w, &FQ);
  int v = 192;
  while (!fe_eq(&b, &FQ.one)) {
    int k = 0;
    fe t2 = b;
    while (!fe_eq(&t2, &FQ.one)) {
      fe_sqr(&t2, &t2, &FQ);
      if (++k == v) return 0;
    }
    fe ww = z;
    for (int i = 0; i < v - k - 1; i++) fe_sqr(&ww, &ww, &FQ);
    fe_sqr(&z, &ww, &FQ);
    fe_mul(&b, &b, &z, &FQ);
    fe_mul(&x, &x, &ww, &FQ);
    v = k;
  }
  *r = x;
  return 1;
}
/* canonical y > p - y ?  <=>  y > (p - 1) / 2 */
static int y_is_larger(const fe* y) {
  static const uint64_t HALF[4] = {0, 0, 0x8000000000000000ull, 0x0400000000000008ull};
  uint64_t c[4];
  fe_to_raw(c, y, &FQ);
  for (int i = 3; i >= 0; i--)
    if (c[i] != HALF[i]) return c[i] > HALF[i];
  return 0;
}
int oc_points_compress(const uint8_t* points, uint64_t n, uint8_t* out) {
  oracle_init();
  for (uint64_t i = 0; i < n; i++) {
    aff a;
    aff_from_bytes64(&a, points + 64 * i);
    uint8_t* o = out + 32 * i;
    if (a.inf) { memset(o, 0, 32); o[31] = 0x40; continue; }
    fe_to_bytes(o, &a.x, &FQ);
    if (y_is_larger(&a.y)) o[31] |= 0x80;
  }
  return 0;
}
/* statuses[i] = 0 ok, 1 malformed (non-canonical x / stray flags), 2 x not on the curve */
int oc_points_decompress(const uint8_t* in, uint64_t n, uint8_t* out, int32_t* statuses) {
  oracle_init();
  sqrt_init();
  _Pragma("omp parallel for schedule(dynamic, 8) num_threads(g_threads)")
  for (int64_t i = 0; i < (int64_t)n; i++) {
    uint8_t xb[32];
    memcpy(xb, in + 32 * i, 32);
    const int flags = xb[31] & 0xC0;
    xb[31] &= 0x3F;
    uint8_t* o = out + 64 * i;
    memset(o, 0, 64);
    statuses[i] = 0;
    int zero = 1;
    for (int k = 0; k < 32; k++) zero &= xb[k] == 0;
    if (flags & 0x40) { statuses[i] = (zero && !(flags & 0x80)) ? 0 : 1; continue; }
    /* canonical: x < p */
    static const uint8_t PB[32] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x11, 0, 0, 0, 0, 0, 0, 0x08};
    int lt = 0;
    for (int k = 31; k >= 0; k--)
      if (xb[k] != PB[k]) { lt = xb[k] < PB[k]; break; }
    if (!lt) { statuses[i] = 1; continue; }
    fe x, rhs, y;
    fe_from_bytes(&x, xb, &FQ);
    fe_sqr(&rhs, &x, &FQ);
    fe_mul(&rhs, &rhs, &x, &FQ);
    fe_add(&rhs, &rhs, &x, &FQ);
    fe_add(&rhs, &rhs, &CURVE_B, &FQ);
    if (!fe_sqrt(&y, &rhs)) { statuses[i] = 2; continue; }
    if (y_is_larger(&y) != ((flags & 0x80) != 0)) fe_neg(&y, &y, &FQ);
    fe_to_bytes(o, &x, &FQ);
    fe_to_bytes(o + 32, &y, &FQ);
  }
  return 0;
}
