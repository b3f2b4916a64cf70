// Host-only half of the shuffle-argument driver: proof layout, transcript schedule, the
// verifier's rewriting of every check into "sum scalar * point == identity" jobs, and the verdict.
// No CUDA in this header -- tests/host/host_shim.cpp compiles it with g++ so the algebra can be
// checked against the oracle on a machine without a GPU (tests/test_host_verify_plan.py); the
// product includes it from the shuffle_*.cu files.
//
// Reference anchors: DLCards::verify_shuffle / shuffle_and_remask call sites
// (reference src/discrete_log_cards/mod.rs:380-443), transcript seed mod.rs:84,408,436; algebra
// and transcript order: SURVEY.md Appendix B (B.6 for the byte / draw order).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/mpshuffle.h"
#include "fr.cuh"
#include "msm_job.h"
#include "transcript.hpp"

namespace mp {

// Byte widths of the curve fq.cuh selects: a C-ABI point is x || y (64 bytes on the Stark curve, 96 on
// BLS12-377 G1), a ciphertext two points; scalars are 32 bytes on both.  Everything below is written
// against these, so this header compiles for either instantiation (tests/test_host_verify_plan.py runs both).
static constexpr size_t kPointBytes = 8 * (size_t)kFqLimbs;
static constexpr size_t kCtBytes = 2 * kPointBytes;

// DLCards `Parameters` as the host sees them (reference mod.rs:37-61)
struct ShuffleParamsHost {
  int m = 0, n = 0;
  std::vector<uint8_t> ck64;  // (n+1) points canonical: h, g_1 .. g_n   (MSM order of a commitment)
  uint8_t enc_g[kPointBytes], ghat[kPointBytes], gsum[kPointBytes];  // gsum = g_1 + .. + g_n  (com(c,..,c; 0) = c * gsum)
};

inline uint64_t shuffle_proof_len(int32_t m, int32_t n) { return (uint64_t)(11 * m + 8) * kPointBytes + (uint64_t)(5 * n + 9) * 32; }
inline uint64_t shuffle_randomness_len(int32_t m, int32_t n) { return (uint64_t)11 * m + (uint64_t)5 * n; }

// ------------------------------------------------------------------------------------------
// host-side scalar helpers
// ------------------------------------------------------------------------------------------
inline fr h_fr(const uint8_t* b) {
  uint32_t w[8];
  memcpy(w, b, 32);
  return fr_from_canonical(w);
}
// canonical 32-byte scalar, i.e. below the group order?  (ark-serialize rejects anything else at deserialisation)
inline bool h_fr_is_canonical(const uint8_t* b) {
  uint32_t w[8];
  memcpy(w, b, 32);
  for (int i = 7; i >= 0; i--) {
    const uint32_t mi = fr_modulus_limb(i);
    if (w[i] != mi) return w[i] < mi;
  }
  return false;
}
inline void h_fr_out(const fr& a, uint8_t* b) {
  uint32_t w[8];
  fr_to_canonical(a, w);
  memcpy(b, w, 32);
}
inline std::vector<fr> h_powers(const fr& x, int count) {  // x^0 .. x^(count-1)
  std::vector<fr> p((size_t)std::max(count, 0));
  if (count > 0) p[0] = fr_one();
  for (int k = 1; k < count; k++) p[k] = fr_mul(p[k - 1], x);
  return p;
}
inline std::vector<fr> h_frs(const uint8_t* b, int count) {
  std::vector<fr> v((size_t)count);
  for (int i = 0; i < count; i++) v[i] = h_fr(b + 32 * (size_t)i);
  return v;
}
inline fr h_dot(const fr* a, const fr* b, int n) {
  fr acc = fr_zero();
  for (int i = 0; i < n; i++) acc = fr_add(acc, fr_mul(a[i], b[i]));
  return acc;
}
inline bool all_zero(const uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (p[i]) return false;
  return true;
}

// A batch of small G1 MSM jobs assembled on the host: job = list of (point, scalar) terms.
struct TermList {
  std::vector<uint8_t> pts;    // one canonical point (kPointBytes) per term
  std::vector<uint32_t> scal;  // 8 words canonical per term
  std::vector<MsmJob> jobs;
  uint32_t start = 0;
  uint32_t count() const { return (uint32_t)(scal.size() / 8); }
  void term(const uint8_t* p64, const fr& s) {
    pts.insert(pts.end(), p64, p64 + kPointBytes);
    uint32_t w[8];
    fr_to_canonical(s, w);
    scal.insert(scal.end(), w, w + 8);
  }
  void close_job() {
    jobs.push_back(MsmJob{start, start, count() - start});
    start = count();
  }
};

// ------------------------------------------------------------------------------------------
// proof layout (include/mpshuffle.h)
// ------------------------------------------------------------------------------------------
struct Layout {
  size_t cA, cB, cb, hB, zpts, za, zb, zr, zs, zt, svpts, sva, svb, svr, svs, mepts, meE, mea, mer, meb, mes, metau, end;
  Layout(int m, int n) {
    const size_t P = kPointBytes, F = 32;
    cA = 0; cB = cA + m * P; cb = cB + m * P; hB = cb + P; zpts = hB + m * P;
    za = zpts + (2 * (size_t)m + 3) * P; zb = za + n * F; zr = zb + n * F; zs = zr + F; zt = zs + F;
    svpts = zt + F; sva = svpts + 3 * P; svb = sva + n * F; svr = svb + n * F; svs = svr + F;
    mepts = svs + F; meE = mepts + (2 * (size_t)m + 1) * P; mea = meE + 4 * (size_t)m * P;
    mer = mea + n * F; meb = mer + F; mes = meb + F; metau = mes + F; end = metau + F;
  }
};

// Every one of the 5n + 9 scalars of a proof must be a canonical residue: h_fr() would reduce s + order to s
// silently and the proof would still verify (malleability); the reference never sees such a proof because
// `Proof: CanonicalDeserialize` (bounds at reference src/lib.rs:45-71) rejects the bytes.
inline bool proof_scalars_canonical(const uint8_t* proof, const Layout& L) {
  const size_t runs[3][2] = {{L.za, L.svpts}, {L.sva, L.mepts}, {L.mea, L.end}};
  for (auto& r : runs)
    for (size_t off = r[0]; off < r[1]; off += 32)
      if (!h_fr_is_canonical(proof + off)) return false;
  return true;
}

// The statement absorb is ONE Blake2s over  label | parameters | pk | deck | deck' | c_A | seed  --
// serial by construction and, at 2^16 cards, 17 MB.  It is fed in two parts so that the prover can
// hash what it already has (everything up to the input deck, then the shuffled deck) while the GPU
// is still remasking / committing, and only then wait for c_A.
inline void absorb_statement_head(Transcript& fs, const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* deck, size_t N) {
  fs.begin();
  fs.feed_label("shuffle_argument");
  fs.feed_points64(S->enc_g, 1);
  fs.feed_points64(pk, 1);
  fs.feed_points64(S->ck64.data() + kPointBytes, (size_t)S->n);
  fs.feed_points64(S->ck64.data(), 1);
  fs.feed_points64(S->ghat, 1);
  fs.feed_points64(deck, 2 * N);
}
inline void absorb_statement_deck2(Transcript& fs, const uint8_t* deck2, size_t N) { fs.feed_points64(deck2, 2 * N); }
// the head for up to 8 proofs of a batch at once (same parameters and key, decks[l] = the input deck of lane l)
inline void absorb_statement_head_lanes(TranscriptLanes& tl, const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* const* decks,
                                        size_t N) {
  tl.feed_label_all("shuffle_argument");
  tl.feed_points64_all(S->enc_g, 1);
  tl.feed_points64_all(pk, 1);
  tl.feed_points64_all(S->ck64.data() + kPointBytes, (size_t)S->n);
  tl.feed_points64_all(S->ck64.data(), 1);
  tl.feed_points64_all(S->ghat, 1);
  tl.feed_points64(decks, 2 * N);
}
inline void absorb_statement_tail(Transcript& fs, const uint8_t* cA, int m) {
  fs.feed_points64(cA, (size_t)m);
  fs.end();
}
inline void absorb_statement(Transcript& fs, const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* deck,
                             const uint8_t* deck2, size_t N, const uint8_t* cA) {
  absorb_statement_head(fs, S, pk, deck, N);
  absorb_statement_deck2(fs, deck2, N);
  absorb_statement_tail(fs, cA, S->m);
}

// ------------------------------------------------------------------------------------------
// verifier pieces shared by the single-proof and the batched entry points
// ------------------------------------------------------------------------------------------
struct Challenges {
  fr x, y, z, xh, yh, xz, xs, xm;
};

// Every challenge derives from statement + proof bytes (no device round trip).
// (stmt_started: a hasher that has already absorbed the statement up to and including the shuffled deck -- the batch
//  verifier hashes those 17 MB for several proofs at once, shuffle_internal.cuh StatementHashes)
inline Challenges derive_challenges(const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                                    size_t N, const uint8_t* proof, const Layout& L, const Blake2s* stmt_started = nullptr) {
  const int m = S->m;
  Challenges ch;
  Transcript fs;
  if (stmt_started) {
    fs.adopt_pending(*stmt_started);
    absorb_statement_tail(fs, proof + L.cA, m);
  } else {
    absorb_statement(fs, S, pk, deck, deck2, N, proof + L.cA);
  }
  ch.x = fs.challenge();
  fs.begin(); fs.feed_label("shuffle_argument_b"); fs.feed_points64(proof + L.cB, m); fs.end();
  ch.y = fs.challenge();
  ch.z = fs.challenge();
  fs.begin(); fs.feed_label("hadamard_argument"); fs.feed_points64(proof + L.cb, 1); fs.feed_points64(proof + L.hB, m); fs.end();
  ch.xh = fs.challenge();
  ch.yh = fs.challenge();
  fs.begin(); fs.feed_label("zero_argument"); fs.feed_points64(proof + L.zpts, 2 * (size_t)m + 3); fs.end();
  ch.xz = fs.challenge();
  fs.begin(); fs.feed_label("single_value_product_argument"); fs.feed_points64(proof + L.svpts, 3); fs.end();
  ch.xs = fs.challenge();
  fs.begin(); fs.feed_label("multi_exponentiation_argument");
  fs.feed_points64(proof + L.mepts, 2 * (size_t)m + 1); fs.feed_points64(proof + L.meE, 4 * (size_t)m); fs.end();
  ch.xm = fs.challenge();
  return ch;
}

// What the host still has to compare once the device reports which jobs are the identity.
struct HostChecks {
  bool hadamard_bytes_ok;   // c_B'[m-1] == c_b
  bool zero_bytes_ok;       // zero-argument c_D[m+1] == O
  bool svp_first_ok;        // b~_1 == a~_1
  fr svp_last, xs;          // b~_n must equal xs * bstar
  bool multiexp_bytes_ok;   // multi-exp c_B[m] == O
};
static const int kG1Checks = 8;  // H1, Z1, Z2, Z3, S1, S2, M1, M2 -- one MSM job each

// Appends the eight commitment-space equations of the verifier as "sum scalar*point == O" jobs.
inline void append_g1_checks(TermList& tl, const ShuffleParamsHost* S, const uint8_t* proof, const Layout& L,
                             const Challenges& ch, HostChecks* hc) {
  const int m = S->m, n = S->n;
  const uint8_t* ck_h = S->ck64.data();
  auto ck_g = [&](int j) { return S->ck64.data() + kPointBytes * (size_t)(j + 1); };  // g_{j+1}, j = 0..n-1
  auto P = [&](size_t off, size_t i) { return proof + off + kPointBytes * i; };
  const fr &y = ch.y, &z = ch.z, &xh = ch.xh, &yh = ch.yh, &xs = ch.xs;
  const std::vector<fr> xzp = h_powers(ch.xz, 2 * m + 1);
  const std::vector<fr> xhp = h_powers(xh, m);
  const std::vector<fr> xmp = h_powers(ch.xm, 2 * m);
  const std::vector<fr> z_a = h_frs(proof + L.za, n), z_b = h_frs(proof + L.zb, n);
  const fr z_r = h_fr(proof + L.zr), z_s = h_fr(proof + L.zs), z_t = h_fr(proof + L.zt);
  const std::vector<fr> sv_a = h_frs(proof + L.sva, n), sv_b = h_frs(proof + L.svb, n);
  const fr sv_r = h_fr(proof + L.svr), sv_s = h_fr(proof + L.svs);
  const std::vector<fr> me_a = h_frs(proof + L.mea, n);
  const fr me_r = h_fr(proof + L.mer), me_b = h_fr(proof + L.meb), me_s = h_fr(proof + L.mes);
  const fr one = fr_one();
  // H1: hB[0] == c_D[0] = y*c_A[0] + c_B[0] - z*gsum
  tl.term(P(L.cA, 0), y); tl.term(P(L.cB, 0), one); tl.term(S->gsum, fr_neg(z)); tl.term(P(L.hB, 0), fr_neg(one));
  tl.close_job();
  // Z1: c_A0 + sum_{i=1}^{m-1} xz^i c_D[i] + xz^m (-gsum) - com(a; r)
  {
    fr s1 = fr_zero();
    tl.term(P(L.zpts, 0), one);
    for (int i = 1; i < m; i++) {
      tl.term(P(L.cA, i), fr_mul(xzp[i], y));
      tl.term(P(L.cB, i), xzp[i]);
      s1 = fr_add(s1, xzp[i]);
    }
    tl.term(S->gsum, fr_neg(fr_add(fr_mul(z, s1), xzp[m])));
    tl.term(ck_h, fr_neg(z_r));
    for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(z_a[j]));
    tl.close_job();
  }
  // Z2: sum_{t=0}^{m-2} xz^{m-t} xh^{t+1} hB[t] + xz * sum_{i=1}^{m-1} xh^i hB[i] + c_Bm1 - com(b; s)
  {
    for (int t = 0; t < m; t++) {
      fr c = fr_zero();
      if (t <= m - 2) c = fr_add(c, fr_mul(xzp[m - t], fr_mul(xhp[t], xh)));
      if (t >= 1) c = fr_add(c, fr_mul(xzp[1], xhp[t]));
      tl.term(P(L.hB, t), c);
    }
    tl.term(P(L.zpts, 1), one);
    tl.term(ck_h, fr_neg(z_s));
    for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(z_b[j]));
    tl.close_job();
  }
  // Z3: sum xz^k c_D_k - com(a * b; t)
  {
    fr ab = fr_zero(), yp = one;
    for (int j = 0; j < n; j++) {
      yp = fr_mul(yp, yh);
      ab = fr_add(ab, fr_mul(fr_mul(z_a[j], z_b[j]), yp));
    }
    for (int k = 0; k <= 2 * m; k++) tl.term(P(L.zpts, 2 + k), xzp[k]);
    tl.term(ck_h, fr_neg(z_t));
    tl.term(ck_g(0), fr_neg(ab));
    tl.close_job();
  }
  // S1: xs*c_b + c_d - com(a~; r~)
  tl.term(P(L.cb, 0), xs); tl.term(P(L.svpts, 0), one); tl.term(ck_h, fr_neg(sv_r));
  for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(sv_a[j]));
  tl.close_job();
  // S2: xs*c_Delta + c_delta - com((xs*b~_{i+1} - b~_i*a~_{i+1})_i; s~)
  tl.term(P(L.svpts, 2), xs); tl.term(P(L.svpts, 1), one); tl.term(ck_h, fr_neg(sv_s));
  for (int i = 0; i + 1 < n; i++)
    tl.term(ck_g(i), fr_neg(fr_sub(fr_mul(xs, sv_b[i + 1]), fr_mul(sv_b[i], sv_a[i + 1]))));
  tl.close_job();
  // M1: c_A0 + sum_{j=1}^{m} xm^j c_B[j-1] - com(a; r)
  tl.term(P(L.mepts, 0), one);
  for (int j = 1; j <= m; j++) tl.term(P(L.cB, j - 1), xmp[j]);
  tl.term(ck_h, fr_neg(me_r));
  for (int j = 0; j < n; j++) tl.term(ck_g(j), fr_neg(me_a[j]));
  tl.close_job();
  // M2: sum xm^k c_B_k - com(b; s)
  for (int k = 0; k < 2 * m; k++) tl.term(P(L.mepts, 1 + k), xmp[k]);
  tl.term(ck_h, fr_neg(me_s));
  tl.term(ck_g(0), fr_neg(me_b));
  tl.close_job();
  hc->hadamard_bytes_ok = memcmp(P(L.hB, m - 1), P(L.cb, 0), kPointBytes) == 0;
  hc->zero_bytes_ok = all_zero(P(L.zpts, 2 + m + 1), kPointBytes);
  hc->svp_first_ok = fr_eq(sv_b[0], sv_a[0]);
  hc->svp_last = sv_b[n - 1];
  hc->xs = xs;
  hc->multiexp_bytes_ok = all_zero(P(L.mepts, 1 + m), kPointBytes);
}

// The two ciphertext equations of the verifier for ONE proof, with host-computed scalars (the
// small-deck / batched path; the 2^16-card path fills the same arena with device kernels):
//   eq 0:  sum_i x^i C_i  +  (-1) E_m                                              == O   (Chat == E_m)
//   eq 1:  sum_ij -(xm^{m-i} a_j) C'_ij  +  sum_k xm^k E_k - tau (g, pk) - b (O, ghat) == O
// sx, s2: N canonical scalars for deck / shuffled deck; ss: 2m + 3 scalars for the small
// ciphertext points written to small_pts (kCtBytes each): E_m | E_0..E_{2m-1} | (g, pk) | (O, ghat).
// Also returns bstar = prod_{i=1..N} (y i + x^i - z), the product-argument statement.
inline void build_ct_plan(const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* proof, const Layout& L,
                          const Challenges& ch, uint32_t* sx, uint32_t* s2, uint32_t* ss, uint8_t* small_pts, fr* bstar) {
  const int m = S->m, n = S->n;
  const size_t N = (size_t)m * n;
  fr xi = fr_one(), yi = fr_zero(), prod = fr_one();
  for (size_t i = 0; i < N; i++) {
    xi = fr_mul(xi, ch.x);
    yi = fr_add(yi, ch.y);
    prod = fr_mul(prod, fr_sub(fr_add(yi, xi), ch.z));
    fr_to_canonical(xi, sx + 8 * i);
  }
  *bstar = prod;
  const std::vector<fr> xmp = h_powers(ch.xm, 2 * m);
  const std::vector<fr> me_a = h_frs(proof + L.mea, n);
  for (int i = 1; i <= m; i++) {
    const fr cf = fr_neg(xmp[m - i]);
    for (int j = 0; j < n; j++) fr_to_canonical(fr_mul(cf, me_a[j]), s2 + 8 * ((size_t)(i - 1) * n + j));
  }
  fr_to_canonical(fr_neg(fr_one()), ss);
  for (int k = 0; k < 2 * m; k++) fr_to_canonical(xmp[k], ss + 8 * (size_t)(1 + k));
  fr_to_canonical(fr_neg(h_fr(proof + L.metau)), ss + 8 * (size_t)(1 + 2 * m));
  fr_to_canonical(fr_neg(h_fr(proof + L.meb)), ss + 8 * (size_t)(2 + 2 * m));
  memcpy(small_pts, proof + L.meE + kCtBytes * (size_t)m, kCtBytes);
  memcpy(small_pts + kCtBytes, proof + L.meE, 2 * (size_t)m * kCtBytes);
  uint8_t* t = small_pts + kCtBytes * (size_t)(1 + 2 * m);
  memcpy(t, S->enc_g, kPointBytes);
  memcpy(t + kPointBytes, pk, kPointBytes);
  memset(t + 2 * kPointBytes, 0, kPointBytes);
  memcpy(t + 3 * kPointBytes, S->ghat, kPointBytes);
}

// The same two equations laid out as TWO CONTIGUOUS 2-component MSM jobs over one ciphertext array and one scalar
// array (the form a batched-MSM entry point takes; used by the host-scalar verifier of capi_bls12_377.cu):
//   cts  = deck (N) | E_m | deck' (N) | E_0..E_{2m-1} | (g, pk) | (O, ghat)          2N + 2m + 3 ciphertexts
//   scal = x^i (N)  | -1  | -(xm^{m-i} a_j) (N) | xm^k .. | -tau | -b                  same order, 8 words each
//   jobs = (0, 0, N + 1), (N + 1, N + 1, N + 2m + 2)   as (scalar_off, point_off, len)
inline void assemble_ct_jobs(const ShuffleParamsHost* S, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2,
                             const uint8_t* proof, const Layout& L, const Challenges& ch, std::vector<uint8_t>& cts,
                             std::vector<uint32_t>& scal, uint32_t jobs[6], fr* bstar) {
  const size_t N = (size_t)S->m * S->n, nsmall = 2 * (size_t)S->m + 3, nct = 2 * N + nsmall;
  std::vector<uint8_t> small(nsmall * kCtBytes);
  std::vector<uint32_t> sx(8 * N), s2(8 * N), ss(8 * nsmall);
  build_ct_plan(S, pk, proof, L, ch, sx.data(), s2.data(), ss.data(), small.data(), bstar);
  cts.resize(nct * kCtBytes);
  scal.resize(8 * nct);
  memcpy(cts.data(), deck, N * kCtBytes);
  memcpy(cts.data() + N * kCtBytes, small.data(), kCtBytes);
  memcpy(cts.data() + (N + 1) * kCtBytes, deck2, N * kCtBytes);
  memcpy(cts.data() + (2 * N + 1) * kCtBytes, small.data() + kCtBytes, (nsmall - 1) * kCtBytes);
  memcpy(scal.data(), sx.data(), 32 * N);
  memcpy(scal.data() + 8 * N, ss.data(), 32);
  memcpy(scal.data() + 8 * (N + 1), s2.data(), 32 * N);
  memcpy(scal.data() + 8 * (2 * N + 1), ss.data() + 8, 32 * (nsmall - 1));
  jobs[0] = 0; jobs[1] = 0; jobs[2] = (uint32_t)(N + 1);
  jobs[3] = (uint32_t)(N + 1); jobs[4] = (uint32_t)(N + 1); jobs[5] = (uint32_t)(N + nsmall - 1);
}

// Verdict in the order the reference reaches the checks (product argument first: Hadamard ->
// zero -> single-value product; then multi-exponentiation).  g1_id[0..8): H1 Z1 Z2 Z3 S1 S2 M1 M2;
// ct_ok: both ciphertext equations (Chat == E_m and the multi-exp opening) hold.
inline int32_t verdict(const HostChecks& hc, const fr& bstar, const bool* g1_id, bool ct_ok) {
  if (!g1_id[0] || !hc.hadamard_bytes_ok) return MP_VERIFY_HADAMARD;
  if (!hc.zero_bytes_ok || !g1_id[1] || !g1_id[2] || !g1_id[3]) return MP_VERIFY_ZERO;
  if (!g1_id[4] || !g1_id[5] || !hc.svp_first_ok || !fr_eq(hc.svp_last, fr_mul(hc.xs, bstar))) return MP_VERIFY_SVP;
  if (!hc.multiexp_bytes_ok || !ct_ok || !g1_id[6] || !g1_id[7]) return MP_VERIFY_MULTIEXP;
  return MP_OK;
}

// flat prover randomness, consumed in the order of SURVEY.md Appendix B.6 (include/mpshuffle.h)
struct RandCursor {
  const uint8_t* p;
  size_t i = 0;
  fr one() { return h_fr(p + 32 * (i++)); }
  std::vector<fr> vec(int k) {
    std::vector<fr> v((size_t)k);
    for (int j = 0; j < k; j++) v[j] = one();
    return v;
  }
};

}  // namespace mp
