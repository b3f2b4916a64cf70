// TEST INFRASTRUCTURE ONLY: the __host__ side of csrc/fq_bls12_377.cuh + ec.cuh (compiled with
// -DMP_CURVE_BLS12_377) under g++, so the 12-limb Montgomery arithmetic, the lazy bounds and the XYZZ
// formulas with a = 0 are checked against the Python oracle without a GPU.  Not linked into the product.
#include "../../mental-poker_b200/csrc/ec.cuh"
#include <string.h>
using namespace mp;
static_assert(kFqLimbs == 12, "build with -DMP_CURVE_BLS12_377");
extern "C" {
void h_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  fq x, y; memcpy(x.v, a, 48); memcpy(y.v, b, 48);
  fq r = fq_mul(x, y); memcpy(out, r.v, 48);
}
void h_fq_sqr(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 48); fq r = fq_sqr(x); memcpy(out, r.v, 48);
}
void h_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t kb, uint32_t* out) {
  fq x, y; memcpy(x.v, a, 48); memcpy(y.v, b, 48);
  fq r = fq_sub(x, y, kb); memcpy(out, r.v, 48);
}
void h_fq_reduce_weak(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 48); fq r = fq_reduce_weak(x); memcpy(out, r.v, 48);
}
void h_fq_reduce_full(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 48); fq r = fq_reduce_full(x); memcpy(out, r.v, 48);
}
int h_fq_is_zero_mod_p_2(const uint32_t* a) { fq x; memcpy(x.v, a, 48); return fq_is_zero_mod_p_2(x); }
void h_fq_inv_canonical(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 48);
  fq r = fq_from_mont(fq_inv(fq_to_mont(x))); memcpy(out, r.v, 48);
}
int h_on_curve(const uint32_t* p) { return affine_on_curve(affine_from_canonical(p)); }
void h_point_add(const uint32_t* p, const uint32_t* q, uint32_t* out) {
  xyzz acc = xyzz_from_affine(affine_from_canonical(p));
  xyzz_madd(acc, affine_from_canonical(q));
  affine_to_canonical(xyzz_to_affine(acc), out);
}
void h_scalar_mul(const uint32_t* p, const uint32_t* k, uint32_t* out) {
  affine P = affine_from_canonical(p);
  xyzz acc = xyzz_identity();
  for (int i = 255; i >= 0; i--) {
    acc = xyzz_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_madd(acc, P);
  }
  affine_to_canonical(xyzz_to_affine(acc), out);
}
void h_lincomb2(const uint32_t* p, const uint32_t* k1, const uint32_t* q, const uint32_t* k2, uint32_t* out) {
  affine P = affine_from_canonical(p), Q = affine_from_canonical(q);
  xyzz a = xyzz_identity(), b = xyzz_identity();
  for (int i = 255; i >= 0; i--) {
    a = xyzz_dbl(a); b = xyzz_dbl(b);
    if ((k1[i >> 5] >> (i & 31)) & 1) xyzz_madd(a, P);
    if ((k2[i >> 5] >> (i & 31)) & 1) xyzz_madd(b, Q);
  }
  xyzz_add(a, b);
  affine_to_canonical(xyzz_to_affine(a), out);
}
// 2 * P through the affine doubling entry (mdbl), which madd takes when both operands are equal
void h_dbl_affine(const uint32_t* p, uint32_t* out) {
  affine_to_canonical(xyzz_to_affine(xyzz_dbl_affine(affine_from_canonical(p))), out);
}
}

// ---- host-only verifier plan (csrc/shuffle_host.hpp) over this curve: 96-byte points, ark_bls12_377::Fr
// challenges (3 shaved bits), 97-byte points in the transcript.  Same wrapper as tests/host/host_shim.cpp.
#include "../../mental-poker_b200/csrc/shuffle_host.hpp"
extern "C" {
int h_point_bytes() { return (int)kPointBytes; }
void h_fs_challenges(const uint8_t* data, uint64_t len, int count, uint8_t* out) {
  Transcript fs;
  if (len) { fs.begin(); fs.feed(data, len / 2); fs.feed(data + len / 2, len - len / 2); fs.end(); }
  for (int i = 0; i < count; i++) { fr c = fs.challenge(); uint32_t w[8]; fr_to_canonical(c, w); memcpy(out + 32 * i, w, 32); }
}
void h_fr_mul_canonical(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  fr x = fr_from_canonical(a), y = fr_from_canonical(b);
  fr_to_canonical(fr_mul(x, y), out);
}
int h_verify_plan(int m, int n, const uint8_t* enc_g, const uint8_t* ck_g, const uint8_t* ck_h, const uint8_t* ghat,
                  const uint8_t* gsum, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2, const uint8_t* proof,
                  uint8_t* g1_pts, uint8_t* g1_scal, int* job_lens, uint8_t* sx, uint8_t* s2, uint8_t* ss, uint8_t* small_pts,
                  int* host_flags) {
  const size_t PB = kPointBytes;
  ShuffleParamsHost S;
  S.m = m; S.n = n;
  S.ck64.resize((size_t)(n + 1) * PB);
  memcpy(S.ck64.data(), ck_h, PB);
  memcpy(S.ck64.data() + PB, ck_g, (size_t)n * PB);
  memcpy(S.enc_g, enc_g, PB); memcpy(S.ghat, ghat, PB); memcpy(S.gsum, gsum, PB);
  const Layout L(m, n);
  const size_t N = (size_t)m * n;
  const Challenges ch = derive_challenges(&S, pk, deck, deck2, N, proof, L);
  TermList tl;
  HostChecks hc;
  append_g1_checks(tl, &S, proof, L, ch, &hc);
  memcpy(g1_pts, tl.pts.data(), tl.pts.size());
  memcpy(g1_scal, tl.scal.data(), tl.scal.size() * 4);
  for (int j = 0; j < kG1Checks; j++) job_lens[j] = (int)tl.jobs[j].len;
  fr bstar;
  build_ct_plan(&S, pk, proof, L, ch, (uint32_t*)sx, (uint32_t*)s2, (uint32_t*)ss, small_pts, &bstar);
  host_flags[0] = hc.hadamard_bytes_ok; host_flags[1] = hc.zero_bytes_ok; host_flags[2] = hc.svp_first_ok;
  host_flags[3] = fr_eq(hc.svp_last, fr_mul(hc.xs, bstar)); host_flags[4] = hc.multiexp_bytes_ok;
  return (int)tl.count();
}
// the two ciphertext equations as the contiguous jobs mp377_shuffle_verify launches (assemble_ct_jobs)
int h_ct_jobs(int m, int n, const uint8_t* enc_g, const uint8_t* ck_g, const uint8_t* ck_h, const uint8_t* ghat,
              const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2, const uint8_t* proof, uint8_t* cts_out,
              uint8_t* scal_out, uint32_t* jobs_out) {
  const size_t PB = kPointBytes;
  ShuffleParamsHost S;
  S.m = m; S.n = n;
  S.ck64.resize((size_t)(n + 1) * PB);
  memcpy(S.ck64.data(), ck_h, PB);
  memcpy(S.ck64.data() + PB, ck_g, (size_t)n * PB);
  memcpy(S.enc_g, enc_g, PB); memcpy(S.ghat, ghat, PB); memset(S.gsum, 0, PB);
  const Layout L(m, n);
  const Challenges ch = derive_challenges(&S, pk, deck, deck2, (size_t)m * n, proof, L);
  std::vector<uint8_t> cts;
  std::vector<uint32_t> scal;
  fr bstar;
  assemble_ct_jobs(&S, pk, deck, deck2, proof, L, ch, cts, scal, jobs_out, &bstar);
  memcpy(cts_out, cts.data(), cts.size());
  memcpy(scal_out, scal.data(), scal.size() * 4);
  return (int)(scal.size() / 8);
}
int h_verdict(const int* g1_id, int ct_ok, const int* host_flags) {
  HostChecks hc;
  hc.hadamard_bytes_ok = host_flags[0]; hc.zero_bytes_ok = host_flags[1]; hc.svp_first_ok = host_flags[2];
  hc.multiexp_bytes_ok = host_flags[4];
  hc.xs = fr_one(); hc.svp_last = host_flags[3] ? fr_mul(fr_one(), fr_one()) : fr_zero();
  bool ids[kG1Checks];
  for (int j = 0; j < kG1Checks; j++) ids[j] = g1_id[j] != 0;
  return verdict(hc, fr_one(), ids, ct_ok != 0);
}
}

// ---- host half of the batched sigma protocols (csrc/sigma_host.hpp) over this curve
#include "../../mental-poker_b200/csrc/sigma_host.hpp"
extern "C" {
void h_cp_challenge(int which, const uint8_t* g, const uint8_t* h, const uint8_t* s0, const uint8_t* s1, const uint8_t* a,
                    const uint8_t* b, uint8_t* out) {
  const char* seed = which == 0 ? kSeedMasking : which == 1 ? kSeedRemasking : kSeedReveal;
  fr_to_bytes(cp_challenge(Transcript(seed, strlen(seed)), g, h, s0, s1, a, b), out);
}
void h_schnorr_challenge(const uint8_t* info, uint64_t info_len, const uint8_t* g, const uint8_t* pk, const uint8_t* commit,
                         uint8_t* out) {
  fr_to_bytes(schnorr_challenge(info, info_len, g, pk, commit), out);
}
int h_fr_bytes_canonical(const uint8_t* b) { return fr_bytes_canonical(b); }
int h_sigma_proof_lens(int which) { return which == 0 ? (int)kCpProofLen : (int)kSchnorrProofLen; }
}

// ---- wire format over this curve: square roots in F_q (csrc/fq_sqrt.cuh: two-adicity 46, 2-bit windows)
#include "../../mental-poker_b200/csrc/fq_sqrt.cuh"
extern "C" {
// canonical a -> canonical root; returns 1 if a is a square, 0 otherwise
int h_fq_sqrt(const uint32_t* a, uint32_t* out) {
  static fq T[kTwoAdicity];
  static bool ready = false;
  if (!ready) { fq_sqrt_table(T); ready = true; }
  fq x; memcpy(x.v, a, 48);
  bool ok;
  fq r = fq_sqrt(fq_reduce_full(fq_to_mont(x)), T, &ok);
  fq c = fq_from_mont(r);
  memcpy(out, c.v, 48);
  return ok ? 1 : 0;
}
// the windowed form the GPU runs; *distinct_keys = number of distinct lut keys (must be kSqrtRadix = 4)
int h_fq_sqrt_win(const uint32_t* a, uint32_t* out, int* distinct_keys) {
  static fq T[kTwoAdicity], Tinv[kTwoAdicity];
  static std::vector<fq> U(kSqrtUCount), V(kSqrtVCount);
  static std::vector<uint8_t> lut(65536, 0xff), hit(65536, 0);
  static bool ready = false;
  static int keys = 0;
  if (!ready) {
    fq_sqrt_table(T);
    fq_sqrt_inverse_table(T, Tinv);
    for (size_t g = 0; g < kSqrtUCount + kSqrtVCount + kSqrtRadix; g++) fq_sqrt_fill_entry(T, Tinv, g, U.data(), V.data(), lut.data());
    for (uint32_t j = 0; j < (uint32_t)kSqrtRadix; j++) {
      fq acc = fq_one();
      for (int k = 0; k < kSqrtWin; k++)
        if ((j >> k) & 1u) acc = fq_mul(acc, T[kTwoAdicity - kSqrtWin + k]);
      uint32_t key = fq_sqrt_key(fq_reduce_full(acc));
      if (!hit[key]) { hit[key] = 1; keys++; }
    }
    ready = true;
  }
  *distinct_keys = keys;
  fq x; memcpy(x.v, a, 48);
  bool ok;
  SqrtTables tb{U.data(), V.data(), lut.data()};
  fq r = fq_sqrt_win(fq_reduce_full(fq_to_mont(x)), tb, &ok);
  fq c = fq_from_mont(r);
  memcpy(out, c.v, 48);
  return ok ? 1 : 0;
}
int h_fq_half_is_half(const uint32_t* q_minus_1_over_2) {
  for (int i = 0; i < kFqLimbs; i++) if (fq_half_limb(i) != q_minus_1_over_2[i]) return 0;
  return 1;
}
}
