// TEST INFRASTRUCTURE ONLY: compiles the __host__ side of csrc/fq.cuh + ec.cuh with g++ so the
// word-level algorithms (sparse Montgomery reduction, lazy bounds, XYZZ formulas) can be
// checked against the Python oracle on a machine without a GPU.  Not linked into the product.
#include "../../mental-poker_b200/csrc/ec.cuh"
#include <string.h>
using namespace mp;
extern "C" {
void h_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  fq x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
  fq r = fq_mul(x, y); memcpy(out, r.v, 32);
}
void h_fq_reduce_weak(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 32); fq r = fq_reduce_weak(x); memcpy(out, r.v, 32);
}
void h_fq_reduce_full(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 32); fq r = fq_reduce_full(x); memcpy(out, r.v, 32);
}
void h_fq_inv_canonical(const uint32_t* a, uint32_t* out) {
  fq x; memcpy(x.v, a, 32);
  fq r = fq_from_mont(fq_inv(fq_to_mont(x))); memcpy(out, r.v, 32);
}
int h_on_curve(const uint32_t* p) { return affine_on_curve(affine_from_canonical(p)); }
// out = P + Q via madd (P lifted to XYZZ)
void h_point_add(const uint32_t* p, const uint32_t* q, uint32_t* out) {
  xyzz acc = xyzz_from_affine(affine_from_canonical(p));
  xyzz_madd(acc, affine_from_canonical(q));
  affine_to_canonical(xyzz_to_affine(acc), out);
}
// out = k*P by MSB-first double-and-add using xyzz_dbl + xyzz_madd;  k = 8 words LE
void h_scalar_mul(const uint32_t* p, const uint32_t* k, uint32_t* out) {
  affine P = affine_from_canonical(p);
  xyzz acc = xyzz_identity();
  for (int i = 255; i >= 0; i--) {
    acc = xyzz_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_madd(acc, P);
  }
  affine_to_canonical(xyzz_to_affine(acc), out);
}
// out = (k1*P) + (k2*Q) with the final addition done by xyzz_add (general add)
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
}

// ---- scalar field + transcript (host side of csrc/fr.cuh, csrc/transcript.hpp)
#include "../../mental-poker_b200/csrc/transcript.hpp"
extern "C" {
void h_fr_mul_canonical(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  fr x = fr_from_canonical(a), y = fr_from_canonical(b);
  fr_to_canonical(fr_mul(x, y), out);
}
void h_fr_addsub_canonical(const uint32_t* a, const uint32_t* b, uint32_t* sum, uint32_t* diff, uint32_t* neg) {
  fr x = fr_from_canonical(a), y = fr_from_canonical(b);
  fr_to_canonical(fr_add(x, y), sum);
  fr_to_canonical(fr_sub(x, y), diff);
  fr_to_canonical(fr_neg(x), neg);
}
void h_blake2s(const uint8_t* data, uint64_t len, uint64_t split, uint8_t* out) {
  Blake2s h;
  if (split > len) split = len;
  h.update(data, split);          // exercise the buffering paths
  h.update(data + split, len - split);
  h.finish(out);
}
// challenges after absorbing `data` (in two feeds) into a fresh "Shuffle Proof" transcript
void h_fs_challenges(const uint8_t* data, uint64_t len, int count, uint8_t* out) {
  Transcript fs;
  if (len) { fs.begin(); fs.feed(data, len / 2); fs.feed(data + len / 2, len - len / 2); fs.end(); }
  for (int i = 0; i < count; i++) { fr c = fs.challenge(); uint32_t w[8]; fr_to_canonical(c, w); memcpy(out + 32 * i, w, 32); }
}
// absorb `count` 64-byte points through feed_points64 and return one challenge
void h_fs_points_challenge(const uint8_t* pts, uint64_t count, uint8_t* out) {
  Transcript fs;
  fs.begin(); fs.feed_label("label"); fs.feed_points64(pts, count); fs.end();
  fr c = fs.challenge(); uint32_t w[8]; fr_to_canonical(c, w); memcpy(out, w, 32);
}
}

// ---- multi-stream Blake2s (Blake2sLanes / TranscriptLanes of csrc/transcript.hpp)
extern "C" {
// `lanes` streams of `len` bytes (stream l at data + l * stride), fed in pieces of `piece` bytes; out = lanes * 32 digest bytes.
// mode 0: vectorised if the CPU allows, 1: also finish through extract() after only `cut` bytes went through the lanes
void h_blake2s_lanes(const uint8_t* data, uint64_t stride, uint64_t len, int lanes, uint64_t piece, uint64_t cut, uint8_t* out) {
  Blake2sLanes mb(lanes);
  if (cut > len) cut = len;
  const uint8_t* p[Blake2sLanes::kLanes];
  for (uint64_t off = 0; off < cut;) {
    const uint64_t take = piece < cut - off ? piece : cut - off;
    for (int l = 0; l < lanes; l++) p[l] = data + l * stride + off;
    mb.update(p, take);
    off += take;
  }
  for (int l = 0; l < lanes; l++) {
    Blake2s h;
    mb.extract(l, &h);
    h.update(data + l * stride + cut, len - cut);   // the rest goes through the single-stream hasher
    h.finish(out + 32 * l);
  }
}
int h_blake2s_lanes_vectorised() { return Blake2sLanes::vectorised() ? 1 : 0; }
// lanes transcripts: label | shared points | per-lane points (lockstep), then per lane: more points + end -> one challenge each
void h_fs_lanes_challenges(const uint8_t* shared_pts, uint64_t n_shared, const uint8_t* lane_pts, uint64_t n_lane, int lanes,
                           const uint8_t* tail_pts, uint64_t n_tail, uint8_t* out) {
  TranscriptLanes tl(lanes);
  tl.feed_label_all("label");
  tl.feed_points64_all(shared_pts, n_shared);
  const uint8_t* p[Blake2sLanes::kLanes];
  for (int l = 0; l < lanes; l++) p[l] = lane_pts + 64 * n_lane * l;
  tl.feed_points64(p, n_lane);
  for (int l = 0; l < lanes; l++) {
    Transcript fs;
    tl.hand_over(l, &fs);
    fs.feed_points64(tail_pts, n_tail);
    fs.end();
    fr c = fs.challenge(); uint32_t w[8]; fr_to_canonical(c, w); memcpy(out + 32 * l, w, 32);
  }
}
// the same through one ordinary transcript per lane
void h_fs_single_challenge(const uint8_t* shared_pts, uint64_t n_shared, const uint8_t* lane_pts, uint64_t n_lane,
                           const uint8_t* tail_pts, uint64_t n_tail, uint8_t* out) {
  Transcript fs;
  fs.begin(); fs.feed_label("label"); fs.feed_points64(shared_pts, n_shared); fs.feed_points64(lane_pts, n_lane);
  fs.feed_points64(tail_pts, n_tail); fs.end();
  fr c = fs.challenge(); uint32_t w[8]; fr_to_canonical(c, w); memcpy(out, w, 32);
}
}

// ---- host-only verifier plan (csrc/shuffle_host.hpp): challenges, the eight commitment-space
// jobs and the two ciphertext equations of one proof, returned as flat (point, scalar) lists
#include "../../mental-poker_b200/csrc/shuffle_host.hpp"
extern "C" {
// g1_pts: T1*64, g1_scal: T1*32, job_lens: 8 ints; ct arrays as documented at build_ct_plan.
// host_flags[5]: hadamard bytes, zero bytes, svp first, svp last (vs bstar), multi-exp bytes.
int h_verify_plan(int m, int n, const uint8_t* enc_g, const uint8_t* ck_g, const uint8_t* ck_h, const uint8_t* ghat,
                  const uint8_t* gsum, const uint8_t* pk, const uint8_t* deck, const uint8_t* deck2, const uint8_t* proof,
                  uint8_t* g1_pts, uint8_t* g1_scal, int* job_lens, uint8_t* sx, uint8_t* s2, uint8_t* ss, uint8_t* small_pts,
                  int* host_flags) {
  ShuffleParamsHost S;
  S.m = m; S.n = n;
  S.ck64.resize((size_t)(n + 1) * 64);
  memcpy(S.ck64.data(), ck_h, 64);
  memcpy(S.ck64.data() + 64, ck_g, (size_t)n * 64);
  memcpy(S.enc_g, enc_g, 64); memcpy(S.ghat, ghat, 64); memcpy(S.gsum, gsum, 64);
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
int h_verdict(const int* g1_id, int ct_ok, const int* host_flags) {
  HostChecks hc;
  hc.hadamard_bytes_ok = host_flags[0]; hc.zero_bytes_ok = host_flags[1]; hc.svp_first_ok = host_flags[2];
  hc.multiexp_bytes_ok = host_flags[4];
  // svp_last check is passed pre-evaluated: encode it as (svp_last, xs, bstar) = (1, 1, 1) or (0, 1, 1)
  hc.xs = fr_one(); hc.svp_last = host_flags[3] ? fr_mul(fr_one(), fr_one()) : fr_zero();
  bool ids[kG1Checks];
  for (int j = 0; j < kG1Checks; j++) ids[j] = g1_id[j] != 0;
  return verdict(hc, fr_one(), ids, ct_ok != 0);
}
}

// ---- host-only plan of the prover's Karatsuba diagonal products (csrc/diag_plan.hpp)
#include "../../mental-poker_b200/csrc/diag_plan.hpp"
extern "C" {
// sizes[0] = nleaf, sizes[1] = number of CSR entries; arrays may be null to query the sizes
void h_diag_plan(int m, uint32_t* sizes, uint32_t* leaf_mask, uint32_t* leaf_val, uint32_t* single, uint32_t* row_start,
                 uint32_t* entries) {
  const DiagPlan p = diag_plan_build(m);
  sizes[0] = p.nleaf();
  sizes[1] = (uint32_t)p.entries.size();
  if (!leaf_mask) return;
  memcpy(leaf_mask, p.leaf_mask.data(), 4 * p.leaf_mask.size());
  memcpy(leaf_val, p.leaf_val.data(), 4 * p.leaf_val.size());
  memcpy(single, p.single.data(), 4 * p.single.size());
  memcpy(row_start, p.row_start.data(), 4 * p.row_start.size());
  memcpy(entries, p.entries.data(), 4 * p.entries.size());
}
}

// ---- host-only half of the batched sigma protocols (csrc/sigma_host.hpp)
#include "../../mental-poker_b200/csrc/sigma_host.hpp"
extern "C" {
// which: 0 masking, 1 remasking, 2 reveal seeds
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
}

// ---- wire format: square roots (csrc/fq_sqrt.cuh) and compression (csrc/wire_host.hpp)
#include "../../mental-poker_b200/csrc/fq_sqrt.cuh"
#include "../../mental-poker_b200/csrc/wire_host.hpp"
extern "C" {
// canonical a -> canonical root; returns 1 if a is a square, 0 otherwise
int h_fq_sqrt(const uint32_t* a, uint32_t* out) {
  static fq T[kTwoAdicity];
  static bool ready = false;
  if (!ready) { fq_sqrt_table(T); ready = true; }
  fq x; memcpy(x.v, a, 32);
  bool ok;
  fq r = fq_sqrt(fq_reduce_full(fq_to_mont(x)), T, &ok);
  fq c = fq_from_mont(r);
  memcpy(out, c.v, 32);
  return ok ? 1 : 0;
}
// the windowed form the GPU runs; returns 1 if a is a square.  *distinct_keys = number of distinct lut keys (must be 256)
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
    // the 256 roots of unity must land on 256 distinct lut keys
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
  fq x; memcpy(x.v, a, 32);
  bool ok;
  SqrtTables tb{U.data(), V.data(), lut.data()};
  fq r = fq_sqrt_win(fq_reduce_full(fq_to_mont(x)), tb, &ok);
  fq c = fq_from_mont(r);
  memcpy(out, c.v, 32);
  return ok ? 1 : 0;
}
void h_wire_compress(const uint8_t* points, uint64_t n, uint8_t* out) {
  for (uint64_t i = 0; i < n; i++) wire_compress_point(points + 64 * i, out + 32 * i);
}
uint64_t h_wire_proof_serialize(int m, int n, const uint8_t* proof, uint8_t* out) {
  wire_proof_serialize(m, n, proof, out);
  return wire_proof_len(m, n);
}
}
