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
