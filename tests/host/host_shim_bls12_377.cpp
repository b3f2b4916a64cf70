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
