// Device kernels over vectors of scalar-field elements (kernel family K6 of SURVEY.md 2b):
// the O(N) and O(m^2 n) scalar work inside `ShuffleArgument::{prove,verify}` (reference call
// sites src/discrete_log_cards/mod.rs:409-415,437-442; algebra in SURVEY.md Appendix B).
// Vectors are arrays of `fr` (Montgomery form, fully reduced) unless a parameter says
// "canonical" (plain little-endian integers < n, the form the MSM digit extractor reads).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fr.cuh"

namespace mp {

// pow2[k] = x^(2^k), k = 0..31 (host-computed).
struct FrPow2Table {
  fr p[32];
};

// out_canon[i] = x^(i+1)  (i < N, canonical; may be null), out_mont likewise in Montgomery form
// (may be null).  If bstar != null:  *bstar = prod_{i=1..N} (y*i + x^i - z)  (Montgomery);
// `partials` must hold >= fr_powers_blocks(N) elements.
unsigned fr_powers_blocks(uint64_t N);
cudaError_t fr_powers(const FrPow2Table& tab, uint64_t N, uint32_t* out_canon, fr* out_mont,
                      const fr* yz /* device: {y, z} */, fr* partials, fr* bstar, cudaStream_t stream);

// out_canon[i*n + j] = coef[i] * a[j]   (i < m, j < n)
cudaError_t fr_outer_canonical(const fr* coef, const fr* a, int m, int n, uint32_t* out_canon,
                               cudaStream_t stream);

// canonical <-> Montgomery conversion of `count` elements
cudaError_t fr_to_canonical_vec(const fr* in, uint32_t* out_canon, uint64_t count, cudaStream_t stream);
cudaError_t fr_from_canonical_vec(const uint32_t* in_canon, fr* out, uint64_t count, cudaStream_t stream);

// ---- prover-side vector kernels (Appendix B.1-B.5') ---------------------------------------
// a[i] = perm[i] + 1;  b[i] = xpow[perm[i]]  (xpow[k] = x^(k+1));  both Montgomery
cudaError_t fr_perm_vectors(const uint32_t* perm, const fr* xpow, uint64_t N, fr* a, fr* b, cudaStream_t stream);
// d[i] = y*a[i] + b[i] - z
cudaError_t fr_affine_comb(const fr* a, const fr* b, const fr* yz /* device: {y, z} */, uint64_t N, fr* d, cudaStream_t stream);
// column prefix products over the m rows of D (m x n, row-major): Bv[k][j] = prod_{k' <= k} D[k'][j]
cudaError_t fr_column_prefix_products(const fr* D, int m, int n, fr* Bv, cudaStream_t stream);
// out[k][j] = coef[k] * in[k][j]   (rows rows)
cudaError_t fr_scale_rows(const fr* in, const fr* coef, int rows, int n, fr* out, cudaStream_t stream);
// out[j] = sum_k coef[k] * rows[k][j]  where row k starts at rows + k*stride
cudaError_t fr_lincomb_rows(const fr* rows, uint64_t stride, const fr* coef, int count, int n, fr* out, cudaStream_t stream);
// *out = sum_i a[i] * b[i]   (N elements; `partials` >= fr_reduce_blocks(N))
unsigned fr_reduce_blocks(uint64_t N);
cudaError_t fr_dot(const fr* a, const fr* b, uint64_t N, fr* partials, fr* out, cudaStream_t stream);
// Zero-argument diagonals (B.4):  A, B are (rows x n) row-major, ypow[j] = y^(j+1).
// d[k] = sum over (i, j) with k == i + rows - 1 - j ... of sum_t A[i][t] * B[j][t] * ypow[t];
// precisely d[k] = sum_{i=0..rows-1} sum_{j=0..rows-1, i + (rows-1) - j == k} <A_i, B_j>_y, k = 0..2*rows-2.
cudaError_t fr_bilinear_diagonals(const fr* A, const fr* B, const fr* ypow, int rows, int n, fr* pair_scratch,
                                  fr* d, cudaStream_t stream);
// scatter `count` Montgomery values into a canonical array: out_canon[dst_off + i*dst_stride] = in[i]
cudaError_t fr_scatter_canonical(const fr* in, uint64_t count, uint32_t* out_canon, uint64_t dst_off,
                                 uint64_t dst_stride, cudaStream_t stream);

}  // namespace mp
