// Karatsuba evaluation of the prover's diagonal ciphertext products (see diag.cu / diag_plan.hpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "diag_plan.hpp"
#include "ec.cuh"

struct mp_ctx;

namespace mp {

struct DiagDevice;  // plan + its device copy, cached per (m) in the ShuffleState
void diag_device_destroy(DiagDevice*);

// Schoolbook diagonals over the pre-shifted deck table (window c_table) or Karatsuba leaves?
// MP_DIAG_KARATSUBA=0/1 overrides the cost model (the tests drive both at the same sizes).
bool diag_karatsuba_selected(int m, int n, int c_table);

// Leaf point rows from the shuffled deck (Montgomery affine, 2 components interleaved, N = m*n
// ciphertexts).  Independent of every challenge.
int32_t diag_karatsuba_points(mp_ctx* ctx, const affine* d_deck2, cudaStream_t st);
// d_rows_canon: canonical scalar rows a0 | b_1 .. b_m ((m+1)*n scalars).  Writes the 2m diagonal
// products (without their Enc(b_k; tau_k) terms) to d_E[2k + comp].  Runs on `st` with workspace `ws`.
struct MsmWorkspace;
int32_t diag_karatsuba_products(mp_ctx* ctx, const uint32_t* d_rows_canon, xyzz* d_E, cudaStream_t st, MsmWorkspace* ws);

}  // namespace mp
