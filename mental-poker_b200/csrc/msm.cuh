// Batched windowed-Pippenger variable-base MSM over the Stark curve (kernel family K1/K2 of
// SURVEY.md section 2b).  Replaces the scalar-mul / dot-product loops that run inside
// proof-essentials' `ShuffleArgument::{prove,verify}` (reference call sites
// src/discrete_log_cards/mod.rs:409-415,437-442).
//
// One call evaluates `njobs` MSMs that share a scalar array and a point array; job k covers
// scalars [scalar_off, scalar_off+len) against points [point_off, point_off+len) (x ncomp
// interleaved components, ncomp = 2 for ElGamal ciphertexts: both components share digits
// and the sorted index list).  Pipeline, all on one stream, no host synchronisation:
//
//   k_digits         signed c-bit digits of every scalar                 (HBM streaming)
//   k_count          histogram of (job, window, |digit|) buckets         (global atomics)
//   scan             exclusive prefix sum of bucket sizes                (3 small kernels)
//   k_scatter        counting-sort scatter of point indices              (global atomics)
//   k_accumulate     fixed-length chunks of the sorted list, XYZZ mixed adds, segmented by
//                    bucket -> perfectly load balanced for ANY digit distribution
//   k_stitch         fold the partial sums of bucket runs that straddle chunk boundaries
//   k_reduce_seg     per segment of L buckets: S = sum, T = sum (i+1) * B_i (running sums)
//   k_reduce_group(_quad)  per window: groups of 4 segment sums folded level after level (thread or quad per group)
//   k_fold           per job: Horner over windows (c doublings each)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ec.cuh"
#include "msm_job.h"

namespace mp {


struct MsmWorkspace;  // opaque, owns device scratch that grows on demand

MsmWorkspace* msm_workspace_create();
void msm_workspace_destroy(MsmWorkspace*);

// Heuristic window width for a launch of `njobs` jobs of average length `avg_len`.
int msm_pick_window(uint64_t avg_len, uint64_t njobs = 1);

// d_scalars: canonical little-endian 256-bit scalars (< group order), 8 words each.
// d_points : Montgomery-form affine points, index = point * ncomp + comp.
// d_out    : njobs * ncomp XYZZ results (not normalised), index = job * ncomp + comp.
// Returns cudaSuccess or the first CUDA error.  Asynchronous on `stream`.
cudaError_t msm_run(MsmWorkspace* ws, const uint32_t* d_scalars, uint64_t n_scalars,
                    const affine* d_points, int ncomp, const MsmJob* h_jobs, int njobs, int c,
                    xyzz* d_out, cudaStream_t stream, int w_begin = 0, int w_count = -1, uint32_t table_nb = 0);
// With a window range [w_begin, w_begin + w_count) of the W = ceil(253/c) windows the result is
//   sum_{w in range} 2^(c*(w - w_begin)) * (window sum w)
// so that  full MSM = sum over ranges of 2^(c*w_begin) * partial  (window-range split across GPUs).
inline int msm_num_windows(int c) { return (kScalarBits + c - 1) / c; }
// out[comp] = sum_r 2^(shift below r) * parts[r * ncomp + comp]: Horner from the highest rank, shifts[r] doublings
// between rank r + 1's partial and rank r's (window-range split across GPUs; nranks <= 64)
cudaError_t msm_fold_ranges(const xyzz* d_parts, int nranks, int ncomp, const int* shifts, xyzz* d_out, cudaStream_t stream);

// Fixed-base mode (Pedersen commitments over a constant key).  msm_build_table fills
//   d_table[w * nb + first + i] = 2^(c*w) * d_bases[first + i],  w < msm_num_windows(c), i < count
// (affine Montgomery; rebuild a sub-range when one base changes, e.g. the public key).  msm_run with
// table_nb = nb, d_points = d_table and the same c then needs ONE bucket set per job: entries of
// all windows go to bucket |digit|, the reduction runs once per job and no fold doublings remain.
cudaError_t msm_build_table(MsmWorkspace* ws, const affine* d_bases, uint32_t nb, uint32_t first, uint32_t count,
                            int c, affine* d_table, cudaStream_t stream);
int msm_pick_table_window(uint64_t typical_len);

// number of kernels the last msm_run launched / EC additions it performed (host-side count
// of scheduled bucket additions, for EC-adds/s reporting)
int msm_last_launches(const MsmWorkspace* ws);
// Per-launch CUDA-event timing of k_accumulate (the dominant kernel) on the launch stream.
// collect() waits for the recorded events, returns the summed duration, the bucket additions
// those launches were scheduled with and the launch count, and clears the list.
void msm_profile_enable(MsmWorkspace* ws, bool on);
cudaError_t msm_profile_collect(MsmWorkspace* ws, double* ms, uint64_t* adds, uint64_t* launches, double* big_ms = nullptr,
                                uint64_t* big_adds = nullptr, uint64_t* big_launches = nullptr);
// (big_*: the same sums restricted to the dominant launches -- within a factor two of the largest)

// canonical 64-byte points -> Montgomery affine (validating on-curve; bad points set *d_bad)
cudaError_t points_to_mont(const uint32_t* d_canonical, affine* d_out, uint64_t n, int* d_bad,
                           cudaStream_t stream);
// per-item form: d_bad_items[i / per_item] is set when point i is rejected (the batch verifiers, one flag per proof)
cudaError_t points_to_mont_items(const uint32_t* d_canonical, affine* d_out, uint64_t n, int* d_bad_items,
                                 uint64_t per_item, cudaStream_t stream);
// XYZZ -> canonical 64-byte affine (x || y, all-zero = identity); one inversion per point
cudaError_t xyzz_to_canonical(const xyzz* d_in, uint32_t* d_out, uint64_t n, cudaStream_t stream);

}  // namespace mp
