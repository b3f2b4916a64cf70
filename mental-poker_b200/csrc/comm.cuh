// Communicator plumbing shared by the translation units that have a multi-GPU path (comm.cu, diag.cu,
// shuffle_verify.cu).  See comm.cu for the design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct mp_ctx;

namespace mp {
int comm_size(const mp_ctx* ctx);   // 1 without a communicator
int comm_rank(const mp_ctx* ctx);
// balanced contiguous shard [begin, end) of `total` items for `rank` of `nranks`
void comm_shard(uint64_t total, int rank, int nranks, uint64_t* begin, uint64_t* end);
// in-place all-gather on `st`: rank r's bytes_per_rank bytes live at d_buf + r * bytes_per_rank
int32_t comm_allgather(mp_ctx* ctx, void* d_buf, size_t bytes_per_rank, cudaStream_t st);
void comm_destroy(mp_ctx* ctx);
// true while one of the *_multi protocol entry points is running on this context: the call is collective (every
// rank runs it on the same inputs) and the large independent pieces of work are split by rank
bool comm_collective(const mp_ctx* ctx);
}  // namespace mp
