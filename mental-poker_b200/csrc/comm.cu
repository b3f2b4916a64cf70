// Multi-GPU entry points of the C ABI: one context per GPU -- one process per GPU, or one
// thread per GPU inside one process -- joined by an NCCL communicator, so that a non-Python host behind
// BarnettSmartProtocol::{shuffle_and_remask, verify_shuffle} (reference src/lib.rs:181-197; e.g. the four chained
// shuffles of examples/round.rs:265-350) has the same multi-GPU paths bench.py measures:
//
//   * one large MSM, window-range split (BASELINE config 5): every rank holds the inputs, runs the windows
//     shard(W, rank, nranks) end to end, the 128-byte XYZZ partials are all-gathered ON THE CONTEXT'S STREAM (no
//     host synchronisation in between) and one quad-cooperative kernel folds them:  sum_r 2^(c * w_begin_r) * P_r.
//     EC addition is not an ncclRedOp, so the "reduce of partial bucket sums" is bytes + a local fold;
//   * batches of independent proofs, proof-index split (config 4): no data-path collective; the verdicts are
//     all-gathered so that every rank holds the whole status vector;
//   * one large proof across GPUs (config 3): the prover's Karatsuba leaf products -- 2 187 independent MSMs at
//     m = 128, 70 % of the prover's device time -- are split by leaf index and their 256-byte results all-gathered
//     (diag.cu); the verifier's two ciphertext equations go to two ranks (shuffle_verify.cu).  All ranks run the
//     same call on the same inputs and return the same bytes.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy already in the process if there is one -- e.g.
// PyTorch's) so that hosts without NCCL still load the library; without it these entry points return MP_ERR_NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "../../include/mpshuffle.h"
#include "comm.cuh"
#include "ctx.cuh"
#include "msm.cuh"
#include "shuffle.cuh"

using namespace mp;

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);  // the copy already mapped into the process, if any
      if (!a.handle) a.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.handle) break;
    }
    if (!a.handle) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
    return a;
  }();
  return &api;
}
int32_t nccl_fail(mp_ctx* ctx, ncclResult_t r, const char* where) {
  return ctx->fail(MP_ERR_NCCL, "NCCL error in %s: %s", where, nccl_api()->GetErrorString(r));
}
}  // namespace

struct mp_comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};

namespace mp {
int comm_size(const mp_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->nranks : 1; }
int comm_rank(const mp_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->rank : 0; }
void comm_shard(uint64_t total, int rank, int nranks, uint64_t* begin, uint64_t* end) {
  const uint64_t base = total / nranks, rem = total % nranks;
  *begin = (uint64_t)rank * base + std::min<uint64_t>(rank, rem);
  *end = *begin + base + ((uint64_t)rank < rem ? 1 : 0);
}
// in place: rank r's `bytes_per_rank` bytes are expected at d_buf + r * bytes_per_rank
int32_t comm_allgather(mp_ctx* ctx, void* d_buf, size_t bytes_per_rank, cudaStream_t st) {
  if (!ctx->comm) return MP_OK;
  const ncclResult_t r = nccl_api()->AllGather((const uint8_t*)d_buf + (size_t)ctx->comm->rank * bytes_per_rank, d_buf, bytes_per_rank,
                                               ncclUint8, ctx->comm->comm, st);
  if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
  return MP_OK;
}
bool comm_collective(const mp_ctx* ctx) { return ctx && ctx->comm && ctx->collective && ctx->comm->nranks > 1; }
void comm_destroy(mp_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  if (ctx->comm->comm) nccl_api()->CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
}
}  // namespace mp

extern "C" int32_t mp_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return MP_ERR_INVALID_ARG;
  NcclApi* api = nccl_api();
  if (!api->ok) return MP_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == MP_COMM_ID_BYTES, "NCCL unique id size");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return MP_ERR_NCCL;
  memcpy(id_out, &id, sizeof id);
  return MP_OK;
}

extern "C" int32_t mp_comm_init(mp_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t* id) {
  if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return ctx ? ctx->fail(MP_ERR_INVALID_ARG, "bad communicator shape") : MP_ERR_INVALID_ARG;
  NcclApi* api = nccl_api();
  if (!api->ok) return ctx->fail(MP_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
  comm_destroy(ctx);
  cudaSetDevice(ctx->device);
  mp_comm* c = new mp_comm();
  c->nranks = nranks;
  c->rank = rank;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  const ncclResult_t r = api->CommInitRank(&c->comm, nranks, uid, rank);
  if (r != ncclSuccess) { delete c; return nccl_fail(ctx, r, "ncclCommInitRank"); }
  ctx->comm = c;
  return MP_OK;
}
extern "C" int32_t mp_comm_destroy(mp_ctx* ctx) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_destroy(ctx);
  return MP_OK;
}
extern "C" int32_t mp_comm_size(mp_ctx* ctx) { return comm_size(ctx); }
extern "C" int32_t mp_comm_rank(mp_ctx* ctx) { return comm_rank(ctx); }

// ---- one MSM, window-range split --------------------------------------------------------------------------------
extern "C" int32_t mp_msm_g1_multi_device(mp_ctx* ctx, const void* d_bases, const void* d_scalars, uint64_t n,
                                          int32_t window_bits, void* d_out) {
  if (!ctx || (!d_bases && n) || (!d_scalars && n) || !d_out) return MP_ERR_INVALID_ARG;
  if (n >= (1ull << 31)) return ctx->fail(MP_ERR_INVALID_ARG, "MSM size %llu too large", (unsigned long long)n);
  cudaSetDevice(ctx->device);
  ctx->launches = 0;
  const int G = comm_size(ctx), rank = comm_rank(ctx);
  const int c = window_bits > 0 ? window_bits : msm_pick_window(n);
  if (c < 2 || c > 16) return ctx->fail(MP_ERR_INVALID_ARG, "window_bits %d out of range [2,16]", c);
  if (G > 64) return ctx->fail(MP_ERR_INVALID_ARG, "at most 64 ranks");
  const int W = msm_num_windows(c);
  ctx->last_window = c;
  uint64_t wb, we;
  comm_shard((uint64_t)W, rank, G, &wb, &we);
  affine* mont = (affine*)ctx->scratch(mp_ctx::kSlotPointsMont, sizeof(affine) * n);
  xyzz* parts = (xyzz*)ctx->scratch(mp_ctx::kSlotMsmOut, sizeof(xyzz) * (size_t)(G + 1));  // [G] partials, then the result
  int* bad = (int*)ctx->scratch(mp_ctx::kSlotFlags, 256);
  if (!mont || !parts || !bad) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaError_t e;
  if ((e = cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "memset");
  if ((e = cudaMemsetAsync(parts, 0, sizeof(xyzz) * (size_t)(G + 1), ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "memset");  // identity
  if ((e = points_to_mont((const uint32_t*)d_bases, mont, n, bad, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "points_to_mont");
  ctx->launches += n ? 1 : 0;
  ctx->last_ec_adds = 0;
  if (we > wb) {
    const MsmJob job{0, 0, (uint32_t)n};
    if ((e = msm_run(ctx->ws, (const uint32_t*)d_scalars, n, mont, 1, &job, 1, c, parts + rank, ctx->stream, (int)wb, (int)(we - wb))) != cudaSuccess)
      return ctx->cuda_fail(e, "msm_run");
    ctx->launches += msm_last_launches(ctx->ws);
    const uint64_t B = 1ull << (c - 1);
    ctx->last_ec_adds = (we - wb) * (n + 2 * B) + (we - wb) * (uint64_t)c;
  }
  int32_t st = comm_allgather(ctx, parts, sizeof(xyzz), ctx->stream);  // on the stream: no host round trip
  if (st != MP_OK) return st;
  int shifts[64];  // doublings between rank r + 1's partial and rank r's: c * (windows of rank r)
  for (int r = 0; r < G; r++) {
    uint64_t b, en;
    comm_shard((uint64_t)W, r, G, &b, &en);
    shifts[r] = c * (int)(en - b);
  }
  if ((e = msm_fold_ranges(parts, G, 1, shifts, parts + G, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "msm_fold_ranges");
  if ((e = xyzz_to_canonical(parts + G, (uint32_t*)d_out, 1, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "xyzz_to_canonical");
  ctx->launches += 2;
  return MP_OK;
}

// ---- batches of independent proofs: proof-index split, verdicts all-gathered ---------------------------------------
extern "C" int32_t mp_shuffle_verify_batch_multi(mp_ctx* ctx, const uint8_t* pk, const uint8_t* decks, const uint8_t* shuffled_decks,
                                                 const uint8_t* proofs, uint64_t batch_per_rank, int32_t* statuses_all,
                                                 int32_t host_threads) {
  if (!ctx || !statuses_all) return MP_ERR_INVALID_ARG;
  const int G = comm_size(ctx), rank = comm_rank(ctx);
  int32_t st = shuffle_verify_batch(ctx, pk, decks, shuffled_decks, proofs, batch_per_rank, statuses_all + (size_t)rank * batch_per_rank,
                                    host_threads);
  if (st != MP_OK || G == 1 || batch_per_rank == 0) return st;
  cudaSetDevice(ctx->device);
  int32_t* d = (int32_t*)ctx->scratch(mp_ctx::kSlotStageOut, sizeof(int32_t) * batch_per_rank * G + 64);
  if (!d) return ctx->fail(MP_ERR_CUDA, "device allocation failed");
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d + (size_t)rank * batch_per_rank, statuses_all + (size_t)rank * batch_per_rank, sizeof(int32_t) * batch_per_rank,
                           cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "H2D statuses");
  if ((st = comm_allgather(ctx, d, sizeof(int32_t) * batch_per_rank, ctx->stream)) != MP_OK) return st;
  if ((e = cudaMemcpyAsync(statuses_all, d, sizeof(int32_t) * batch_per_rank * G, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
    return ctx->cuda_fail(e, "D2H statuses");
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return ctx->cuda_fail(e, "status all-gather");
  return MP_OK;
}

// ---- one large proof across GPUs: the same call on every rank, same inputs, same bytes out ----
namespace {
struct CollectiveScope {
  mp_ctx* ctx;
  explicit CollectiveScope(mp_ctx* c) : ctx(c) { ctx->collective = true; }
  ~CollectiveScope() { ctx->collective = false; }
};
}  // namespace
extern "C" int32_t mp_shuffle_and_remask_multi(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint32_t* perm,
                                               const uint8_t* rho, const uint8_t* randomness, uint8_t* out_deck, uint8_t* proof_out) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  CollectiveScope scope(ctx);
  return mp_shuffle_and_remask(ctx, pk, deck, perm, rho, randomness, out_deck, proof_out);
}
extern "C" int32_t mp_shuffle_verify_multi(mp_ctx* ctx, const uint8_t* pk, const uint8_t* deck, const uint8_t* shuffled_deck,
                                           const uint8_t* proof) {
  if (!ctx) return MP_ERR_INVALID_ARG;
  CollectiveScope scope(ctx);
  return mp_shuffle_verify(ctx, pk, deck, shuffled_deck, proof);
}
