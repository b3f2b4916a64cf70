// Engine context plumbing shared by both curve builds: the mp_ctx methods and the in-library create / destroy used
// by the protocol drivers for their worker contexts.  Compiled twice (Makefile): for the Stark curve as is, and
// with -DMP_CURVE_BLS12_377 -Dmp=mp_bls12_377 -Dmp_ctx=mp377_pctx for the second curve, so that the protocol
// translation units (shuffle_setup.cu, shuffle_prove_batch.cu) link into one library once per curve.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/mpshuffle.h"
#include "comm.cuh"
#include "ctx.cuh"
#include "msm.cuh"
#include "shuffle.cuh"

int32_t mp_ctx::fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err = buf;
  return code;
}
int32_t mp_ctx::cuda_fail(cudaError_t e, const char* where) {
  return fail(MP_ERR_CUDA, "CUDA error in %s: %s", where, cudaGetErrorString(e));
}
void* mp_ctx::scratch(int slot, size_t bytes) {
  if ((size_t)slot >= bufs.size()) bufs.resize(slot + 1);
  auto& b = bufs[slot];
  if (b.cap < bytes) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&b.ptr, want) != cudaSuccess) return nullptr;
    b.cap = want;
  }
  return b.ptr;
}

namespace mp {
int32_t ctx_create(mp_ctx** out, int32_t device) {
  if (!out) return MP_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
    fprintf(stderr, "mpshuffle: no usable CUDA device %d (%s); there is no CPU fallback\n", device,
            e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    return MP_ERR_CUDA;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return MP_ERR_CUDA;
  mp_ctx* ctx = new mp_ctx();
  ctx->device = device;
  // Stream priorities were measured and rejected (round 1): with the main stream above the bulk stream
  // (ShuffleState::bulk) the block scheduler holds back the bulk kernel's pending blocks whenever a
  // main-stream kernel is waiting for resources, the SMs drain, and the bulk kernel's launch time grows
  // by exactly what the small kernels took (12.8 vs 10.1 ms) -- same end-to-end time, muddier kernels.
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete ctx; return MP_ERR_CUDA; }
  ctx->ws = msm_workspace_create();
  *out = ctx;
  return MP_OK;
}
void ctx_destroy(mp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
#ifndef MP_CURVE_BLS12_377
  comm_destroy(ctx);
#endif
  msm_workspace_destroy(ctx->ws);
  if (ctx->shuffle) shuffle_state_destroy(ctx->shuffle);
  for (auto& b : ctx->bufs)
    if (b.ptr) cudaFree(b.ptr);
  if (ctx->wait_ev) cudaEventDestroy(ctx->wait_ev);
  if (ctx->mark_ev) cudaEventDestroy(ctx->mark_ev);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
}  // namespace mp
