// Engine context behind the opaque `mp_ctx` of include/mpshuffle.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <vector>

namespace mp {
struct MsmWorkspace;
struct ShuffleState;
}

struct mp_comm;  // NCCL communicator of a multi-GPU job (comm.cu); null = single GPU

struct mp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  mp::MsmWorkspace* ws = nullptr;
  mp::ShuffleState* shuffle = nullptr;  // protocol parameters + staging (shuffle_internal.cuh)
  mp_comm* comm = nullptr;
  bool collective = false;  // set for the duration of a *_multi protocol call (comm.cuh)
  std::string err = "";
  int launches = 0;
  uint64_t last_ec_adds = 0;
  int last_window = 0;
  bool wire_table_ready = false;  // sqrt table of wire.cu built in its scratch slot
  bool blocking = false;          // host waits sleep instead of spinning (mp::stream_wait below; set by the batch drivers)
  cudaEvent_t wait_ev = nullptr, mark_ev = nullptr;  // blocking-sync events of mp::stream_wait / mark_record (created on first use)

  struct Buf { void* ptr = nullptr; size_t cap = 0; };
  std::vector<Buf> bufs;
  enum { kSlotStageIn = 0, kSlotStageOut, kSlotPointsMont, kSlotMsmOut, kSlotFlags, kSlotUser };
  // growable device scratch; returns nullptr on allocation failure
  void* scratch(int slot, size_t bytes);

  int32_t fail(int32_t code, const char* fmt, ...);
  int32_t cuda_fail(cudaError_t e, const char* where);
};

namespace mp {
// in-library create / destroy (pctx.cu): what mp_ctx_create / mp_ctx_destroy wrap, and what the protocol drivers use
// for their worker contexts (they must not go through the extern "C" names: those exist once, for the Stark curve)
int32_t ctx_create(mp_ctx** out, int32_t device);
void ctx_destroy(mp_ctx* ctx);

// How a host thread waits for the device.  A batch of large proofs keeps 16 worker threads per GPU (8 provers, 8
// verifiers), each of which alternates a serial Blake2s statement hash (23 ms of real CPU work) with waits for its
// kernels.  cudaStreamSynchronize spins by default, so the waiting threads compete for cores with the hashing ones --
// harmless with a core per thread, costly once several GPUs' workers share one host.  A context whose `blocking` flag
// is set waits on events created with cudaEventBlockingSync (the thread sleeps) instead.  The batch drivers set the
// flag on the worker contexts of LARGE-deck batches; small-deck batches (many short waits per sub-batch) and single
// calls (latency) keep spinning.  Measured with 8 ranks on one 32-vCPU host: 2^16-card proofs 276 -> 292 per second
// blocking, 52-card batches 128 k -> 119 k; one GPU on 16 vCPUs: no difference either way.
// MP_BLOCKING_SYNC=0 / 1 forces spinning / sleeping everywhere.
inline int blocking_policy() {   // -1 = by call (default), 0 = never, 1 = always
  static const int v = [] { const char* e = getenv("MP_BLOCKING_SYNC"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
  return v;
}
inline bool blocking_waits(const mp_ctx* ctx) { return blocking_policy() < 0 ? ctx->blocking : blocking_policy() == 1; }
inline cudaError_t blocking_event(cudaEvent_t* ev) {
  return *ev ? cudaSuccess : cudaEventCreateWithFlags(ev, cudaEventDisableTiming | cudaEventBlockingSync);
}
// wait until everything queued on `st` so far has finished
inline cudaError_t stream_wait(mp_ctx* ctx, cudaStream_t st) {
  if (!blocking_waits(ctx)) return cudaStreamSynchronize(st);
  cudaError_t e;
  if ((e = blocking_event(&ctx->wait_ev)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(ctx->wait_ev, st)) != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->wait_ev);
}
// mark a point of `st` the host will wait for later (work queued after the mark is not waited for); `spin_ev` is the
// caller's ordinary event
inline cudaError_t mark_record(mp_ctx* ctx, cudaEvent_t spin_ev, cudaStream_t st) {
  if (!blocking_waits(ctx)) return cudaEventRecord(spin_ev, st);
  cudaError_t e = blocking_event(&ctx->mark_ev);
  return e != cudaSuccess ? e : cudaEventRecord(ctx->mark_ev, st);
}
inline cudaError_t mark_wait(mp_ctx* ctx, cudaEvent_t spin_ev) {
  return cudaEventSynchronize(blocking_waits(ctx) ? ctx->mark_ev : spin_ev);
}
}  // namespace mp
