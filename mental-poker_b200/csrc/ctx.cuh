// Engine context behind the opaque `mp_ctx` of include/mpshuffle.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
}  // namespace mp
