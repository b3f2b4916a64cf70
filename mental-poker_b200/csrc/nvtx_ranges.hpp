// NVTX ranges around the protocol drivers and the MSM pipeline (profiling aids), so that a
// timeline tool (Nsight Systems / Nsight Compute with --nvtx) shows which call a kernel belongs to.  NVTX v3 is
// header-only: without an attached tool every call is a load and a branch.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace mp {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
inline void nvtx_mark(const char* what) { nvtxMarkA(what); }
}  // namespace mp
