// nvtx_range.hpp - NVTX ranges around the host-side phases of the engine (profilers: nsys / ncu --nvtx).
// nvtx3 is header-only: without an attached tool a push/pop is a load and a predictable branch.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace cordum {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace cordum
