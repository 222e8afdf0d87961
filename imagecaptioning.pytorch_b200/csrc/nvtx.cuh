// NVTX ranges around the phases of the hot path (tracing).  nvtx3 is header-only and resolves the tool's injection
// library lazily: without a profiler attached a push/pop is a branch on a null function pointer.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace capb200 {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace capb200
#define CAPB_NVTX_CAT2(a, b) a##b
#define CAPB_NVTX_CAT(a, b) CAPB_NVTX_CAT2(a, b)
#define CAPB_NVTX(name) capb200::NvtxRange CAPB_NVTX_CAT(nvtx_range_, __LINE__)(name)
