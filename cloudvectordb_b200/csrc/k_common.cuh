// Shared launch helper for the k_*.cu translation units.
#pragma once
#include "gemm_topk.cuh"
#include "launchers.h"

namespace cvdb {

template <typename Kern, typename... Args>
cudaError_t launch_kernel(Kern kern, size_t smem, bool& configured, int grid, cudaStream_t st, Args... args) {
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, 256, smem, st>>>(args...);
    return cudaGetLastError();
}

#define CVDB_DISPATCH_E(E_, CALL)                          \
    switch (E_) {                                          \
        case 0: { constexpr int E = 0; return CALL; }      \
        case 1: { constexpr int E = 1; return CALL; }      \
        case 2: { constexpr int E = 2; return CALL; }      \
        case 4: { constexpr int E = 4; return CALL; }      \
        case 8: { constexpr int E = 8; return CALL; }      \
        case 16: { constexpr int E = 16; return CALL; }    \
        default: return cudaErrorInvalidValue;             \
    }

}  // namespace cvdb
