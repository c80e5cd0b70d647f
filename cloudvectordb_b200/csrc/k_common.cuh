// Shared launch helper for the k_*.cu translation units.
#pragma once
#include "gemm_topk.cuh"
#include "launchers.h"

namespace cvdb {

// `configured` is a per-kernel bit mask over device ordinals: the opt-in to > 48 KB of dynamic shared
// memory is a per-device function attribute.
template <typename Kern, typename... Args>
cudaError_t launch_kernel(Kern kern, size_t smem, unsigned long long& configured, int grid, cudaStream_t st, Args... args) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured & bit)) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        configured |= bit;
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
        case 32: { constexpr int E = 32; return CALL; }    \
        case 64: { constexpr int E = 64; return CALL; }    \
        case 128: { constexpr int E = 128; return CALL; }  \
        default: return cudaErrorInvalidValue;             \
    }

}  // namespace cvdb
