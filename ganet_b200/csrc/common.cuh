// Shared helpers for the ganet_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ganet_b200.h"

#define GANET_API extern "C" __attribute__((visibility("default")))

#define GANET_RETURN_IF_LAUNCH_FAILED()                  \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

#define GANET_RETURN_IF_CUDA(call)                       \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

namespace ganet {

constexpr unsigned kFullMask = 0xffffffffu;

// Lanes of a warp are laid out as lane = j * G + gl: `gl` selects one of the
// G = 32 / L scan lines the warp works on, `j` selects the depth chunk.  All
// group collectives below run over the L lanes that share `gl`.

template <int L>
__device__ __forceinline__ float group_max(float v)
{
    if constexpr (L == 32) {
        float m;   // Blackwell warp-wide fp32 max in one instruction (CREDUX.MAX.F32)
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
        return m;
    } else {
#pragma unroll
        for (int o = 16; o >= 32 / L; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
        return v;
    }
}

template <int L>
__device__ __forceinline__ int group_min(int v)
{
    if constexpr (L == 32) {
        return __reduce_min_sync(kFullMask, v);
    } else {
#pragma unroll
        for (int o = 16; o >= 32 / L; o >>= 1) v = min(v, __shfl_xor_sync(kFullMask, v, o));
        return v;
    }
}

template <int L>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int o = 16; o >= 32 / L; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// value held by the lane owning the previous / next depth chunk of the same line
template <int L>
__device__ __forceinline__ float from_prev_chunk(float v)
{
    return __shfl_up_sync(kFullMask, v, 32 / L);
}
template <int L>
__device__ __forceinline__ float from_next_chunk(float v)
{
    return __shfl_down_sync(kFullMask, v, 32 / L);
}

__device__ __forceinline__ float ld_nc(const float *p) { return __ldg(p); }

// address = 64-bit row base + 32-bit byte offset in ONE instruction (IMAD.WIDE.U32).
// Left to itself nvcc rebuilds base + (int64)(off + pix) * 4 with four integer
// instructions per access, which made the scan kernels issue-bound.
typedef unsigned long long addr_t;
template <typename T>
__device__ __forceinline__ T *at(addr_t base, unsigned off_bytes)
{
    addr_t a;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(a) : "r"(off_bytes), "l"(base));
    return reinterpret_cast<T *>(a);
}

}  // namespace ganet
