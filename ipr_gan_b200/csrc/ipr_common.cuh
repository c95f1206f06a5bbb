// ipr_common.cuh -- shared helpers for libipr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/ipr_b200.h"

#ifndef __CUDA_ARCH_LIST__
#endif

extern unsigned long long g_ipr_launches;   // defined in core.cu

#define IPR_REQUIRE(cond, code) do { if (!(cond)) return (code); } while (0)

// Every kernel launch goes through this: counts it and returns the launch error (never synchronises).
#define IPR_LAUNCH_CHECK() do {                                   \
        __atomic_fetch_add(&g_ipr_launches, 1ULL, __ATOMIC_RELAXED); \
        cudaError_t e_ = cudaGetLastError();                      \
        if (e_ != cudaSuccess) return (int)e_;                    \
    } while (0)

static inline cudaStream_t ipr_cu(ipr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline bool ipr_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline int ipr_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

__device__ __forceinline__ float ipr_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int ipr_warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a fixed order (deterministic); `red` holds >= 32 floats of shared memory.
// Result valid in every thread.
__device__ __forceinline__ float ipr_block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = ipr_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.0f;
    t = ipr_warp_sum(t);
    return t;
}

// Streaming 128-bit accesses that bypass L1 allocation (data touched once).
__device__ __forceinline__ float4 ipr_ldg_stream4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void ipr_stg_stream4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel of the step is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs may become resident and run their
// prologue (barrier init, TMEM allocation, tensor-map prefetch, index math) while the previous kernel drains; each
// kernel calls ipr_pdl_wait() before it touches global memory (full completion + visibility of the predecessor) and
// then ipr_pdl_trigger() so its own successor can start early.  Inside the captured CUDA graph this turns the ~250
// kernel-to-kernel launch gaps of a step into overlapped prologues.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ipr_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ipr_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t ipr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define IPR_LAUNCH_PDL(kernel, grid, block, smem, st, ...)                                         \
    do {                                                                                           \
        cudaError_t e__ = ipr_launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__); \
        if (e__ != cudaSuccess) return (int)e__;                                                   \
    } while (0)
