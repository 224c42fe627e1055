// Shared helpers for the afldm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/afldm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "afldm_b200 is written for sm_100a (B200) only"
#endif

namespace afldm {

// Host-side count of kernels launched through the C ABI (afldm_launch_count()).
void note_launch(int n = 1);

// Every ABI entry point ends with this: count the launch, surface launch errors.
inline int launched(int n = 1) {
    note_launch(n);
    return (int)cudaGetLastError();
}

// Programmatic dependent launch (PDL): every kernel of this library starts with pdl_wait(), so it may be
// launched while its predecessor on the stream is still draining - launch latency and the prologue
// (barrier init, TMEM allocation, descriptor prefetch) overlap the previous kernel's tail.  Works under
// CUDA-graph capture (programmatic edges).  Opt-in with AFLDM_PDL=1.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface via cudaGetLastError()
}

#ifdef __CUDACC__
// Block until every prerequisite grid has completed and its writes are visible (no-op without PDL).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Allow the next kernel on the stream to start launching.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

inline cudaStream_t as_stream(afldm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// SiLU.  exp(-z) as ex2.approx(-z * log2 e) (rel. error ~1e-7 |z| + 2 ulp) and an approximate reciprocal
// (1 ulp): ~6 instructions instead of ~22 for expf + IEEE division, which matters because the filtered
// activation applies it 64 times per thread inside an instruction-cache-bound unrolled kernel.  Measured
// parity vs the reference's exact SiLU stays within the 1e-5 op tolerance (tests/test_gpu_ops.py).
// Written as ex2.approx / rcp.approx directly: __fdividef adds a range-scaling sequence (FSETP + 2 FMUL) that
// 1 + e^-z in [1, inf] never needs; z * rcp(inf) = -0 is the correct limit for very negative finite z.
__device__ __forceinline__ float silu_f(float z) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return z * r;
}

template <int ACT>
__device__ __forceinline__ float apply_act(float z) {
    if constexpr (ACT == AFLDM_ACT_SILU) return silu_f(z);
    return z;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the CURRENT DEVICE's copy of the function, not to the process:
// one bit per device ordinal records where it has been set (idempotent and thread-safe: two racing threads both set it).
template <typename K>
inline cudaError_t set_max_dyn_smem(K kern, int bytes, std::atomic<unsigned long long>& done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done_mask.fetch_or(bit, std::memory_order_release);
    return e;
}

}  // namespace afldm
