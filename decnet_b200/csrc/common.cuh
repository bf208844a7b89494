// decnet_b200/csrc/common.cuh -- shared host/device helpers for libdecnet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <utility>
#include "../../include/decnet_b200.h"

namespace decnet {

// thread-local error text behind decnet_last_error()
void set_error(const char *fmt, ...);
// records + returns DECNET_ERR_CUDA_BASE + err when err != cudaSuccess, else 0
int cuda_status(cudaError_t err, const char *what);
// checks cudaGetLastError() after a launch and bumps the per-thread launch counter
int after_launch(const char *kernel_name);
int sm_count_cached();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

#define DECNET_REQUIRE(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ::decnet::set_error(__VA_ARGS__);     \
            return DECNET_ERR_INVALID;            \
        }                                         \
    } while (0)

#define DECNET_CUDA(expr)                                         \
    do {                                                          \
        int _st = ::decnet::cuda_status((expr), #expr);           \
        if (_st) return _st;                                      \
    } while (0)

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// Programmatic dependent launch (the persistent tensor-core kernels follow one another dozens of times per step): a kernel
// launched with launch_pdl() may start while its predecessor in the stream is still running -- it sets up barriers, allocates
// TMEM, loads its weights -- and must call pdl_wait() before its first access to anything the predecessor wrote or still reads.
// pdl_wait() returns once the predecessor grid has completed and its memory operations are visible to this grid; every access
// that is ordered after the waiting thread's loads (through mbarriers) is ordered after it too.  pdl_trigger() lets the NEXT kernel
// be scheduled as soon as every CTA of this one has started.  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// true unless DECNET_PDL=0 (read once)
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ float group_max(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float group_sum(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace decnet
