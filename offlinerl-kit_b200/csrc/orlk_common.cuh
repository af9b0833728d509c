// Shared helpers for the orlk_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/orlk_b200.h"

namespace orlk {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline int check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define ORLK_REQUIRE(cond, msg)                         \
    do {                                                \
        if (!(cond)) {                                  \
            orlk::set_error("bad argument: %s", msg);   \
            return ORLK_ERR_BAD_ARG;                    \
        }                                               \
    } while (0)

// Programmatic dependent launch.  Every kernel of the library is launched with the stream-serialisation attribute and
// starts with pdl_enter(): it lets the NEXT kernel of the stream be scheduled right away (its CTAs become resident and
// run their own prologue) and then blocks until the PREVIOUS kernel has completed and its writes are visible.  Because
// every kernel waits on its predecessor before touching global memory, ordering stays transitive along the stream; what
// is hidden is the ~1-2 us launch + prologue latency between dependent graph nodes.  ORLK_PDL=0 turns the attribute off.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
#ifdef ORLK_PDL_WAIT_FIRST
    pdl_wait();
    pdl_trigger();
#else
    pdl_trigger();
    pdl_wait();
#endif
}

bool pdl_enabled();
// Profiling aid shared by the GEMM kernels (orlk_tc_set_trace): device buffer for per-CTA clock stamps, or NULL.
unsigned long long* trace_buffer();
void set_trace_buffer(unsigned long long* p);
// First launch of each kernel: pin its shared-memory carveout (ORLK_CARVEOUT=percent, default: leave the driver's choice).
void prepare_kernel(const void* fn);

template <typename... KArgs, typename... Args>
inline void launch_opt(bool allow_pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
    prepare_kernel(reinterpret_cast<const void*>(kernel));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (allow_pdl && pdl_enabled()) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    prepare_kernel(reinterpret_cast<const void*>(kernel));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same as launch(), with the highest scheduling priority the device offers (a kernel-node attribute under stream capture):
// when CTAs of several ready grids compete for an SM, these go first.  For short launches that finish a dependency chain
// (the per-layer optimiser updates) while long tensor-core grids of other branches still have CTAs waiting for an SM.
template <typename... KArgs, typename... Args>
inline void launch_high_priority(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    prepare_kernel(reinterpret_cast<const void*>(kernel));
    static int greatest = 1;
    if (greatest == 1) {
        int least = 0, g = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &g) != cudaSuccess) g = 0;
        greatest = g;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = greatest;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for up to 1024 threads; `red` is a __shared__ float[32]. Result is returned to every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace orlk
