// Shared helpers for libneuromah_b200 (error convention, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <utility>
#include "../../include/neuromah_b200.h"

namespace nm {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

inline int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return e == cudaErrorMemoryAllocation ? NM_ERR_OOM : NM_ERR_CUDA;
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// After a kernel launch: count it and surface launch-configuration errors.
inline int launched(const char* name, int n = 1) {
    g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, name); }
    return NM_OK;
}

}  // namespace nm

#define NM_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return nm::cuda_fail(_e, #call);         \
    } while (0)

#define NM_REQUIRE(cond, msg)                                           \
    do {                                                                \
        if (!(cond)) return nm::fail(NM_ERR_BAD_ARG, "%s: %s", __func__, msg); \
    } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// The kernels of a training step are launched with programmatic stream serialization: a kernel's CTAs may be
// scheduled, and run their prologue (mbarrier init, TMEM allocation, tensor-map prefetch, shared-memory tables),
// while the previous kernel in the stream is still draining.  Contract: pdl_trigger() first thing (lets the NEXT
// kernel be scheduled as early as the hardware allows), pdl_wait() before the first access to global memory
// (it returns once every prerequisite grid has completed and its writes are visible).
namespace nm {
template <typename... KArgs, typename... Args>
inline int launch_pdl(const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, name); }
    return launched(name);
}
// same, plus a thread-block cluster shape (the CTAs of a cluster are co-scheduled and can read each other's shared memory)
template <typename... KArgs, typename... Args>
inline int launch_pdl_cluster(const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              dim3 cluster, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster.x; attr[1].val.clusterDim.y = cluster.y; attr[1].val.clusterDim.z = cluster.z;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, name); }
    return launched(name);
}
}  // namespace nm
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Timing experiments only (scratch/ablate.py): -DNM_ABLATE=<bits> builds a variant library with parts of some kernels
// switched off at COMPILE time; the product build (NM_ABLATE undefined) contains none of it.
#ifdef NM_ABLATE
#define NM_DBG(bit) (((NM_ABLATE) >> (bit)) & 1)
#else
#define NM_DBG(bit) 0
#endif

static inline unsigned nm_cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
