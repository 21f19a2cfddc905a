// Shared helpers for libneuromah_b200 (error convention, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <atomic>
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

// Timing experiments only (scratch/ablate.py): -DNM_ABLATE=<bits> builds a variant library with parts of some kernels
// switched off at COMPILE time; the product build (NM_ABLATE undefined) contains none of it.
#ifdef NM_ABLATE
#define NM_DBG(bit) (((NM_ABLATE) >> (bit)) & 1)
#else
#define NM_DBG(bit) 0
#endif

static inline unsigned nm_cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
