// Host-side helpers of the tensor-core kernels: TMA tensor maps through the driver entry point, SM count.
#pragma once
#include "nm_common.cuh"
#include <cuda.h>
#include <mutex>

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int get_encoder(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
    });
    if (!fn) return nm::fail(NM_ERR_CUDA, "%s%s", "cuTensorMapEncodeTiled not available from the driver");
    *out = fn;
    return NM_OK;
}

// 2-D bf16 tensor map: rows x inner, row pitch in bytes, box (box_inner x box_rows), OOB reads as zero
inline int make_map_2d(CUtensorMap* m, const void* gptr, uint64_t inner, uint64_t rows, uint64_t pitch_bytes,
                uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc;
    int rc = get_encoder(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nm::fail(NM_ERR_CUDA, "%s%s", "cuTensorMapEncodeTiled failed");
    return NM_OK;
}

inline int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
    return n;
}

}  // namespace tc
