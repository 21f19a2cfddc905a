// Element-wise re-packing of the FP32 master weights into the operand layouts of the tensor-core kernels
// (shared by the per-layer pack kernels and by the one-launch pack of the whole CNN plan).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace tcpack {

// conv1 [OC=32][1][3][3] -> Wext[(dy,dx,oc)][a*4+b] = W[oc][a-dy][b-dx] inside the 3x3 support, else 0   (i < 128*16)
__device__ __forceinline__ float conv1_value(int i, const float* __restrict__ w) {
    const int row = i >> 4, k = i & 15, pos = row >> 5, oc = row & 31, a = k >> 2, b = k & 3;
    const int r = a - (pos >> 1), s = b - (pos & 1);
    return ((unsigned)r < 3u && (unsigned)s < 3u) ? w[oc * 9 + r * 3 + s] : 0.f;
}
__device__ __forceinline__ void conv1_elem(int i, const float* __restrict__ w, __nv_bfloat16* __restrict__ wext) {
    wext[i] = __float2bfloat16(conv1_value(i, w));
}

// OIHW f32 -> bf16 [OC][tap][C] (fprop B operand) and [C][8-tap][OC] (dgrad B operand)                    (i < OC*C*9)
__device__ __forceinline__ void conv3x3_elem(int i, const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                             __nv_bfloat16* __restrict__ wd, int OC, int C) {
    const int oc = i / (C * 9), rem = i - oc * C * 9, c = rem / 9, tap = rem - c * 9;
    const __nv_bfloat16 v = __float2bfloat16(w[i]);
    if (wf) wf[((size_t)oc * 9 + tap) * C + c] = v;
    if (wd) wd[((size_t)c * 9 + (8 - tap)) * OC + oc] = v;
}

// reference (c,h,w)-row Dense weights [K][n_out] f32 -> wp f32 [n_out][K] (K in (h,w,c) order) and
// whl bf16 [32][K]: rows [0,16) = hi, [16,32) = lo, classes >= n_out zero                                  (i < K*16)
__device__ __forceinline__ void dense_elem(int i, const float* __restrict__ w, float* __restrict__ wp, __nv_bfloat16* __restrict__ whl,
                                           int K, int n_out, int HW, int C) {
    const int j = i / K, k = i % K, hw = k / C, ch = k % C;
    const float v = j < n_out ? w[(size_t)(ch * HW + hw) * n_out + j] : 0.f;
    if (wp && j < n_out) wp[i] = v;
    if (whl) {
        const __nv_bfloat16 h = __float2bfloat16(v);
        whl[(size_t)j * K + k] = h;
        whl[(size_t)(16 + j) * K + k] = __float2bfloat16(v - __bfloat162float(h));
    }
}

}  // namespace tcpack
