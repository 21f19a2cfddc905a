// Memory-bound glue kernels of the general bf16 plan (haloed NHWC bf16 activations, see nm_tc_conv_gen.cu):
// MaxPooling2D(2, 2) forward / backward and the Dense input gradient.  CUDA cores, 128-bit accesses, one thread per
// (pixel, 8 channels) -- these are HBM-bound (a few bytes per element and no reuse), not GEMM-shaped.
//
// Replaces: Layer_MaxPooling2D.forward / backward (src/layers/MaxPooling2D.py:37-84, maxpooling_backend.cpp:24-133,
// :145-238) together with the ReLU backward of the conv that feeds the pool (Relu.py:12-14), and Layer_Dense's
// `dinputs = dvalues . W^T` (Dense.py:108), for the tensor-core mode.
#include "nm_common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int kNO = 16;          // classes the Dense head supports (nm_tc_dense.cu)

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// x_pad [N][H+2][W+2][C] -> y [N][H/2 (+2)][W/2 (+2)][C], idx [N][H/2][W/2][C/2] nibbles:
// bits 0-1 = position of the FIRST maximum in the reference's row-major window scan (maxpooling_backend.cpp:94-108),
// bit 2 = (max > 0), i.e. the ReLU mask of the conv below at the element the gradient is routed to.
__global__ void maxpool2x2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint32_t* __restrict__ idx,
                                      int N, int H, int W, int C, int out_padded) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8, PH = H / 2, PW = W / 2, Wp = W + 2;
    const int OH = out_padded ? PH + 2 : PH, OW = out_padded ? PW + 2 : PW;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;           // element count < 2^31 (checked on the host): 32-bit index math
    if (i >= (unsigned)N * OH * OW * G) return;
    const int g = (int)(i % (unsigned)G);
    unsigned pix = i / (unsigned)G;
    const int ow = (int)(pix % (unsigned)OW); pix /= (unsigned)OW;
    const int oh = (int)(pix % (unsigned)OH);
    const int n = (int)(pix / (unsigned)OH);
    const int ph = out_padded ? oh - 1 : oh, pw = out_padded ? ow - 1 : ow;
    if (ph < 0 || ph >= PH || pw < 0 || pw >= PW) { y[i] = make_uint4(0u, 0u, 0u, 0u); return; }     // halo of the output
    const size_t base = (((size_t)n * (H + 2) + 2 * ph + 1) * Wp + 2 * pw + 1) * G + g;
    const uint4 v[4] = {__ldg(x + base), __ldg(x + base + G), __ldg(x + base + (size_t)Wp * G), __ldg(x + base + (size_t)Wp * G + G)};
    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(&v[0]);
    uint32_t out[4], nib = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {            // 32-bit word k = channels 2k, 2k+1
        float best[2] = {bf16_lo(w0[k]), bf16_hi(w0[k])};
        uint32_t pos[2] = {0u, 0u};
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            const uint32_t wq = reinterpret_cast<const uint32_t*>(&v[q])[k];
            const float c0 = bf16_lo(wq), c1 = bf16_hi(wq);
            if (c0 > best[0]) { best[0] = c0; pos[0] = q; }                 // strict: the first maximum wins
            if (c1 > best[1]) { best[1] = c1; pos[1] = q; }
        }
        out[k] = (__float_as_uint(best[0]) >> 16) | (__float_as_uint(best[1]) & 0xFFFF0000u);
        nib |= (pos[0] | (best[0] > 0.f ? 4u : 0u)) << (8 * k);
        nib |= (pos[1] | (best[1] > 0.f ? 4u : 0u)) << (8 * k + 4);
    }
    y[i] = make_uint4(out[0], out[1], out[2], out[3]);
    idx[(((size_t)n * PH + ph) * PW + pw) * G + g] = nib;
}

// dpool [N][H/2 (+2)][W/2 (+2)][C] + idx -> dy_pad [N][H+2][W+2][C]: the pooled gradient goes to the arg-max position
// of its window if the ReLU mask bit is set, every other element (halo included) is written as zero.
// Block = (one image row of (ww, 8-channel group) elements) x (a few rows); blockIdx.x walks the rows n * Hp + hh.  The flat
// one-thread-per-element index of the first version needed three 32-bit divisions by run-time values per 16-byte store and
// was INSTRUCTION-bound (block 1 of the VGG stack: 65 us for 198 MB, 0.47 of the HBM bound); here the row split is one
// division per block and the channel-group split a shift.
__global__ void maxpool2x2_bwd_kernel(const uint4* __restrict__ dpool, const uint32_t* __restrict__ idx, uint4* __restrict__ dy,
                                      int N, int H, int W, int C, int in_padded, int g_shift) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8, PH = H / 2, PW = W / 2, Hp = H + 2, Wp = W + 2;
    const int t = blockIdx.y * blockDim.x + threadIdx.x;                // (ww, g) within the row
    const int row = blockIdx.x * blockDim.y + threadIdx.y;              // n * Hp + hh
    if (t >= Wp * G || row >= N * Hp) return;
    const int ww = g_shift >= 0 ? t >> g_shift : t / G, g = t - ww * G;
    const int n = row / Hp, hh = row - n * Hp;
    const size_t i = (size_t)row * Wp * G + t;
    if (hh < 1 || hh > H || ww < 1 || ww > W) { dy[i] = make_uint4(0u, 0u, 0u, 0u); return; }
    const int ph = (hh - 1) >> 1, pw = (ww - 1) >> 1;
    const uint32_t me = (uint32_t)(((hh - 1) & 1) * 2 + ((ww - 1) & 1)) | 4u;        // nibble that selects this element
    const uint32_t nib = __ldg(idx + (((size_t)n * PH + ph) * PW + pw) * G + g);
    const size_t src = in_padded ? (((size_t)n * (PH + 2) + ph + 1) * (PW + 2) + pw + 1) * G + g
                                 : (((size_t)n * PH + ph) * PW + pw) * G + g;
    const uint4 d = __ldg(dpool + src);
    const uint32_t dv[4] = {d.x, d.y, d.z, d.w};
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t lo = ((nib >> (8 * k)) & 7u) == me ? 0x0000FFFFu : 0u;
        const uint32_t hi = ((nib >> (8 * k + 4)) & 7u) == me ? 0xFFFF0000u : 0u;
        out[k] = dv[k] & (lo | hi);
    }
    dy[i] = make_uint4(out[0], out[1], out[2], out[3]);
}

// FP32-exact mode ("bf16x3"): the activation is three bf16 planes h, m, l (`xplane` / `yplane` uint4 apart).  The window's FP32
// values (l + m) + h are compared exactly like the FP32 path compares its floats; the winner's three plane words are copied.
__global__ void maxpool2x2_fwd_x3_kernel(const uint4* __restrict__ x, size_t xplane, uint4* __restrict__ y, size_t yplane,
                                         uint32_t* __restrict__ idx, int N, int H, int W, int C, int out_padded) {
    pdl_trigger();
    pdl_wait();
    const int G = C / 8, PH = H / 2, PW = W / 2, Wp = W + 2;
    const int OH = out_padded ? PH + 2 : PH, OW = out_padded ? PW + 2 : PW;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)N * OH * OW * G) return;
    const int g = (int)(i % (unsigned)G);
    unsigned pix = i / (unsigned)G;
    const int ow = (int)(pix % (unsigned)OW); pix /= (unsigned)OW;
    const int oh = (int)(pix % (unsigned)OH);
    const int n = (int)(pix / (unsigned)OH);
    const int ph = out_padded ? oh - 1 : oh, pw = out_padded ? ow - 1 : ow;
    if (ph < 0 || ph >= PH || pw < 0 || pw >= PW) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        y[i] = z; y[yplane + i] = z; y[2 * yplane + i] = z;
        return;
    }
    const size_t base = (((size_t)n * (H + 2) + 2 * ph + 1) * Wp + 2 * pw + 1) * G + g;
    const size_t off[4] = {0, (size_t)G, (size_t)Wp * G, (size_t)Wp * G + G};
    uint32_t w[3][4][4];                                     // [plane][window position][word]
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(x + pl * xplane + base + off[q]);
            w[pl][q][0] = v.x; w[pl][q][1] = v.y; w[pl][q][2] = v.z; w[pl][q][3] = v.w;
        }
    uint32_t out[3][4], nib = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float best[2]; uint32_t pos[2] = {0u, 0u};
        best[0] = (bf16_lo(w[2][0][k]) + bf16_lo(w[1][0][k])) + bf16_lo(w[0][0][k]);
        best[1] = (bf16_hi(w[2][0][k]) + bf16_hi(w[1][0][k])) + bf16_hi(w[0][0][k]);
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            const float c0 = (bf16_lo(w[2][q][k]) + bf16_lo(w[1][q][k])) + bf16_lo(w[0][q][k]);
            const float c1 = (bf16_hi(w[2][q][k]) + bf16_hi(w[1][q][k])) + bf16_hi(w[0][q][k]);
            if (c0 > best[0]) { best[0] = c0; pos[0] = q; }                 // strict: the first maximum wins
            if (c1 > best[1]) { best[1] = c1; pos[1] = q; }
        }
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            uint32_t lo = w[pl][0][k], hi = w[pl][0][k];
#pragma unroll
            for (int q = 1; q < 4; ++q) { if (pos[0] == (uint32_t)q) lo = w[pl][q][k]; if (pos[1] == (uint32_t)q) hi = w[pl][q][k]; }
            out[pl][k] = (lo & 0x0000FFFFu) | (hi & 0xFFFF0000u);
        }
        nib |= (pos[0] | (best[0] > 0.f ? 4u : 0u)) << (8 * k);
        nib |= (pos[1] | (best[1] > 0.f ? 4u : 0u)) << (8 * k + 4);
    }
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) y[pl * yplane + i] = make_uint4(out[pl][0], out[pl][1], out[pl][2], out[pl][3]);
    idx[(((size_t)n * PH + ph) * PW + pw) * G + g] = nib;
}

// planes [3][N][HW][Cp] (unpadded NHWC grid, channels possibly zero-padded to Cp) -> float32 [N][C][HW], the reference's Flatten
// order (Flatten.py:21-53): the input of the FP32 Dense head in the FP32-exact mode
__global__ void merge3_nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ planes, size_t plane, float* __restrict__ out, int N, int HW, int C,
                                           int Cp) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;         // output index: coalesced writes
    if (i >= (size_t)N * C * HW) return;
    const int hw = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const size_t n = i / ((size_t)HW * C);
    const size_t src = (n * HW + hw) * Cp + c;
    out[i] = (__bfloat162float(planes[2 * plane + src]) + __bfloat162float(planes[plane + src])) + __bfloat162float(planes[src]);
}

// float32 [N][C][HW] -> planes [3][N][HW][Cp] (zeros in the padded channels): the Dense head's input gradient back into the
// conv stack's layout
__global__ void split3_nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ planes, size_t plane, int N, int HW, int C,
                                           int Cp) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;         // destination index
    if (i >= (size_t)N * Cp * HW) return;
    const int c = (int)(i % Cp);
    const int hw = (int)((i / Cp) % HW);
    const size_t n = i / ((size_t)HW * Cp);
    const float v = c < C ? __ldg(x + (n * C + c) * HW + hw) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16(v);
    const float r = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16(r);
    planes[i] = h; planes[plane + i] = m; planes[2 * plane + i] = __float2bfloat16(r - __bfloat162float(m));
}

// dX[b][k] = sum_j dlogits[b][j] * Wp[j][k]   (Wp = the Dense weights packed as [n_out][K] in the kernels' (h,w,c) order)
__global__ void dense_dgrad_bf16_kernel(const float* __restrict__ dlogits, const float* __restrict__ wp, const uint4* __restrict__ mask,
                                        float mask_scale, uint4* __restrict__ dx, int B, int K, int n_out) {
    pdl_trigger();
    pdl_wait();
    const int G = K / 8;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;           // B * K / 8 < 2^31 (checked on the host)
    if (i >= (unsigned)B * G) return;
    const int b = (int)(i / (unsigned)G), g = (int)(i - (unsigned)b * G);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < n_out; ++j) {
        const float d = __ldg(dlogits + (size_t)b * n_out + j);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * K + g * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * K + g * 8 + 4));
        a[0] = fmaf(d, w0.x, a[0]); a[1] = fmaf(d, w0.y, a[1]); a[2] = fmaf(d, w0.z, a[2]); a[3] = fmaf(d, w0.w, a[3]);
        a[4] = fmaf(d, w1.x, a[4]); a[5] = fmaf(d, w1.y, a[5]); a[6] = fmaf(d, w1.z, a[6]); a[7] = fmaf(d, w1.w, a[7]);
    }
    if (mask) {                       // ReLU backward of the layer that produced the Dense input: zero where its output <= 0
        const uint4 mk = __ldg(mask + i);
        const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t h = (mw[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
            a[j] = ((h & 0x8000u) || !(h & 0x7FFFu)) ? 0.f : a[j] * mask_scale;
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(a[2 * k], a[2 * k + 1]);
        o[k] = *reinterpret_cast<const uint32_t*>(&t);
    }
    dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace

extern "C" {

int nm_tc_maxpool2x2_fwd_bf16(const void* x_pad, void* y, void* idx, int N, int H, int W, int C, int out_padded, void* stream) {
    NM_REQUIRE(x_pad && y && idx, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "bad dimensions (even H, W; C % 8 == 0)");
    const long long total = (long long)N * (H / 2 + (out_padded ? 2 : 0)) * (W / 2 + (out_padded ? 2 : 0)) * (C / 8);
    NM_REQUIRE(total < (1ll << 31) - 1024, "tensor too large for 32-bit element indices");
    return nm::launch_pdl("maxpool2x2_fwd", maxpool2x2_fwd_kernel, dim3(nm_cdiv((size_t)total, 256)), dim3(256), 0, nm::S(stream),
                          static_cast<const uint4*>(x_pad), static_cast<uint4*>(y), static_cast<uint32_t*>(idx), N, H, W, C, out_padded);
}

int nm_tc_maxpool2x2_fwd_x3(const void* x_pad3, size_t x_plane, void* y3, size_t y_plane, void* idx, int N, int H, int W, int C, int out_padded,
                            void* stream) {
    NM_REQUIRE(x_pad3 && y3 && idx, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "bad dimensions (even H, W; C % 8 == 0)");
    const long long total = (long long)N * (H / 2 + (out_padded ? 2 : 0)) * (W / 2 + (out_padded ? 2 : 0)) * (C / 8);
    NM_REQUIRE(total < (1ll << 31) - 1024, "tensor too large for 32-bit element indices");
    NM_REQUIRE(x_plane >= (size_t)N * (H + 2) * (W + 2) * C && y_plane >= (size_t)total * 8 && x_plane % 8 == 0 && y_plane % 8 == 0,
               "plane strides must cover a plane and be multiples of 8 elements");
    return nm::launch_pdl("maxpool2x2_fwd_x3", maxpool2x2_fwd_x3_kernel, dim3(nm_cdiv((size_t)total, 128)), dim3(128), 0, nm::S(stream),
                          static_cast<const uint4*>(x_pad3), x_plane / 8, static_cast<uint4*>(y3), y_plane / 8, static_cast<uint32_t*>(idx),
                          N, H, W, C, out_padded);
}

int nm_tc_merge3_nhwc_to_nchw_f32(const void* planes, size_t plane_stride, float* out, int N, int HW, int C, int Cp, void* stream) {
    NM_REQUIRE(planes && out, "null pointer");
    NM_REQUIRE(N > 0 && HW > 0 && C > 0 && Cp >= C && plane_stride >= (size_t)N * HW * Cp, "bad dimensions");
    const size_t total = (size_t)N * HW * C;
    return nm::launch_pdl("merge3_nhwc_to_nchw", merge3_nhwc_to_nchw_kernel, dim3(nm_cdiv(total, 256)), dim3(256), 0, nm::S(stream),
                          static_cast<const __nv_bfloat16*>(planes), plane_stride, out, N, HW, C, Cp);
}

int nm_tc_split3_nchw_to_nhwc(const float* x, void* planes, size_t plane_stride, int N, int HW, int C, int Cp, void* stream) {
    NM_REQUIRE(planes && x, "null pointer");
    NM_REQUIRE(N > 0 && HW > 0 && C > 0 && Cp >= C && plane_stride >= (size_t)N * HW * Cp, "bad dimensions");
    const size_t total = (size_t)N * HW * Cp;
    return nm::launch_pdl("split3_nchw_to_nhwc", split3_nchw_to_nhwc_kernel, dim3(nm_cdiv(total, 256)), dim3(256), 0, nm::S(stream),
                          x, static_cast<__nv_bfloat16*>(planes), plane_stride, N, HW, C, Cp);
}

int nm_tc_maxpool2x2_bwd_bf16(const void* dpool, const void* idx, void* dy_pad, int N, int H, int W, int C, int in_padded, void* stream) {
    NM_REQUIRE(dpool && idx && dy_pad, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "bad dimensions (even H, W; C % 8 == 0)");
    const long long rows = (long long)N * (H + 2);
    NM_REQUIRE(rows < (1ll << 31) - 1024, "tensor too large for 32-bit row indices");
    const int G = C / 8, per_row = (W + 2) * G;
    int g_shift = -1;
    for (int sft = 0; sft < 16; ++sft) if ((1 << sft) == G) g_shift = sft;                 // G a power of two: shift instead of divide
    const int bx = per_row < 1024 ? per_row : 1024, by = 768 / bx > 0 ? 768 / bx : 1;       // a whole row per block where it fits
    return nm::launch_pdl("maxpool2x2_bwd", maxpool2x2_bwd_kernel, dim3(nm_cdiv((size_t)rows, by), nm_cdiv((size_t)per_row, bx)),
                          dim3(bx, by), 0, nm::S(stream), static_cast<const uint4*>(dpool), static_cast<const uint32_t*>(idx),
                          static_cast<uint4*>(dy_pad), N, H, W, C, in_padded, g_shift);
}

int nm_tc_dense_dgrad_bf16(const float* dlogits, const float* w_packed, void* dx, int B, int K, int n_out, void* stream) {
    NM_REQUIRE(dlogits && w_packed && dx, "null pointer");
    NM_REQUIRE(B > 0 && K > 0 && K % 8 == 0 && n_out > 0 && n_out <= kNO, "bad dimensions");
    const long long total = (long long)B * (K / 8);
    NM_REQUIRE(total < (1ll << 31) - 1024, "tensor too large for 32-bit element indices");
    return nm::launch_pdl("dense_dgrad_bf16", dense_dgrad_bf16_kernel, dim3(nm_cdiv((size_t)total, 256)), dim3(256), 0, nm::S(stream),
                          dlogits, w_packed, static_cast<const uint4*>(nullptr), 1.f, static_cast<uint4*>(dx), B, K, n_out);
}

int nm_tc_dense_dgrad_mask_bf16(const float* dlogits, const float* w_packed, const void* relu_mask, float mask_scale, void* dx, int B, int K,
                                int n_out, void* stream) {
    NM_REQUIRE(dlogits && w_packed && dx, "null pointer");
    NM_REQUIRE(B > 0 && K > 0 && K % 8 == 0 && n_out > 0 && n_out <= kNO, "bad dimensions");
    const long long total = (long long)B * (K / 8);
    NM_REQUIRE(total < (1ll << 31) - 1024, "tensor too large for 32-bit element indices");
    return nm::launch_pdl("dense_dgrad_mask_bf16", dense_dgrad_bf16_kernel, dim3(nm_cdiv((size_t)total, 256)), dim3(256), 0, nm::S(stream),
                          dlogits, w_packed, static_cast<const uint4*>(relu_mask), mask_scale, static_cast<uint4*>(dx), B, K, n_out);
}

}  // extern "C"
