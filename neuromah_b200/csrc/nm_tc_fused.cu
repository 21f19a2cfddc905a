// CUDA-core fused kernels that complete the bf16 tensor-core training path around the tcgen05 convolutions:
//   conv1 (C_in = 1, K = 9: a streaming kernel, not a GEMM) + ReLU + 2x2 max-pool, forward and backward,
//   Dense head (n_out <= 16) + softmax + cross-entropy + fused gradient,
//   Dense dgrad + max-pool backward (arg-max routing) + ReLU backward -> conv2's dY in one pass,
//   Dense wgrad with a fixed-order reduction.
// All of them are HBM-bound: coalesced 16-byte accesses, nothing materialised that a
// consumer can recompute from a nibble.
#include "nm_common.cuh"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// ------------------------------------------------------------------------------------------------
// conv1 forward: x f32 [N,1,H,W] -> ReLU(conv3x3 same, no bias) -> 2x2/2 max-pool
//   y_pad bf16 [N, H/2+2, W/2+2, OC] (interior only; the zero halo is never written)
//   idx   nibbles [N, H/2, W/2, OC/2]
// One CTA per image; a work item is (pooled pixel, group of 8 output channels).
// ------------------------------------------------------------------------------------------------
template <int OC>
__global__ void __launch_bounds__(256)
conv1_fwd_pool_kernel(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ y_pad,
                      uint8_t* __restrict__ idx, int H, int W) {
    extern __shared__ float sm[];
    const int Wp = W + 2, PH = H / 2, PW = W / 2;
    float* img = sm;                          // (H+2) x (W+2), zero halo
    float* wt = sm + (H + 2) * Wp;            // [9][OC]  (tap-major so 8 channels are contiguous)
    const int n = blockIdx.x;
    for (int i = threadIdx.x; i < (H + 2) * Wp; i += blockDim.x) {
        int yy = i / Wp - 1, xx = i % Wp - 1;
        img[i] = ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? x[((size_t)n * H + yy) * W + xx] : 0.f;
    }
    for (int i = threadIdx.x; i < OC * 9; i += blockDim.x) wt[(i % 9) * OC + i / 9] = w[i];
    __syncthreads();
    constexpr int G = OC / 8;
    for (int item = threadIdx.x; item < PH * PW * G; item += blockDim.x) {
        const int g = item % G, pix = item / G, py = pix / PW, px = pix % PW;
        float patch[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) patch[a][b] = img[(2 * py + a) * Wp + 2 * px + b];
        float s[4][8];                                   // conv value at the 4 window positions x 8 channels
#pragma unroll
        for (int pos = 0; pos < 4; ++pos)
#pragma unroll
            for (int c = 0; c < 8; ++c) s[pos][c] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 w0 = *reinterpret_cast<const float4*>(wt + t * OC + g * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(wt + t * OC + g * 8 + 4);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                const float xv = patch[(pos >> 1) + t / 3][(pos & 1) + t % 3];
#pragma unroll
                for (int c = 0; c < 8; ++c) s[pos][c] = fmaf(xv, wv[c], s[pos][c]);
            }
        }
        float best[8]; uint32_t nibs = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float bv = -1.f; uint32_t bi = 0;
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {                    // reference scan order (0,0),(0,1),(1,0),(1,1)
                const float v = fmaxf(s[pos][c], 0.f);
                if (v > bv) { bv = v; bi = pos; }                  // strict '>': first max wins
            }
            best[c] = bv;
            nibs |= (bi | (bv > 0.f ? 4u : 0u)) << (4 * c);
        }
        uint4 o = make_uint4(pack2(best[0], best[1]), pack2(best[2], best[3]), pack2(best[4], best[5]), pack2(best[6], best[7]));
        *reinterpret_cast<uint4*>(y_pad + (((size_t)n * (PH + 2) + py + 1) * (PW + 2) + px + 1) * OC + g * 8) = o;
        *reinterpret_cast<uint32_t*>(idx + (((size_t)n * PH + py) * PW + px) * (OC / 2) + g * 4) = nibs;
    }
}

// ------------------------------------------------------------------------------------------------
// conv1 backward: dW1[oc][tap], db1[oc] from the gradient of the POOLED output (bf16, un-shifted grid
// [N, PH+2, PW+2, OC]), the arg-max nibbles and the input image.  Pool backward + ReLU backward are the
// nibble; nothing of the 28x28x32 activation gradient is ever materialised.
// Per image the CTA stages x, dP and the nibbles in shared memory with 16-byte loads (all in flight at
// once -- the first version chased idx -> dP -> x per thread and was latency-bound at 3% of HBM), then a
// work item = (pooled pixel, 8 channels) accumulates 8 x (9 taps + bias) sums in registers.
// Partials per CTA, reduced in a fixed order by a second kernel (deterministic).
// ------------------------------------------------------------------------------------------------
template <int OC>
__global__ void __launch_bounds__(256, 2)
conv1_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dp_grid, const uint8_t* __restrict__ idx,
                 float* __restrict__ partial, int N, int H, int W) {
    static_assert(OC == 32, "thread mapping assumes 4 groups of 8 channels");
    extern __shared__ __align__(16) uint8_t smraw[];
    const int Wp = W + 2, PH = H / 2, PW = W / 2, GW = PW + 2, NPIX = PH * PW;
    float* img = reinterpret_cast<float*>(smraw);                              // (H+2) x (W+2), zero halo
    uint4* dps = reinterpret_cast<uint4*>(img + (((H + 2) * Wp + 3) & ~3));    // [NPIX][4] x 16 B (8 bf16)
    uint32_t* ids = reinterpret_cast<uint32_t*>(dps + NPIX * 4);               // [NPIX][4] x 4 B (8 nibbles)
    float* red = reinterpret_cast<float*>(ids + NPIX * 4);                     // [8 warps][4 groups][80]
    constexpr int G = OC / 8;
    const int g = threadIdx.x % G;
    float acc[8][10];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 10; ++t) acc[c][t] = 0.f;

    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < (H + 2) * Wp; i += blockDim.x) {
            int yy = i / Wp - 1, xx = i % Wp - 1;
            img[i] = ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? __ldg(x + ((size_t)n * H + yy) * W + xx) : 0.f;
        }
        for (int i = threadIdx.x; i < NPIX * G; i += blockDim.x) {
            const int pix = i / G, ch = i % G, py = pix / PW, px = pix % PW;
            dps[i] = __ldg(reinterpret_cast<const uint4*>(dp_grid + (((size_t)n * (PH + 2) + py) * GW + px) * OC) + ch);
            ids[i] = __ldg(reinterpret_cast<const uint32_t*>(idx + ((size_t)n * NPIX + pix) * (OC / 2)) + ch);
        }
        __syncthreads();
        for (int item = threadIdx.x; item < NPIX * G; item += blockDim.x) {
            const int pix = item / G, py = pix / PW, px = pix % PW;
            const uint32_t nibs = ids[item];
            if ((nibs & 0x44444444u) == 0) continue;                      // all 8 channels ReLU-dead
            const uint4 d = dps[item];
            const float gv[8] = {bf_lo(d.x), bf_hi(d.x), bf_lo(d.y), bf_hi(d.y), bf_lo(d.z), bf_hi(d.z), bf_lo(d.w), bf_hi(d.w)};
            float patch[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) patch[a][b] = img[(2 * py + a) * Wp + 2 * px + b];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t nb = (nibs >> (4 * c)) & 0xF;
                const float gc = (nb & 4) ? gv[c] : 0.f;                   // ReLU mask (Relu.py:12-14)
                acc[c][9] += gc;
#pragma unroll
                for (int pos = 0; pos < 4; ++pos) {                        // arg-max routing (pool backward)
                    const float gp = ((nb & 3) == (uint32_t)pos) ? gc : 0.f;
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        acc[c][t] = fmaf(gp, patch[(pos >> 1) + t / 3][(pos & 1) + t % 3], acc[c][t]);
                }
            }
        }
    }
    // fixed-order reduction: lanes with equal (lane % 4) share a channel group -> xor 4, 8, 16; then warps
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int t = 0; t < 10; ++t) {
            float v = acc[c][t];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[c][t] = v;
        }
    __syncthreads();
    if (lane < G) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int t = 0; t < 10; ++t) red[(warp * G + g) * 80 + c * 10 + t] = acc[c][t];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OC * 10; i += blockDim.x) {              // i = (g*8 + c)*10 + t
        const int gg = i / 80, rem = i % 80;
        float sum = 0.f;
        for (int w8 = 0; w8 < 8; ++w8) sum += red[(w8 * G + gg) * 80 + rem];
        partial[(size_t)blockIdx.x * OC * 10 + i] = sum;
    }
}

// 32 outputs x 8 slices of the block range per CTA; slices combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) conv1_bwd_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                               float* __restrict__ db, int n_blocks, int OC) {
    __shared__ float red[8][32];
    const int total = OC * 10;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
    const int per = (n_blocks + 7) / 8, b0 = slice * per, b1 = min(n_blocks, b0 + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < total) {
        int b = b0;
        for (; b + 4 <= b1; b += 4) {
            s0 += __ldg(partial + (size_t)b * total + i);
            s1 += __ldg(partial + (size_t)(b + 1) * total + i);
            s2 += __ldg(partial + (size_t)(b + 2) * total + i);
            s3 += __ldg(partial + (size_t)(b + 3) * total + i);
        }
        for (; b < b1; ++b) s0 += __ldg(partial + (size_t)b * total + i);
    }
    red[slice][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (slice == 0 && i < total) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        const int oc = i / 10, t = i % 10;
        if (t < 9) dw[oc * 9 + t] = s; else db[oc] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Dense dgrad + pool backward + ReLU backward: dY of the last conv, straight into its zero-haloed bf16
// layout [B, 2PH+2, 2PW+2, C].  Also the conv bias gradient (sum of the routed values), one partial per CTA.
// A thread owns ONE (pooled pixel, 8-channel group) for every image its CTA processes and keeps that
// [n_out x 8] slice of the Dense weights in registers: the weights are read once per CTA instead of once
// per image (the per-image version re-read 125 KB x 4096 images from L2 and sat at 10% of HBM).
// The next image's nibbles are prefetched while the current one is routed and stored.
// ------------------------------------------------------------------------------------------------
template <int NO>
__global__ void __launch_bounds__(416, 1)
dense_bwd_unpool_kernel(const float* __restrict__ dlogits, const float* __restrict__ wp, const uint8_t* __restrict__ idx,
                        __nv_bfloat16* __restrict__ dy_pad, float* __restrict__ dbias_cta, int B, int PH, int PW, int C,
                        int n_out) {
    extern __shared__ float red[];                       // [items][8] for the bias-gradient reduction
    const int K = PH * PW * C, G = C / 8, items = PH * PW * G;
    const int item = threadIdx.x;
    const bool active = item < items;
    const int g = item % G, pix = item / G, py = pix / PW, px = pix % PW;
    const int H2 = 2 * PH + 2, W2 = 2 * PW + 2;
    float wv[NO][8];
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (active && j < n_out) {
            a = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * K + pix * C + g * 8));
            b = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * K + pix * C + g * 8 + 4));
        }
        wv[j][0] = a.x; wv[j][1] = a.y; wv[j][2] = a.z; wv[j][3] = a.w;
        wv[j][4] = b.x; wv[j][5] = b.y; wv[j][6] = b.z; wv[j][7] = b.w;
    }
    float bsum[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bsum[c] = 0.f;
    const uint32_t* idx32 = reinterpret_cast<const uint32_t*>(idx);
    int n = blockIdx.x;
    uint32_t nibs_next = (active && n < B) ? __ldg(idx32 + (size_t)n * items + item) : 0u;
    for (; n < B; n += gridDim.x) {
        const uint32_t nibs = nibs_next;
        const int nn = n + gridDim.x;
        if (active && nn < B) nibs_next = __ldg(idx32 + (size_t)nn * items + item);
        if (!active) continue;
        float d[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) d[c] = 0.f;
#pragma unroll
        for (int j = 0; j < NO; ++j) {
            if (j < n_out) {
                const float sdl = __ldg(dlogits + (size_t)n * n_out + j);     // same address for the whole CTA: broadcast
#pragma unroll
                for (int c = 0; c < 8; ++c) d[c] = fmaf(sdl, wv[j][c], d[c]);
            }
        }
        float v[4][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t nb = (nibs >> (4 * c)) & 0xF;
            const float val = (nb & 4) ? d[c] : 0.f;                 // ReLU mask
            bsum[c] += val;
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) v[pos][c] = ((nb & 3) == (uint32_t)pos) ? val : 0.f;   // arg-max routing
        }
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) {
            const int Y = 2 * py + (pos >> 1) + 1, X = 2 * px + (pos & 1) + 1;
            uint4 o = make_uint4(pack2(v[pos][0], v[pos][1]), pack2(v[pos][2], v[pos][3]),
                                 pack2(v[pos][4], v[pos][5]), pack2(v[pos][6], v[pos][7]));
            *reinterpret_cast<uint4*>(dy_pad + (((size_t)n * H2 + Y) * W2 + X) * C + g * 8) = o;
        }
    }
    // fixed-order reduction over the pooled pixels of each channel
    if (active) {
#pragma unroll
        for (int c = 0; c < 8; ++c) red[item * 8 + c] = bsum[c];
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        const int gg = threadIdx.x / 8, c = threadIdx.x % 8;
        float s = 0.f;
        for (int p2 = 0; p2 < PH * PW; ++p2) s += red[(p2 * G + gg) * 8 + c];
        dbias_cta[(size_t)blockIdx.x * C + threadIdx.x] = s;
    }
}

// out[c] = sum_n in[n][c], fixed-shape tree -> deterministic
__global__ void column_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int C) {
    __shared__ float red[256];
    const int c = blockIdx.x;
    float s = 0.f;
    for (int n = threadIdx.x; n < N; n += 256) s += in[(size_t)n * C + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[c] = red[0];
}

constexpr int kNO = 16;

}  // namespace

extern "C" {

int nm_tc_conv1_fwd_pool_bf16(const float* x, const float* w, void* y_pad, void* idx, int N, int H, int W, int OC, void* stream) {
    NM_REQUIRE(x && w && y_pad && idx, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad dimensions");
    if (OC != 32) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv1_fwd_pool_bf16: only OC=32 is instantiated");
    size_t smem = ((size_t)(H + 2) * (W + 2) + 9 * OC) * sizeof(float);
    conv1_fwd_pool_kernel<32><<<N, 256, smem, nm::S(stream)>>>(x, w, static_cast<__nv_bfloat16*>(y_pad),
                                                              static_cast<uint8_t*>(idx), H, W);
    return nm::launched("conv1_fwd_pool");
}

size_t nm_tc_conv1_bwd_workspace(int OC) { return (size_t)296 * OC * 10 * sizeof(float); }

int nm_tc_conv1_bwd_bf16(const float* x, const void* dp_grid, const void* idx, float* dw, float* db,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int OC, void* stream) {
    NM_REQUIRE(x && dp_grid && idx && dw && db && workspace, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad dimensions");
    if (OC != 32) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv1_bwd_bf16: only OC=32 is instantiated");
    int blocks = N < 296 ? N : 296;
    NM_REQUIRE(workspace_bytes >= (size_t)blocks * OC * 10 * sizeof(float), "workspace too small");
    const int npix = (H / 2) * (W / 2);
    size_t smem = ((((size_t)(H + 2) * (W + 2) + 3) & ~(size_t)3) + 8 * 4 * 80) * sizeof(float) + (size_t)npix * 4 * (16 + 4);
    conv1_bwd_kernel<32><<<blocks, 256, smem, nm::S(stream)>>>(x, static_cast<const __nv_bfloat16*>(dp_grid),
                                                              static_cast<const uint8_t*>(idx),
                                                              static_cast<float*>(workspace), N, H, W);
    int rc = nm::launched("conv1_bwd");
    if (rc) return rc;
    conv1_bwd_reduce_kernel<<<nm_cdiv(OC * 10, 32), 256, 0, nm::S(stream)>>>(static_cast<const float*>(workspace), dw, db, blocks, OC);
    return nm::launched("conv1_bwd_reduce");
}

size_t nm_tc_dense_bwd_unpool_workspace(int C) { return (size_t)148 * 2 * C * sizeof(float); }

int nm_tc_dense_bwd_unpool_bf16(const float* dlogits, const float* w_packed, const void* idx, void* dy_pad,
                                float* dbias, void* workspace, size_t workspace_bytes,
                                int B, int PH, int PW, int C, int n_out, void* stream) {
    NM_REQUIRE(dlogits && w_packed && idx && dy_pad && dbias && workspace, "null pointer");
    NM_REQUIRE(B > 0 && PH > 0 && PW > 0 && C % 8 == 0 && n_out > 0 && n_out <= kNO, "bad dimensions");
    const int items = PH * PW * (C / 8);
    if (items > 416 || C > 416)
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_dense_bwd_unpool_bf16: needs PH*PW*C/8 <= 416 (one thread per item)");
    int grid = B < 148 ? B : 148;
    NM_REQUIRE(workspace_bytes >= (size_t)grid * C * sizeof(float), "workspace too small");
    const size_t smem = (size_t)items * 8 * sizeof(float);
    const uint8_t* ip = static_cast<const uint8_t*>(idx);
    __nv_bfloat16* dyp = static_cast<__nv_bfloat16*>(dy_pad);
    float* part = static_cast<float*>(workspace);
    if (n_out <= 10)     // the [n_out x 8] weight slice lives in registers: instantiate the common head width exactly
        dense_bwd_unpool_kernel<10><<<grid, 416, smem, nm::S(stream)>>>(dlogits, w_packed, ip, dyp, part, B, PH, PW, C, n_out);
    else
        dense_bwd_unpool_kernel<kNO><<<grid, 416, smem, nm::S(stream)>>>(dlogits, w_packed, ip, dyp, part, B, PH, PW, C, n_out);
    int rc = nm::launched("dense_bwd_unpool");
    if (rc) return rc;
    column_sum_kernel<<<C, 256, 0, nm::S(stream)>>>(static_cast<const float*>(workspace), dbias, grid, C);
    return nm::launched("column_sum");
}

}  // extern "C"
