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

__global__ void conv1_bwd_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                        int n_blocks, int OC) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OC * 10) return;
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * OC * 10 + i];
    int oc = i / 10, t = i % 10;
    if (t < 9) dw[oc * 9 + t] = s; else db[oc] = s;
}

// ------------------------------------------------------------------------------------------------
// Dense head + softmax + CCE + fused gradient.  x bf16 [B,K], wp f32 [NO][K] (packed, K in the kernel's
// (h,w,c) order), one warp per row, the whole weight matrix staged in shared memory once per CTA.
// ------------------------------------------------------------------------------------------------
template <int NO>
__global__ void __launch_bounds__(256)
dense_softmax_ce_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                        const int32_t* __restrict__ labels, float* __restrict__ probs, float* __restrict__ dlogits,
                        float* __restrict__ sample_loss, int32_t* __restrict__ correct, int B, int K, int n_out,
                        float inv_batch) {
    extern __shared__ float wsm[];                  // [n_out][K]
    for (int i = threadIdx.x * 4; i < n_out * K; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(wsm + i) = *reinterpret_cast<const float4*>(wp + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int row = blockIdx.x * nwarps + warp; row < B; row += gridDim.x * nwarps) {
        float acc[NO];
#pragma unroll
        for (int j = 0; j < NO; ++j) acc[j] = 0.f;
        const uint2* xr = reinterpret_cast<const uint2*>(x + (size_t)row * K);
        for (int k4 = lane; k4 < K / 4; k4 += 32) {
            const uint2 u = __ldg(xr + k4);
            const float x0 = bf_lo(u.x), x1 = bf_hi(u.x), x2 = bf_lo(u.y), x3 = bf_hi(u.y);
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                if (j < n_out) {
                    const float4 wv = *reinterpret_cast<const float4*>(wsm + j * K + k4 * 4);
                    acc[j] = fmaf(x0, wv.x, fmaf(x1, wv.y, fmaf(x2, wv.z, fmaf(x3, wv.w, acc[j]))));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NO; ++j)
#pragma unroll
            for (int o = 16; o; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        if (lane == 0) {
            float z[NO], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NO; ++j) if (j < n_out) { z[j] = acc[j] + bias[j]; mx = fmaxf(mx, z[j]); }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NO; ++j) if (j < n_out) { z[j] = expf(z[j] - mx); s += z[j]; }
            const int yv = labels[row];
            float best = -1.f; int bj = 0; float py = 1e-7f;
#pragma unroll
            for (int j = 0; j < NO; ++j) if (j < n_out) {
                const float pj = z[j] / s;
                probs[(size_t)row * n_out + j] = pj;
                dlogits[(size_t)row * n_out + j] = (pj - (j == yv ? 1.f : 0.f)) * inv_batch;
                if (pj > best) { best = pj; bj = j; }
                if (j == yv) py = pj;
            }
            py = fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f);
            sample_loss[row] = -logf(py);
            correct[row] = (bj == yv) ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dense dgrad + pool backward + ReLU backward: dY of the last conv, straight into its zero-haloed bf16
// layout [B, 2PH+2, 2PW+2, C].  Also the conv bias gradient (sum of the routed values) per image.
// ------------------------------------------------------------------------------------------------
template <int NO>
__global__ void __launch_bounds__(256)
dense_bwd_unpool_kernel(const float* __restrict__ dlogits, const float* __restrict__ wp, const uint8_t* __restrict__ idx,
                        __nv_bfloat16* __restrict__ dy_pad, float* __restrict__ dbias_img, int PH, int PW, int C,
                        int n_out) {
    __shared__ float dl[NO];
    __shared__ float red[256 * 8];
    const int n = blockIdx.x, K = PH * PW * C, G = C / 8;
    if (threadIdx.x < NO) dl[threadIdx.x] = threadIdx.x < n_out ? dlogits[(size_t)n * n_out + threadIdx.x] : 0.f;
    __syncthreads();
    const int H2 = 2 * PH + 2, W2 = 2 * PW + 2;
    float bsum[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bsum[c] = 0.f;
    for (int item = threadIdx.x; item < PH * PW * G; item += blockDim.x) {
        const int g = item % G, pix = item / G, py = pix / PW, px = pix % PW;
        const int k0 = pix * C + g * 8;
        float d[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) d[c] = 0.f;
#pragma unroll
        for (int j = 0; j < NO; ++j) {
            if (j < n_out) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * K + k0));
                const float4 b = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * K + k0 + 4));
                const float s = dl[j];
                d[0] = fmaf(s, a.x, d[0]); d[1] = fmaf(s, a.y, d[1]); d[2] = fmaf(s, a.z, d[2]); d[3] = fmaf(s, a.w, d[3]);
                d[4] = fmaf(s, b.x, d[4]); d[5] = fmaf(s, b.y, d[5]); d[6] = fmaf(s, b.z, d[6]); d[7] = fmaf(s, b.w, d[7]);
            }
        }
        const uint32_t nibs = *reinterpret_cast<const uint32_t*>(idx + (((size_t)n * PH + py) * PW + px) * (C / 2) + g * 4);
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
    // fixed-order block reduction of the bias-gradient partials: thread t always owns channel group t % G
#pragma unroll
    for (int c = 0; c < 8; ++c) red[threadIdx.x * 8 + c] = bsum[c];
    __syncthreads();
    if (threadIdx.x < C) {
        const int g = threadIdx.x / 8, c = threadIdx.x % 8;
        float s = 0.f;
        for (int t = g; t < 256; t += G) s += red[t * 8 + c];
        dbias_img[(size_t)n * C + threadIdx.x] = s;
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

// ------------------------------------------------------------------------------------------------
// Dense wgrad: partial[chunk][k][j] = sum_{rows in chunk} x[row][k] * dlogits[row][j]
// ------------------------------------------------------------------------------------------------
template <int NO>
__global__ void __launch_bounds__(256)
dense_wgrad_partial_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dlogits,
                           float* __restrict__ partial, int B, int K, int n_out, int rows_per_chunk) {
    extern __shared__ float dls[];             // [rows_per_chunk][NO]
    const int chunk = blockIdx.y, r0 = chunk * rows_per_chunk, r1 = min(B, r0 + rows_per_chunk);
    for (int i = threadIdx.x; i < (r1 - r0) * NO; i += blockDim.x) {
        int r = i / NO, j = i % NO;
        dls[i] = j < n_out ? dlogits[(size_t)(r0 + r) * n_out + j] : 0.f;
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > K) return;                               // k == K is a virtual all-ones feature: the bias gradient
    float acc[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) acc[j] = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float xv = k < K ? __bfloat162float(x[(size_t)r * K + k]) : 1.f;
#pragma unroll
        for (int j = 0; j < NO; ++j) acc[j] = fmaf(xv, dls[(r - r0) * NO + j], acc[j]);
    }
    float* dst = partial + ((size_t)chunk * (K + 1) + k) * NO;
#pragma unroll
    for (int j = 0; j < NO; ++j) dst[j] = acc[j];
}

// dw[perm(k)][j] = scale * sum_chunks partial + l1*sign(w) + 2*l2*w ;  db[j] = scale * sum_rows dlogits
template <int NO>
__global__ void dense_wgrad_reduce_kernel(const float* __restrict__ partial,
                                          const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ db,
                                          int n_chunks, int B, int K, int n_out, int HW, int C, float scale, float l1, float l2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K * n_out) {
        const int k = i / n_out, j = i % n_out;
        float s = 0.f;
        for (int c = 0; c < n_chunks; ++c) s += partial[((size_t)c * (K + 1) + k) * NO + j];
        const int hw = k / C, ch = k % C, kr = ch * HW + hw;     // kernel (h,w,c) order -> reference (c,h,w) row
        s *= scale;
        const float wv = w[(size_t)kr * n_out + j];
        if (l1 > 0.f) s += l1 * (wv > 0.f ? 1.f : (wv < 0.f ? -1.f : 0.f));
        if (l2 > 0.f) s += 2.f * l2 * wv;
        dw[(size_t)kr * n_out + j] = s;
    } else if (i < K * n_out + n_out) {
        const int j = i - K * n_out;
        float s = 0.f;
        for (int c = 0; c < n_chunks; ++c) s += partial[((size_t)c * (K + 1) + K) * NO + j];
        db[j] = s * scale;
    }
}

// reference (c,h,w)-row Dense weights [K][n_out] -> packed [n_out][K] with K in (h,w,c) order
__global__ void pack_dense_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int K, int n_out, int HW, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * n_out) return;
    const int j = i / K, k = i % K, hw = k / C, ch = k % C;
    wp[i] = w[(size_t)(ch * HW + hw) * n_out + j];
}

// single block, fixed order -> deterministic accumulation into double[2] {sum of losses, #correct}
__global__ void tc_loss_acc_reduce_kernel(const float* __restrict__ sample_loss, const int32_t* __restrict__ correct,
                                          double* __restrict__ accum, int B) {
    __shared__ double sl[256]; __shared__ double sc[256];
    double l = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { l += (double)sample_loss[i]; c += (double)correct[i]; }
    sl[threadIdx.x] = l; sc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) { sl[threadIdx.x] += sl[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { accum[0] += sl[0]; accum[1] += sc[0]; }
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
    conv1_bwd_reduce_kernel<<<nm_cdiv(OC * 10, 128), 128, 0, nm::S(stream)>>>(static_cast<const float*>(workspace), dw, db, blocks, OC);
    return nm::launched("conv1_bwd_reduce");
}

int nm_tc_pack_dense_weights(const float* w, float* w_packed, int K, int n_out, int HW, int C, void* stream) {
    NM_REQUIRE(w && w_packed, "null pointer");
    NM_REQUIRE(K > 0 && n_out > 0 && HW > 0 && C > 0 && HW * C == K, "bad dimensions");
    pack_dense_weights_kernel<<<nm_cdiv((size_t)K * n_out, 256), 256, 0, nm::S(stream)>>>(w, w_packed, K, n_out, HW, C);
    return nm::launched("pack_dense_weights");
}

int nm_tc_dense_softmax_ce_bf16(const void* x, const float* w_packed, const float* bias, const int32_t* labels,
                                float* probs, float* dlogits, float* sample_loss, int32_t* correct, double* accum,
                                int B, int K, int n_out, float inv_batch, void* stream) {
    NM_REQUIRE(x && w_packed && bias && labels && probs && dlogits && sample_loss && correct, "null pointer");
    NM_REQUIRE(B > 0 && K > 0 && K % 4 == 0 && n_out > 0, "bad dimensions");
    if (n_out > kNO || (size_t)n_out * K * sizeof(float) > 220 * 1024)
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_dense_softmax_ce_bf16: needs n_out <= 16 and n_out*K floats of shared memory");
    auto kern = dense_softmax_ce_kernel<kNO>;
    size_t smem = (size_t)n_out * K * sizeof(float);
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    int grid = (B + 7) / 8 < 148 ? (B + 7) / 8 : 148;
    kern<<<grid, 256, smem, nm::S(stream)>>>(static_cast<const __nv_bfloat16*>(x), w_packed, bias, labels, probs, dlogits,
                                            sample_loss, correct, B, K, n_out, inv_batch);
    int rc = nm::launched("dense_softmax_ce");
    if (rc || !accum) return rc;
    tc_loss_acc_reduce_kernel<<<1, 256, 0, nm::S(stream)>>>(sample_loss, correct, accum, B);
    return nm::launched("tc_loss_acc_reduce");
}

int nm_tc_dense_bwd_unpool_bf16(const float* dlogits, const float* w_packed, const void* idx, void* dy_pad,
                                float* dbias, void* workspace, size_t workspace_bytes,
                                int B, int PH, int PW, int C, int n_out, void* stream) {
    NM_REQUIRE(dlogits && w_packed && idx && dy_pad && dbias && workspace, "null pointer");
    NM_REQUIRE(B > 0 && PH > 0 && PW > 0 && C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && n_out > 0 && n_out <= kNO, "bad dimensions");
    NM_REQUIRE(workspace_bytes >= (size_t)B * C * sizeof(float), "workspace too small");
    dense_bwd_unpool_kernel<kNO><<<B, 256, 0, nm::S(stream)>>>(dlogits, w_packed, static_cast<const uint8_t*>(idx),
                                                             static_cast<__nv_bfloat16*>(dy_pad),
                                                             static_cast<float*>(workspace), PH, PW, C, n_out);
    int rc = nm::launched("dense_bwd_unpool");
    if (rc) return rc;
    column_sum_kernel<<<C, 256, 0, nm::S(stream)>>>(static_cast<const float*>(workspace), dbias, B, C);
    return nm::launched("column_sum");
}

size_t nm_tc_dense_wgrad_workspace(int B, int K) {
    int rows = 64, chunks = (B + rows - 1) / rows;
    return (size_t)chunks * (K + 1) * kNO * sizeof(float);
}

int nm_tc_dense_wgrad_bf16(const void* x, const float* dlogits, const float* w, float* dw, float* db,
                           void* workspace, size_t workspace_bytes, int B, int K, int n_out, int HW, int C,
                           float scale, float l1, float l2, void* stream) {
    NM_REQUIRE(x && dlogits && w && dw && db && workspace, "null pointer");
    NM_REQUIRE(B > 0 && K > 0 && n_out > 0 && n_out <= kNO && HW * C == K, "bad dimensions");
    const int rows = 64, chunks = (B + rows - 1) / rows;
    NM_REQUIRE(workspace_bytes >= (size_t)chunks * (K + 1) * kNO * sizeof(float), "workspace too small");
    dim3 grid(nm_cdiv(K + 1, 256), chunks);
    dense_wgrad_partial_kernel<kNO><<<grid, 256, rows * kNO * sizeof(float), nm::S(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), dlogits, static_cast<float*>(workspace), B, K, n_out, rows);
    int rc = nm::launched("dense_wgrad_partial");
    if (rc) return rc;
    dense_wgrad_reduce_kernel<kNO><<<nm_cdiv((size_t)K * n_out + n_out, 256), 256, 0, nm::S(stream)>>>(
        static_cast<const float*>(workspace), w, dw, db, chunks, B, K, n_out, HW, C, scale, l1, l2);
    return nm::launched("dense_wgrad_reduce");
}

}  // extern "C"
