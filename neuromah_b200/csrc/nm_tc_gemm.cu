// General bf16 tensor-core GEMMs for Layer_Dense (hidden layers of any width) on sm_100a: TMA-fed tcgen05 tiles with
// FP32 accumulation in TMEM, the layer's elementwise work fused into the epilogue.
//
//   fprop  Y[B, n_out]  = X[B, n_in] W[n_in, n_out] + b, ReLU            (Dense.py:79-85)      -> bf16 activations
//   dgrad  dX[B, n_in]  = dY[B, n_out] W^T, then the ReLU backward of the layer below  (Dense.py:108, Relu.py:12-14)
//   wgrad  dW[n_in, n_out] = X^T dY / B (+ L1 / L2 terms), db = sum_b dY / B           (Dense.py:96-105)
//
// fprop and dgrad are ONE kernel, D[M, N] = A[M, K] B[N, K]^T with both operands K-major: fprop takes the packed bf16 copy
// W^T [n_out][n_in], dgrad the bf16 copy of W in its own layout [n_in][n_out].  wgrad contracts over the BATCH, so both
// operands are read MN-major straight from the activation / gradient tensors ([B][features] rows): no transposed copies of
// anything that changes per step.  Dimensions need not be multiples of the tile: TMA zero-fills out-of-range boxes and the
// epilogues mask their stores (784 = 12 * 64 + 16 input features run as 13 K chunks).
//
// FP32-exact on the tensor cores (mode "bf16x3"): every operand is carried as THREE bf16 planes h + m + l (h = bf16(v),
// m = bf16(v - h), l = bf16(v - h - m): 24 mantissa bits) and the product is the six plane pairs whose weight is above
// 2^-24 -- (l,h) (h,l) (m,m) (m,h) (h,m) (h,h), smallest first -- accumulated in the same FP32 TMEM tile.  The kernels
// below run them as six passes over the K chunks with the plane chosen per pass (template parameter SPLIT); the
// epilogues split their FP32 results back into planes.  rtol 1e-4 / atol 1e-5 against the NumPy path, like the CUDA-core
// FP32 kernels (tests/test_gpu_tc_mlp.py).
//
// Replaces for the tensor-core mode: np.dot(inputs, weights) + biases and the embedded activation (Dense.py:79-85),
// np.dot(inputs.T, dvalues) / B, np.sum(dvalues) / B, the regulariser terms and np.dot(dvalues, weights.T)
// (Dense.py:96-108), Activation_ReLU.forward / backward (Relu.py:7-14).
#include "nm_common.cuh"
#include "nm_tc_common.cuh"
#include "nm_tc_host.cuh"
#include "nm_philox.cuh"

namespace {

using namespace tc;

constexpr int GM = 128;                  // rows of one MMA / CTA tile
constexpr int GK = 64;                   // K chunk = one 128-byte SWIZZLE_128B row
constexpr int G_STAGES = 4;

struct GemmParams {
    int M, N, K;                         // D[M][N] = sum_k A[M][k] * B[N][k]
    int k_chunks, n_tiles;
    __nv_bfloat16* out; int ldo;         // bf16 [M][ldo]
    size_t out_plane;                    // SPLIT: elements between the h / m / l planes of `out`
    float* out_f32;                      // SPLIT, optional: FP32 copy of the result [M][ldo]
    const float* bias;                   // optional [N]
    int relu;
    const __nv_bfloat16* mask; int ldm;  // optional [M][ldm]: out = mask > 0 ? v * mask_scale : 0   (ReLU' of the layer below,
    float mask_scale;                    //   times 1 / keep when a Layer_Dropout sits between the two layers)
    float drop_keep;                     // < 1: Layer_Dropout fused behind the activation: out *= Bernoulli(keep) / keep
    unsigned long long drop_seed, drop_offset;      // Philox key / stream offset (same stream as nm_dropout_fwd_f32)
    const long long* drop_offset_dev;    // optional: added to drop_offset on the device (the optimizer's step counter, so that
};                                       //   a CUDA-graph replay of the step draws a fresh mask)

struct Maps3 { CUtensorMap m[3]; };      // one tensor map per plane (bf16 mode: only [0])
template <int NT, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
dense_gemm_tc_kernel(const __grid_constant__ Maps3 maps_a, const __grid_constant__ Maps3 maps_b, GemmParams p) {
    constexpr int PASSES = SPLIT ? 6 : 1;
    constexpr uint64_t TMPL = desc_template(16, 8 * 128, LAYOUT_SW128);         // K-major, SBO = 8 rows of 128 B
    constexpr uint32_t IDESC = idesc_bf16(GM, NT, 0, 0);
    constexpr int A_BYTES = GM * 128, B_BYTES = NT * 128, STAGE = A_BYTES + B_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G_STAGES * STAGE);
    uint64_t* empty = full + G_STAGES;
    uint64_t* tfull = empty + G_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x / p.n_tiles, n0 = (blockIdx.x - m_tile * p.n_tiles) * NT;
    pdl_trigger();
    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int pl = 0; pl < (SPLIT ? 3 : 1); ++pl) { tma_prefetch_desc(&maps_a.m[pl]); tma_prefetch_desc(&maps_b.m[pl]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, NT);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int pr = 0; pr < PASSES; ++pr) {
                const CUtensorMap* ma = &maps_a.m[SPLIT ? split_plane_a(pr) : 0];
                const CUtensorMap* mb = &maps_b.m[SPLIT ? split_plane_b(pr) : 0];
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + stage * STAGE;
                    mbar_expect_tx(&full[stage], STAGE);             // out-of-range parts of a box arrive as zeros, bytes included
                    tma_load_2d(st, ma, kc * GK, m_tile * GM, &full[stage]);
                    tma_load_2d(st + A_BYTES, mb, kc * GK, n0, &full[stage]);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int kc = 0; kc < PASSES * p.k_chunks; ++kc) {
                mbar_wait(&full[stage], phase);
                fence_after_sync();
                const uint32_t st = smem_u32(smem + stage * STAGE);
                const uint64_t a_desc = desc_at(TMPL, st), b_desc = desc_at(TMPL, st + A_BYTES);
#pragma unroll
                for (int kk = 0; kk < GK / 16; ++kk)
                    mma_bf16(tmem_base, a_desc + ((kk * 32) >> 4), b_desc + ((kk * 32) >> 4), IDESC, (kc | kk) != 0);
                mma_commit(&empty[stage]);
                if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
            }
            mma_commit(tfull);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = m_tile * GM + q * 32 + lane;
        mbar_wait(tfull, 0);
        fence_after_sync();
        const bool row_ok = row < p.M;
        const unsigned long long drop_dev = p.drop_offset_dev ? (unsigned long long)__ldg(p.drop_offset_dev) : 0ull;
        __nv_bfloat16* dst = p.out + (size_t)row * p.ldo + n0;
        const __nv_bfloat16* msk = p.mask ? p.mask + (size_t)row * p.ldm + n0 : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 32) {
            uint32_t v[32];
            tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c0 + 16, v + 16);
            tmem_ld_wait();
            if (!row_ok) continue;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int col = n0 + c0 + g * 8;
                if (col >= p.N) break;                                  // N is a multiple of 8 (checked on the host)
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
                if (p.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                }
                if (p.drop_keep < 1.f) {               // Dropout.py:27-30 behind the activation; element index = row * N + col
                    const unsigned long long e = (unsigned long long)row * p.N + col, off = p.drop_offset + drop_dev;
                    const uint32_t kb = dropout_keep4(e >> 2, p.drop_seed, off, p.drop_keep) |
                                        (dropout_keep4((e >> 2) + 1, p.drop_seed, off, p.drop_keep) << 4);
                    const float inv = 1.f / p.drop_keep;
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = (kb >> j) & 1u ? f[j] * inv : 0.f;
                }
                if (msk) {
                    const uint4 mk = __ldg(reinterpret_cast<const uint4*>(msk + c0 + g * 8));
                    const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t h = (mw[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;     // bf16 > 0: sign clear, not zero
                        f[j] = ((h & 0x8000u) || !(h & 0x7FFFu)) ? 0.f : f[j] * p.mask_scale;
                    }
                }
                if constexpr (SPLIT) {
                    float h[8], m[8], l[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) split3(f[j], h[j], m[j], l[j]);
                    __nv_bfloat16* d3 = dst + c0 + g * 8;
                    *reinterpret_cast<uint4*>(d3) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                               pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
                    *reinterpret_cast<uint4*>(d3 + p.out_plane) = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]),
                                                                             pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
                    *reinterpret_cast<uint4*>(d3 + 2 * p.out_plane) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]),
                                                                                 pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
                    if (p.out_f32) {
                        float4* o = reinterpret_cast<float4*>(p.out_f32 + (size_t)row * p.ldo + col);
                        o[0] = make_float4(f[0], f[1], f[2], f[3]);
                        o[1] = make_float4(f[4], f[5], f[6], f[7]);
                    }
                } else {
                    *reinterpret_cast<uint4*>(dst + c0 + g * 8) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                             pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, NT);
}

// ---------------------------------------------------------------------------------------------
// wgrad: P[split][f][n] = sum over the split's batch blocks of X[b][f] * dY[b][n]        (both operands MN-major)
// ---------------------------------------------------------------------------------------------
struct DenseWgradParams {
    int F, N;                            // features (n_in), outputs (n_out)
    int k_blocks, splits, n_tiles;
    float* partial;                      // [splits][F][N]
};

template <int NT, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
dense_wgrad_gen_tc_kernel(const __grid_constant__ Maps3 maps_x, const __grid_constant__ Maps3 maps_dy, DenseWgradParams p) {
    constexpr int PASSES = SPLIT ? 6 : 1;
    constexpr int BOX = GM * 128;                                        // one [128 batch rows][64 columns] box, 16 KB
    constexpr int A_BYTES = 2 * BOX, B_BYTES = (NT / 64) * BOX, STAGE = A_BYTES + B_BYTES;
    // MN-major: 64-column chunks BOX bytes apart (leading byte offset), 8 batch rows = 1024 B (stride byte offset)
    constexpr uint64_t TMPL = desc_template(BOX, 1024, LAYOUT_SW128);
    constexpr uint32_t IDESC = idesc_bf16(GM, NT, 1, 1);
    constexpr int STAGES = NT == 128 ? 3 : 4;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* done = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f_tile = blockIdx.x / p.n_tiles, n0 = (blockIdx.x - f_tile * p.n_tiles) * NT, f0 = f_tile * GM;
    const int split = blockIdx.y;
    pdl_trigger();
    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int pl = 0; pl < (SPLIT ? 3 : 1); ++pl) { tma_prefetch_desc(&maps_x.m[pl]); tma_prefetch_desc(&maps_dy.m[pl]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, NT);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const bool has_work = split < p.k_blocks;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int pr = 0; pr < PASSES; ++pr) {
                const CUtensorMap* mx = &maps_x.m[SPLIT ? split_plane_a(pr) : 0];
                const CUtensorMap* mdy = &maps_dy.m[SPLIT ? split_plane_b(pr) : 0];
                for (int kb = split; kb < p.k_blocks; kb += p.splits) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + stage * STAGE;
                    mbar_expect_tx(&full[stage], STAGE);
                    tma_load_2d(st, mx, f0, kb * GM, &full[stage]);
                    tma_load_2d(st + BOX, mx, f0 + 64, kb * GM, &full[stage]);
#pragma unroll
                    for (int j = 0; j < NT / 64; ++j) tma_load_2d(st + A_BYTES + j * BOX, mdy, n0 + j * 64, kb * GM, &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0; bool first = true;
            for (int pr = 0; pr < PASSES; ++pr)
            for (int kb = split; kb < p.k_blocks; kb += p.splits) {
                mbar_wait(&full[stage], phase);
                fence_after_sync();
                const uint32_t st = smem_u32(smem + stage * STAGE);
                const uint64_t a_desc = desc_at(TMPL, st), b_desc = desc_at(TMPL, st + A_BYTES);
#pragma unroll
                for (int ks = 0; ks < GM / 16; ++ks)
                    mma_bf16(tmem_base, a_desc + ((ks * 16 * 128) >> 4), b_desc + ((ks * 16 * 128) >> 4), IDESC, !(first && ks == 0));
                first = false;
                mma_commit(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (has_work) mma_commit(done);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int f = f0 + q * 32 + lane;
        if (has_work) { mbar_wait(done, 0); fence_after_sync(); }
        float* dst = p.partial + ((size_t)split * p.F + f) * p.N + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t v[16];
            if (has_work) { tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v); tmem_ld_wait(); }
            if (f < p.F && n0 + c0 < p.N) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + c0 + j) = has_work
                        ? make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, NT);
}

// dW[f][n] = scale * sum_s P[s][f][n] + l1 sign(W) + 2 l2 W      (splits summed in order: bitwise repeatable)
__global__ void __launch_bounds__(256)
dense_wgrad_gen_finish_kernel(const float* __restrict__ partial, const float* __restrict__ w, float* __restrict__ dw,
                              int F, int N, int splits, float scale, float l1, float l2) {
    pdl_trigger();
    pdl_wait();
    const size_t total = (size_t)F * N / 4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + (size_t)s * total + i);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    if (l1 != 0.f || l2 != 0.f) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + i);
        auto reg = [&](float x) { return l1 * (x > 0.f ? 1.f : x < 0.f ? -1.f : 0.f) + 2.f * l2 * x; };
        a.x += reg(ww.x); a.y += reg(ww.y); a.z += reg(ww.z); a.w += reg(ww.w);
    }
    reinterpret_cast<float4*>(dw)[i] = a;
}

// db[n] = scale * sum_b dY[b][n] over the bf16 gradient [B][N] (`planes` planes `plane` elements apart, summed per element):
// one block per 32 columns, 8 row lanes, fixed tree
__global__ void __launch_bounds__(256)
dense_dbias_bf16_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db, int B, int N, float scale, int planes, size_t plane) {
    __shared__ float red[8][32];
    pdl_trigger();
    pdl_wait();
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    float s = 0.f;
    if (col < N)
        for (int b = rl; b < B; b += 8) {
            float v = __bfloat162float(dy[(size_t)b * N + col]);
            for (int pl = 1; pl < planes; ++pl) v += __bfloat162float(dy[(size_t)pl * plane + (size_t)b * N + col]);
            s += v;
        }
    red[rl][threadIdx.x & 31] = s;
    __syncthreads();
    if (rl == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        db[col] = t * scale;
    }
}

// float32 (or raw uint8 pixels * scale) -> bf16, 8 elements per thread: the A operand of the first Dense layer
template <typename XT>
__global__ void to_bf16_kernel(const XT* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n, float scale) {
    pdl_trigger();
    pdl_wait();
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = i + j < n ? (float)x[i + j] * scale : 0.f;
    if (i + 8 <= n)
        *reinterpret_cast<uint4*>(y + i) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    else
        for (int j = 0; i + j < n; ++j) y[i + j] = __float2bfloat16(f[j]);
}

// float32 (or raw uint8 pixels / denom) -> the three bf16 planes h, m, l (`plane` elements apart); with `mask` (bf16, same
// indexing): zero where mask <= 0, x mask_scale elsewhere (the ReLU backward of the layer the gradient enters)
template <typename XT>
__global__ void split3_kernel(const XT* __restrict__ x, float denom, const __nv_bfloat16* __restrict__ mask, float mask_scale,
                              __nv_bfloat16* __restrict__ y, size_t n, size_t plane) {
    pdl_trigger();
    pdl_wait();
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= n) return;
    float h[8], m[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = 0.f;
        if (i0 + j < n) {
            if constexpr (sizeof(XT) == 1) v = (float)x[i0 + j] / denom; else v = (float)x[i0 + j];
            if (mask) v = __bfloat162float(mask[i0 + j]) > 0.f ? v * mask_scale : 0.f;
        }
        split3(v, h[j], m[j], l[j]);
    }
    if (i0 + 8 <= n && plane % 8 == 0) {
        *reinterpret_cast<uint4*>(y + i0) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        *reinterpret_cast<uint4*>(y + plane + i0) = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
        *reinterpret_cast<uint4*>(y + 2 * plane + i0) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    } else {
        for (int j = 0; i0 + j < n && j < 8; ++j) {
            y[i0 + j] = __float2bfloat16(h[j]); y[plane + i0 + j] = __float2bfloat16(m[j]); y[2 * plane + i0 + j] = __float2bfloat16(l[j]);
        }
    }
}

// Dense weights f32 [n_in][n_out] -> three bf16 planes in the same layout and three transposed [n_out][n_in]
__global__ void pack_dense_gen_x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_kn, __nv_bfloat16* __restrict__ w_nk,
                                         int K, int N) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const int k = i / N, n = i - k * N;
    float h, m, l;
    split3(w[i], h, m, l);
    const size_t plane = (size_t)K * N, t = (size_t)n * K + k;
    if (w_kn) { w_kn[i] = __float2bfloat16(h); w_kn[plane + i] = __float2bfloat16(m); w_kn[2 * plane + i] = __float2bfloat16(l); }
    if (w_nk) { w_nk[t] = __float2bfloat16(h); w_nk[plane + t] = __float2bfloat16(m); w_nk[2 * plane + t] = __float2bfloat16(l); }
}

// Dense weights f32 [n_in][n_out] -> bf16 copy in the same layout (dgrad operand) and transposed [n_out][n_in] (fprop operand)
__global__ void pack_dense_gen_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_kn, __nv_bfloat16* __restrict__ w_nk,
                                      int K, int N) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const int k = i / N, n = i - k * N;
    const __nv_bfloat16 v = __float2bfloat16(w[i]);
    if (w_kn) w_kn[i] = v;
    if (w_nk) w_nk[(size_t)n * K + k] = v;
}

// tensor maps of the planes of a bf16 [rows][inner] matrix (`planes` = 1 or 3, `plane` elements apart)
int make_maps(Maps3* m, const void* base, int planes, size_t plane, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
    for (int pl = 0; pl < planes; ++pl) {
        const int rc = make_map_2d(&m->m[pl], static_cast<const __nv_bfloat16*>(base) + (size_t)pl * plane, inner, rows, inner * 2,
                                   box_inner, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    for (int pl = planes; pl < 3; ++pl) m->m[pl] = m->m[0];
    return NM_OK;
}

template <int NT, bool SPLIT>
int launch_gemm(const void* a, size_t a_plane, const void* b, size_t b_plane, GemmParams p, cudaStream_t st) {
    Maps3 ma, mb;
    int rc = make_maps(&ma, a, SPLIT ? 3 : 1, a_plane, (uint64_t)p.K, (uint64_t)p.M, GK, GM);
    if (rc) return rc;
    rc = make_maps(&mb, b, SPLIT ? 3 : 1, b_plane, (uint64_t)p.K, (uint64_t)p.N, GK, NT);
    if (rc) return rc;
    const size_t smem = (size_t)G_STAGES * (GM * 128 + NT * 128) + 256;
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(dense_gemm_tc_kernel<NT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    p.n_tiles = (p.N + NT - 1) / NT;
    const int m_tiles = (p.M + GM - 1) / GM;
    return nm::launch_pdl(SPLIT ? "dense_gemm_tc_x3" : "dense_gemm_tc", dense_gemm_tc_kernel<NT, SPLIT>, dim3(m_tiles * p.n_tiles), dim3(256), smem, st, ma, mb, p);
}
template <bool SPLIT>
int launch_gemm_nt(const void* a, size_t a_plane, const void* b, size_t b_plane, const GemmParams& p, cudaStream_t st) {
    return p.N % 128 == 0 ? launch_gemm<128, SPLIT>(a, a_plane, b, b_plane, p, st) : launch_gemm<64, SPLIT>(a, a_plane, b, b_plane, p, st);
}

struct DenseWgradPlan { int nt, f_tiles, n_tiles, k_blocks, splits; };
DenseWgradPlan dense_wgrad_plan(int B, int F, int N) {
    DenseWgradPlan g;
    g.nt = N % 128 == 0 ? 128 : 64;
    g.f_tiles = (F + GM - 1) / GM;
    g.n_tiles = (N + g.nt - 1) / g.nt;
    g.k_blocks = (B + GM - 1) / GM;
    int s = sm_count() / (g.f_tiles * g.n_tiles);
    if (s < 1) s = 1;
    if (s > g.k_blocks) s = g.k_blocks;
    g.splits = s;
    return g;
}

size_t wgrad_gen_workspace(int B, int n_in, int n_out) {
    if (B <= 0 || n_in <= 0 || n_out <= 0) return 0;
    const DenseWgradPlan g = dense_wgrad_plan(B, n_in, n_out);
    return (size_t)g.splits * n_in * n_out * sizeof(float);
}

template <int NT, bool SPLIT>
int launch_wgrad_kernel(const Maps3& mx, const Maps3& mdy, const DenseWgradParams& p, dim3 grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(dense_wgrad_gen_tc_kernel<NT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const size_t smem = NT == 128 ? (size_t)3 * (4 * GM * 128) + 256 : (size_t)4 * (3 * GM * 128) + 256;
    return nm::launch_pdl(SPLIT ? "dense_wgrad_gen_tc_x3" : "dense_wgrad_gen_tc", dense_wgrad_gen_tc_kernel<NT, SPLIT>, grid, dim3(256), smem, st, mx, mdy, p);
}

// dw[n_in][n_out] = scale * x^T dy (+ l1 sign(w) + 2 l2 w), db[n_out] = scale * sum_b dy; x, dy bf16 [B][.] in 1 or 3 planes
template <bool SPLIT>
int dense_wgrad_gen(const void* x, size_t x_plane, const void* dy, size_t dy_plane, const float* w, float* dw, float* db, void* workspace,
                    size_t workspace_bytes, int B, int n_in, int n_out, float scale, float l1, float l2, cudaStream_t st) {
    NM_REQUIRE(x && dy && dw && workspace, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0 && n_in % 8 == 0 && n_out % 64 == 0, "bad dimensions (n_in % 8, n_out % 64)");
    NM_REQUIRE(w || (l1 == 0.f && l2 == 0.f), "the regulariser terms need the weights");
    const DenseWgradPlan g = dense_wgrad_plan(B, n_in, n_out);
    NM_REQUIRE(workspace_bytes >= wgrad_gen_workspace(B, n_in, n_out), "workspace too small");
    DenseWgradParams p{};
    p.F = n_in; p.N = n_out; p.k_blocks = g.k_blocks; p.splits = g.splits; p.n_tiles = g.n_tiles;
    p.partial = static_cast<float*>(workspace);
    Maps3 mx, mdy;
    int rc = make_maps(&mx, x, SPLIT ? 3 : 1, x_plane, (uint64_t)n_in, (uint64_t)B, 64, GM);
    if (rc) return rc;
    rc = make_maps(&mdy, dy, SPLIT ? 3 : 1, dy_plane, (uint64_t)n_out, (uint64_t)B, 64, GM);
    if (rc) return rc;
    const dim3 grid(g.f_tiles * g.n_tiles, g.splits);
    rc = g.nt == 128 ? launch_wgrad_kernel<128, SPLIT>(mx, mdy, p, grid, st) : launch_wgrad_kernel<64, SPLIT>(mx, mdy, p, grid, st);
    if (rc) return rc;
    NM_REQUIRE(((size_t)n_in * n_out) % 4 == 0, "n_in * n_out must be a multiple of 4");
    rc = nm::launch_pdl("dense_wgrad_gen_finish", dense_wgrad_gen_finish_kernel, dim3(nm_cdiv((size_t)n_in * n_out / 4, 256)), dim3(256), 0,
                        st, (const float*)p.partial, w, dw, n_in, n_out, g.splits, scale, l1, l2);
    if (rc || !db) return rc;
    return nm::launch_pdl("dense_dbias_bf16", dense_dbias_bf16_kernel, dim3(nm_cdiv(n_out, 32)), dim3(256), 0, st,
                          static_cast<const __nv_bfloat16*>(dy), db, B, n_out, scale, SPLIT ? 3 : 1, dy_plane);
}

}  // namespace

extern "C" {

int nm_tc_to_bf16(const float* x, void* y, size_t n, void* stream) {
    NM_REQUIRE(x && y, "null pointer");
    if (!n) return NM_OK;
    return nm::launch_pdl("to_bf16", to_bf16_kernel<float>, dim3(nm_cdiv(n, 8 * 256)), dim3(256), 0, nm::S(stream), x,
                          static_cast<__nv_bfloat16*>(y), n, 1.f);
}
int nm_tc_u8_to_bf16(const uint8_t* x, float scale, void* y, size_t n, void* stream) {
    NM_REQUIRE(x && y, "null pointer");
    if (!n) return NM_OK;
    return nm::launch_pdl("u8_to_bf16", to_bf16_kernel<uint8_t>, dim3(nm_cdiv(n, 8 * 256)), dim3(256), 0, nm::S(stream), x,
                          static_cast<__nv_bfloat16*>(y), n, scale);
}

// x (float32) -> bf16 planes h, m, l at planes + {0, 1, 2} * plane_stride elements (h + m + l == x to 24 bits); with relu_mask
// (bf16 [n]): zero where the mask is <= 0, x * mask_scale elsewhere
int nm_tc_split3_f32(const float* x, const void* relu_mask, float mask_scale, void* planes, size_t n, size_t plane_stride, void* stream) {
    NM_REQUIRE(x && planes, "null pointer");
    NM_REQUIRE(plane_stride >= n, "plane_stride must cover the n elements of a plane");
    if (!n) return NM_OK;
    return nm::launch_pdl("split3_f32", split3_kernel<float>, dim3(nm_cdiv(n, 8 * 256)), dim3(256), 0, nm::S(stream), x, 1.f,
                          static_cast<const __nv_bfloat16*>(relu_mask), mask_scale, static_cast<__nv_bfloat16*>(planes), n, plane_stride);
}
// raw uint8 pixels: planes of float32(x) / denom (the division Layer_Input's FP32 path performs, nm_u8_to_f32_div)
int nm_tc_split3_u8(const uint8_t* x, float denom, void* planes, size_t n, size_t plane_stride, void* stream) {
    NM_REQUIRE(x && planes, "null pointer");
    NM_REQUIRE(plane_stride >= n && denom != 0.f, "plane_stride must cover the n elements of a plane; denom != 0");
    if (!n) return NM_OK;
    return nm::launch_pdl("split3_u8", split3_kernel<uint8_t>, dim3(nm_cdiv(n, 8 * 256)), dim3(256), 0, nm::S(stream), x, denom,
                          static_cast<const __nv_bfloat16*>(nullptr), 1.f, static_cast<__nv_bfloat16*>(planes), n, plane_stride);
}

int nm_tc_pack_dense_gen_bf16(const float* w, void* w_kn, void* w_nk, int n_in, int n_out, void* stream) {
    NM_REQUIRE(w && (w_kn || w_nk), "null pointer");
    NM_REQUIRE(n_in > 0 && n_out > 0, "bad dimensions");
    return nm::launch_pdl("pack_dense_gen", pack_dense_gen_kernel, dim3(nm_cdiv((size_t)n_in * n_out, 256)), dim3(256), 0, nm::S(stream),
                          w, static_cast<__nv_bfloat16*>(w_kn), static_cast<__nv_bfloat16*>(w_nk), n_in, n_out);
}
// the same as three planes each: w_kn3 [3][n_in][n_out], w_nk3 [3][n_out][n_in]
int nm_tc_pack_dense_gen_x3(const float* w, void* w_kn3, void* w_nk3, int n_in, int n_out, void* stream) {
    NM_REQUIRE(w && (w_kn3 || w_nk3), "null pointer");
    NM_REQUIRE(n_in > 0 && n_out > 0, "bad dimensions");
    return nm::launch_pdl("pack_dense_gen_x3", pack_dense_gen_x3_kernel, dim3(nm_cdiv((size_t)n_in * n_out, 256)), dim3(256), 0, nm::S(stream),
                          w, static_cast<__nv_bfloat16*>(w_kn3), static_cast<__nv_bfloat16*>(w_nk3), n_in, n_out);
}

// y[B][n_out] = x[B][n_in] w_nk[n_out][n_in]^T (+ bias) (ReLU) (x Bernoulli(drop_keep) / drop_keep: a fused Layer_Dropout)
int nm_tc_dense_fprop_gen_bf16(const void* x, const void* w_nk, const float* bias, void* y, int B, int n_in, int n_out, int relu,
                               float drop_keep, unsigned long long drop_seed, unsigned long long drop_offset,
                               const long long* drop_offset_dev, void* stream) {
    NM_REQUIRE(x && w_nk && y, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0 && n_in % 8 == 0 && n_out % 8 == 0, "bad dimensions (n_in, n_out multiples of 8)");
    NM_REQUIRE(drop_keep > 0.f && drop_keep <= 1.f, "drop_keep (the KEEP probability) must be in (0, 1]");
    GemmParams p{};
    p.M = B; p.N = n_out; p.K = n_in; p.k_chunks = (n_in + GK - 1) / GK;
    p.out = static_cast<__nv_bfloat16*>(y); p.ldo = n_out; p.bias = bias; p.relu = relu;
    p.mask_scale = 1.f; p.drop_keep = drop_keep; p.drop_seed = drop_seed; p.drop_offset = drop_offset; p.drop_offset_dev = drop_offset_dev;
    return launch_gemm_nt<false>(x, 0, w_nk, 0, p, nm::S(stream));
}
// FP32-exact variant: x3 / y3 are three bf16 planes x_plane / y_plane elements apart, w_nk3 [3][n_out][n_in]; y_f32 (optional)
// receives the FP32 result as well
int nm_tc_dense_fprop_gen_x3(const void* x3, size_t x_plane, const void* w_nk3, const float* bias, void* y3, size_t y_plane, float* y_f32,
                             int B, int n_in, int n_out, int relu, float drop_keep, unsigned long long drop_seed,
                             unsigned long long drop_offset, const long long* drop_offset_dev, void* stream) {
    NM_REQUIRE(x3 && w_nk3 && y3, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0 && n_in % 8 == 0 && n_out % 8 == 0, "bad dimensions (n_in, n_out multiples of 8)");
    NM_REQUIRE(x_plane >= (size_t)B * n_in && y_plane >= (size_t)B * n_out && x_plane % 8 == 0 && y_plane % 8 == 0,
               "plane strides must cover a [B][features] plane and be multiples of 8 elements");
    NM_REQUIRE(drop_keep > 0.f && drop_keep <= 1.f, "drop_keep (the KEEP probability) must be in (0, 1]");
    GemmParams p{};
    p.M = B; p.N = n_out; p.K = n_in; p.k_chunks = (n_in + GK - 1) / GK;
    p.out = static_cast<__nv_bfloat16*>(y3); p.ldo = n_out; p.out_plane = y_plane; p.out_f32 = y_f32; p.bias = bias; p.relu = relu;
    p.mask_scale = 1.f; p.drop_keep = drop_keep; p.drop_seed = drop_seed; p.drop_offset = drop_offset; p.drop_offset_dev = drop_offset_dev;
    return launch_gemm_nt<true>(x3, x_plane, w_nk3, (size_t)n_in * n_out, p, nm::S(stream));
}

// dx[B][n_in] = dy[B][n_out] w_kn[n_in][n_out]^T; where relu_mask[B][n_in] is given: zero where it is <= 0, x mask_scale elsewhere
int nm_tc_dense_dgrad_gen_bf16(const void* dy, const void* w_kn, const void* relu_mask, float mask_scale, void* dx, int B, int n_in,
                               int n_out, void* stream) {
    NM_REQUIRE(dy && w_kn && dx, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0 && n_in % 8 == 0 && n_out % 8 == 0, "bad dimensions (n_in, n_out multiples of 8)");
    GemmParams p{};
    p.M = B; p.N = n_in; p.K = n_out; p.k_chunks = (n_out + GK - 1) / GK;
    p.out = static_cast<__nv_bfloat16*>(dx); p.ldo = n_in;
    p.mask = static_cast<const __nv_bfloat16*>(relu_mask); p.ldm = n_in; p.mask_scale = mask_scale; p.drop_keep = 1.f;
    return launch_gemm_nt<false>(dy, 0, w_kn, 0, p, nm::S(stream));
}
// FP32-exact variant on planes (relu_mask: the h plane of the activation, bf16 [B][n_in])
int nm_tc_dense_dgrad_gen_x3(const void* dy3, size_t dy_plane, const void* w_kn3, const void* relu_mask, float mask_scale, void* dx3,
                             size_t dx_plane, int B, int n_in, int n_out, void* stream) {
    NM_REQUIRE(dy3 && w_kn3 && dx3, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0 && n_in % 8 == 0 && n_out % 8 == 0, "bad dimensions (n_in, n_out multiples of 8)");
    NM_REQUIRE(dy_plane >= (size_t)B * n_out && dx_plane >= (size_t)B * n_in && dy_plane % 8 == 0 && dx_plane % 8 == 0,
               "plane strides must cover a [B][features] plane and be multiples of 8 elements");
    GemmParams p{};
    p.M = B; p.N = n_in; p.K = n_out; p.k_chunks = (n_out + GK - 1) / GK;
    p.out = static_cast<__nv_bfloat16*>(dx3); p.ldo = n_in; p.out_plane = dx_plane;
    p.mask = static_cast<const __nv_bfloat16*>(relu_mask); p.ldm = n_in; p.mask_scale = mask_scale; p.drop_keep = 1.f;
    return launch_gemm_nt<true>(dy3, dy_plane, w_kn3, (size_t)n_in * n_out, p, nm::S(stream));
}

size_t nm_tc_dense_wgrad_gen_workspace(int B, int n_in, int n_out) { return wgrad_gen_workspace(B, n_in, n_out); }

// dw[n_in][n_out] = scale * x^T dy (+ l1 sign(w) + 2 l2 w), db[n_out] = scale * sum_b dy      (x, dy bf16 [B][.])
int nm_tc_dense_wgrad_gen_bf16(const void* x, const void* dy, const float* w, float* dw, float* db, void* workspace, size_t workspace_bytes,
                               int B, int n_in, int n_out, float scale, float l1, float l2, void* stream) {
    return dense_wgrad_gen<false>(x, 0, dy, 0, w, dw, db, workspace, workspace_bytes, B, n_in, n_out, scale, l1, l2, nm::S(stream));
}
// FP32-exact variant on planes (same workspace size)
int nm_tc_dense_wgrad_gen_x3(const void* x3, size_t x_plane, const void* dy3, size_t dy_plane, const float* w, float* dw, float* db,
                             void* workspace, size_t workspace_bytes, int B, int n_in, int n_out, float scale, float l1, float l2,
                             void* stream) {
    NM_REQUIRE(x_plane >= (size_t)B * (n_in > 0 ? n_in : 0) && dy_plane >= (size_t)B * (n_out > 0 ? n_out : 0) && x_plane % 8 == 0 && dy_plane % 8 == 0,
               "plane strides must cover a [B][features] plane and be multiples of 8 elements");
    return dense_wgrad_gen<true>(x3, x_plane, dy3, dy_plane, w, dw, db, workspace, workspace_bytes, B, n_in, n_out, scale, l1, l2, nm::S(stream));
}

}  // extern "C"
