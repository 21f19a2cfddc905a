// bf16 tensor-core Dense head for sm_100a: TMA-fed tcgen05 tiles with FP32 accumulation in TMEM.
//
//   forward  logits[B, n_out<=16] = X[B,K] . W[K,n_out]   split-K over CTAs, then one finalize kernel:
//            + bias, softmax, CCE loss, accuracy flag, fused gradient (p - onehot) * inv_batch, loss/accuracy sums
//   wgrad    dW[K, n_out] = X^T . dlogits                  split over the batch rows, fixed-order reduction
//
// Precision: X is bf16 (the pooled conv activations); the narrow operand (weights / dlogits) is carried as
// bf16 hi + bf16 lo (w = hi + lo to ~16 mantissa bits), i.e. N = 32 columns instead of 16.  The MMA cost of these
// tiles is set by the 128 B/cycle shared-memory read of the 128-row X operand (scratch/mma_rate.cu), so the 16
// extra columns are free and the head keeps the FP32-weight numerics of the CUDA-core version it replaces.
//
// Replaces: Layer_Dense.forward/backward (src/layers/Dense.py:79-80, :96-105), Activation_Softmax.forward
// (src/activations/Softmax.py:7-11), Loss_CategoricalCrossentropy.forward (src/losses/categorical_crossentropy.py:8-25)
// and Activation_Softmax_Loss_CategoricalCrossentropy.backward (:4-19) for the tensor-core mode.
#include "nm_common.cuh"
#include "nm_tc_common.cuh"
#include "nm_tc_host.cuh"
#include "nm_tc_pack.cuh"

namespace {

using namespace tc;

constexpr int kNO = 16;          // padded number of classes
constexpr int kNB = 2 * kNO;     // hi + lo columns

// ---------------------------------------------------------------------------------------------
// forward: partial[split][row][16] = sum_{k in split} X[row,k] * (Whi + Wlo)[j,k]
// ---------------------------------------------------------------------------------------------
struct DenseFwdParams {
    int B, kblocks, per_split, n_out;
    float inv_batch;
    const float* bias; const int32_t* labels;
    float* probs; float* dlogits; __nv_bfloat16* dl_hilo; float* sample_loss; int32_t* correct;
    double* block_part; unsigned* counter; double* accum;
};

constexpr int FWD_BK = 64;                               // 128-byte K rows, SWIZZLE_128B
constexpr int FWD_A_BYTES = 128 * 128, FWD_B_BYTES = kNB * 128, FWD_STAGE = FWD_A_BYTES + FWD_B_BYTES;
constexpr int FWD_STAGES = 8;

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t ra;
    float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta_rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}

// One launch for the whole head.  Grid (row tiles of 128, K splits); the `splits` CTAs of a row tile form a thread-block
// CLUSTER: each accumulates its K range on the tensor core, parks its [128 x 16] FP32 partial in its own shared memory,
// and after a cluster barrier CTA r finishes rows [r*R, (r+1)*R) of the tile -- it reads the partials of all splits over
// distributed shared memory in split order (fixed order => deterministic), adds the bias and does softmax, loss,
// accuracy flag and the gradient.  No partials in HBM, no second launch (the separate finalize kernel it replaces sat on
// a chain of L2 round trips: ~10 us for 4096 rows).  Loss / accuracy sums: per-CTA partials combined by the last CTA.
__global__ void __launch_bounds__(192, 1)
dense_head_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, DenseFwdParams p) {
    constexpr uint64_t TMPL = desc_template(16, 1024, LAYOUT_SW128);
    constexpr uint32_t IDESC = idesc_bf16(128, kNB, 0, 0);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + FWD_STAGES * FWD_STAGE);
    uint64_t* empty = full + FWD_STAGES;
    uint64_t* done = empty + FWD_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    double* red = reinterpret_cast<double*>(tmem_slot + 2);            // [4 warps][2]
    int* last_flag = reinterpret_cast<int*>(red + 8);
    float* part = reinterpret_cast<float*>(smem);                      // [128][16] partial logits, reuses stage 0 once the MMAs are done
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int splits = gridDim.y, split = blockIdx.y;                  // cluster = (1, splits, 1): the rank in the cluster is blockIdx.y
    const int kb0 = split * p.per_split, kb1 = min(p.kblocks, kb0 + p.per_split);
    pdl_trigger();

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < FWD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 32);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + (size_t)stage * FWD_STAGE;
                mbar_expect_tx(&full[stage], FWD_STAGE);
                tma_load_2d(st, &map_x, kb * FWD_BK, blockIdx.x * 128, &full[stage]);
                tma_load_2d(st + FWD_A_BYTES, &map_w, kb * FWD_BK, 0, &full[stage]);
                if (++stage == FWD_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[stage], phase);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t a_base = smem_u32(smem + (size_t)stage * FWD_STAGE), b_base = a_base + FWD_A_BYTES;
#pragma unroll
                for (int kk = 0; kk < FWD_BK / 16; ++kk)
                    mma_bf16(tmem_base, desc_at(TMPL, a_base + kk * 32), desc_at(TMPL, b_base + kk * 32), IDESC,
                             (kb > kb0 || kk > 0) ? 1u : 0u);
                mma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == FWD_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    } else {
        const int q = warp & 3;                                   // TMEM sub-partition of this warp
        mbar_wait(done, 0);                                       // every MMA has retired: the stage memory is free
        fence_after_sync();
        uint32_t raw[kNB];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        tmem_ld_x16(taddr, raw);
        tmem_ld_x16(taddr + 16, raw + 16);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(part + (q * 32 + lane) * kNO);
#pragma unroll
        for (int j = 0; j < kNO; j += 4)
            dst[j / 4] = make_float4(__uint_as_float(raw[j]) + __uint_as_float(raw[kNO + j]),
                                     __uint_as_float(raw[j + 1]) + __uint_as_float(raw[kNO + j + 1]),
                                     __uint_as_float(raw[j + 2]) + __uint_as_float(raw[kNO + j + 2]),
                                     __uint_as_float(raw[j + 3]) + __uint_as_float(raw[kNO + j + 3]));
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 32);
    cluster_sync_all();                                           // every split's partial is in its CTA's shared memory

    // ---- finish rows [split*R, split*R + R) of this tile: one thread per row (warps 2..5) ----
    const int R = (128 + splits - 1) / splits;
    const int t = (int)threadIdx.x - 64;
    float loss = 0.f, corr = 0.f;
    if (t >= 0 && t < R && split * R + t < 128) {
        const int lr = split * R + t, row = blockIdx.x * 128 + lr;
        if (row < p.B) {
            float z[kNO];
#pragma unroll
            for (int j = 0; j < kNO; ++j) z[j] = 0.f;
            const uint32_t local = smem_u32(part + lr * kNO);
            for (int s0 = 0; s0 < splits; s0 += 4) {               // four splits' loads in flight together, summed in split order
                float4 v[4][kNO / 4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int j = 0; j < kNO / 4; ++j) v[u][j] = ld_cluster_f4(local + j * 16, (uint32_t)min(s0 + u, splits - 1));
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (s0 + u < splits) {
#pragma unroll
                        for (int j = 0; j < kNO; j += 4) {
                            z[j] += v[u][j / 4].x; z[j + 1] += v[u][j / 4].y; z[j + 2] += v[u][j / 4].z; z[j + 3] += v[u][j / 4].w;
                        }
                    }
            }
            const int n_out = p.n_out;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < kNO; ++j) if (j < n_out) { z[j] += __ldg(p.bias + j); mx = fmaxf(mx, z[j]); }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < kNO; ++j) if (j < n_out) { z[j] = expf(z[j] - mx); sum += z[j]; }
            const int yv = p.labels[row];
            float best = -1.f, py = 1e-7f; int bj = 0;
            uint32_t hl[kNB / 2];                                  // bf16 pairs: [0,8) hi, [8,16) lo
#pragma unroll
            for (int j = 0; j < kNO; j += 2) {
                float d[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int jj = j + u;
                    d[u] = 0.f;
                    if (jj < n_out) {
                        const float pj = z[jj] / sum;
                        p.probs[(size_t)row * n_out + jj] = pj;
                        d[u] = (pj - (jj == yv ? 1.f : 0.f)) * p.inv_batch;
                        p.dlogits[(size_t)row * n_out + jj] = d[u];
                        if (pj > best) { best = pj; bj = jj; }
                        if (jj == yv) py = pj;
                    }
                }
                const __nv_bfloat16 h0 = __float2bfloat16(d[0]), h1 = __float2bfloat16(d[1]);
                hl[j / 2] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
                hl[kNO / 2 + j / 2] = pack_bf16x2(d[0] - __bfloat162float(h0), d[1] - __bfloat162float(h1));
            }
            if (p.dl_hilo) {
                uint4* dst = reinterpret_cast<uint4*>(p.dl_hilo + (size_t)row * kNB);
#pragma unroll
                for (int j = 0; j < kNB / 2; j += 4) dst[j / 4] = make_uint4(hl[j], hl[j + 1], hl[j + 2], hl[j + 3]);
            }
            py = fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f);
            loss = -logf(py);
            p.sample_loss[row] = loss;
            corr = (bj == yv) ? 1.f : 0.f;
            p.correct[row] = (bj == yv) ? 1 : 0;
        }
    }
    // ---- loss / accuracy sums: warp -> CTA -> (last CTA of the grid) all CTAs, every level in a fixed order ----
    if (p.accum) {
        double l = loss, c = corr;
#pragma unroll
        for (int o = 16; o; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
        if (warp >= 2 && lane == 0) { red[(warp - 2) * 2] = l; red[(warp - 2) * 2 + 1] = c; }
        __syncthreads();
        const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x, n_cta = gridDim.x * gridDim.y;
        if (threadIdx.x == 0) {
            p.block_part[2 * cta] = (red[0] + red[2]) + (red[4] + red[6]);
            p.block_part[2 * cta + 1] = (red[1] + red[3]) + (red[5] + red[7]);
            __threadfence();
            *last_flag = atomicAdd(p.counter, 1u) == n_cta - 1;
        }
        __syncthreads();
        if (*last_flag && warp == 0) {
            // the last CTA sums every CTA's pair: lane l takes CTAs l, l+32, ... in order, then a fixed shuffle tree
            // (deterministic), all loads in flight at once; the running sums are bumped with fire-and-forget reductions
            __threadfence();
            double tl = 0.0, tc_ = 0.0;
            for (unsigned b = lane; b < n_cta; b += 32) { tl += __ldcg(p.block_part + 2 * b); tc_ += __ldcg(p.block_part + 2 * b + 1); }
#pragma unroll
            for (int o = 16; o; o >>= 1) { tl += __shfl_xor_sync(0xffffffffu, tl, o); tc_ += __shfl_xor_sync(0xffffffffu, tc_, o); }
            if (lane == 0) {
                atomicAdd(p.accum, tl);                          // one writer per launch, stream-ordered: still deterministic
                atomicAdd(p.accum + 1, tc_);
                *p.counter = 0u;
            }
        }
    }
    cluster_sync_all();                                           // nobody leaves while a peer may still read its partial
}

// reference (c,h,w)-row Dense weights [K][n_out] f32 ->
//   wp   f32  [n_out][K]   K in the kernels' (h,w,c) order (CUDA-core dgrad)         (may be NULL)
//   whl  bf16 [32][K]      rows [0,16) = hi, [16,32) = lo, rows >= n_out zero          (may be NULL)
__global__ void pack_dense_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, __nv_bfloat16* __restrict__ whl,
                                          int K, int n_out, int HW, int C) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K * kNO) tcpack::dense_elem(i, w, wp, whl, K, n_out, HW, C);
}

// f32 dlogits [B][n_out] -> bf16 hi/lo [B][32] (public-API backward after a label-free forward)
__global__ void pack_dlogits_kernel(const float* __restrict__ dl, __nv_bfloat16* __restrict__ hilo, int B, int n_out) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * kNO) return;
    const int row = i / kNO, j = i % kNO;
    const float v = j < n_out ? dl[(size_t)row * n_out + j] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16(v);
    hilo[(size_t)row * kNB + j] = h;
    hilo[(size_t)row * kNB + kNO + j] = __float2bfloat16(v - __bfloat162float(h));
}

// ---------------------------------------------------------------------------------------------
// wgrad: partial[split][feature][16] = sum_{rows in split} X[row, feature] * (dl_hi + dl_lo)[row, j]
// Both operands MN-major (the reduction dimension = batch rows is the smem row index), M = 64 features per CTA.
// ---------------------------------------------------------------------------------------------
struct DenseWgradParams { int K, rblocks, per_split; float* partial; };

constexpr int WG_A_BYTES = 128 * 128, WG_B_BYTES = 128 * 64, WG_STAGE = WG_A_BYTES + WG_B_BYTES, WG_STAGES = 6;

__global__ void __launch_bounds__(192, 1)
dense_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dl, DenseWgradParams p) {
    constexpr uint64_t TMPL_A = desc_template(16, 1024, LAYOUT_SW128);     // MN-major: SBO = 8 k-rows of 128 B
    constexpr uint64_t TMPL_B = desc_template(16, 512, LAYOUT_SW64);       //           SBO = 8 k-rows of 64 B
    constexpr uint32_t IDESC = idesc_bf16(64, kNB, 1, 1);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* done = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rb0 = blockIdx.y * p.per_split, rb1 = min(p.rblocks, rb0 + p.per_split);
    pdl_trigger();

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_dl); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 32);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int rb = rb0; rb < rb1; ++rb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + (size_t)stage * WG_STAGE;
                mbar_expect_tx(&full[stage], WG_STAGE);
                tma_load_2d(st, &map_x, blockIdx.x * 64, rb * 128, &full[stage]);
                tma_load_2d(st + WG_A_BYTES, &map_dl, 0, rb * 128, &full[stage]);
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        int stage = 0; uint32_t phase = 0;
        for (int rb = rb0; rb < rb1; ++rb) {
            mbar_wait(&full[stage], phase);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t a_base = smem_u32(smem + (size_t)stage * WG_STAGE), b_base = a_base + WG_A_BYTES;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma_bf16(tmem_base, desc_at(TMPL_A, a_base + ks * 16 * 128), desc_at(TMPL_B, b_base + ks * 16 * 64), IDESC,
                             (rb > rb0 || ks > 0) ? 1u : 0u);
                mma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    } else {
        const int q = warp & 3;
        mbar_wait(done, 0);
        fence_after_sync();
        uint32_t raw[kNB];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        tmem_ld_x16(taddr, raw);
        tmem_ld_x16(taddr + 16, raw + 16);
        tmem_ld_wait();
        // M = 64: accumulator row m sits in TMEM lane (m % 16) + 32 * (m / 16)
        const int feat = blockIdx.x * 64 + 16 * q + lane;
        if (lane < 16 && feat < p.K) {
            float4* dst = reinterpret_cast<float4*>(p.partial + ((size_t)blockIdx.y * p.K + feat) * kNO);
#pragma unroll
            for (int j = 0; j < kNO; j += 4)
                dst[j / 4] = make_float4(__uint_as_float(raw[j]) + __uint_as_float(raw[kNO + j]),
                                         __uint_as_float(raw[j + 1]) + __uint_as_float(raw[kNO + j + 1]),
                                         __uint_as_float(raw[j + 2]) + __uint_as_float(raw[kNO + j + 2]),
                                         __uint_as_float(raw[j + 3]) + __uint_as_float(raw[kNO + j + 3]));
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 32);
}

// dw[perm(k)][j] = scale * sum_splits partial + l1*sign(w) + 2*l2*w ; blocks past the dW range: db[j] = scale * sum_rows dl
__global__ void __launch_bounds__(256)
dense_wgrad_finish_kernel(const float* __restrict__ partial, int n_split, const float* __restrict__ dlogits,
                          const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ db,
                          int B, int K, int n_out, int HW, int C, float scale, float l1, float l2, int dw_blocks) {
    pdl_trigger();
    pdl_wait();
    if ((int)blockIdx.x < dw_blocks) {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= K * kNO) return;
        const int k = i / kNO, j = i % kNO;
        if (j >= n_out) return;
        float s = 0.f;
        for (int sp = 0; sp < n_split; ++sp) s += __ldg(partial + ((size_t)sp * K + k) * kNO + j);
        const int hw = k / C, ch = k % C, kr = ch * HW + hw;     // kernel (h,w,c) order -> reference (c,h,w) row
        s *= scale;
        const float wv = w[(size_t)kr * n_out + j];
        if (l1 > 0.f) s += l1 * (wv > 0.f ? 1.f : (wv < 0.f ? -1.f : 0.f));
        if (l2 > 0.f) s += 2.f * l2 * wv;
        dw[(size_t)kr * n_out + j] = s;
    } else {
        __shared__ float red[256];
        const int j = blockIdx.x - dw_blocks;
        float s = 0.f;
        for (int r = threadIdx.x; r < B; r += 256) s += __ldg(dlogits + (size_t)r * n_out + j);
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) db[j] = red[0] * scale;
    }
}

int fwd_split(int B, int K, int* per_split) {
    const int kblocks = (K + FWD_BK - 1) / FWD_BK, mtiles = (B + 127) / 128;
    int want = sm_count() / mtiles;
    if (want < 1) want = 1;
    if (want > 8) want = 8;                       // the splits of a row tile form one cluster (portable limit 8)
    if (want > kblocks) want = kblocks;
    const int per = (kblocks + want - 1) / want;
    *per_split = per;
    return (kblocks + per - 1) / per;
}

int wgrad_split(int B, int K, int* per_split) {
    const int rblocks = (B + 127) / 128, mtiles = (K + 63) / 64;
    int want = sm_count() / mtiles;
    if (want < 1) want = 1;
    if (want > rblocks) want = rblocks;
    const int per = (rblocks + want - 1) / want;
    *per_split = per;
    return (rblocks + per - 1) / per;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

int nm_tc_pack_dense_weights(const float* w, float* w_packed, void* w_hilo, int K, int n_out, int HW, int C, void* stream) {
    NM_REQUIRE(w && (w_packed || w_hilo), "null pointer");
    NM_REQUIRE(K > 0 && n_out > 0 && n_out <= kNO && HW > 0 && C > 0 && HW * C == K, "bad dimensions");
    return nm::launch_pdl("pack_dense_weights", pack_dense_weights_kernel, dim3(nm_cdiv((size_t)K * kNO, 256)), dim3(256), 0, nm::S(stream),
                          w, w_packed, static_cast<__nv_bfloat16*>(w_hilo), K, n_out, HW, C);
}

int nm_tc_pack_dlogits_bf16(const float* dlogits, void* dl_hilo, int B, int n_out, void* stream) {
    NM_REQUIRE(dlogits && dl_hilo, "null pointer");
    NM_REQUIRE(B > 0 && n_out > 0 && n_out <= kNO, "bad dimensions");
    return nm::launch_pdl("pack_dlogits", pack_dlogits_kernel, dim3(nm_cdiv((size_t)B * kNO, 256)), dim3(256), 0, nm::S(stream),
                          dlogits, static_cast<__nv_bfloat16*>(dl_hilo), B, n_out);
}

// Workspace layout: [0,256) last-block ticket (persists, always left at zero) | per-CTA loss/accuracy sums.  The size is an
// upper bound that also covers every SMALLER batch (a validation pass through the same buffers).
size_t nm_tc_dense_softmax_ce_workspace(int B, int K) {
    if (B <= 0 || K <= 0) return 0;
    const size_t ctas = ((size_t)(B + 127) / 128) * 8 + (size_t)sm_count();
    return 256 + align256(ctas * 2 * sizeof(double));
}

int nm_tc_dense_softmax_ce_bf16(const void* x, const void* w_hilo, const float* bias, const int32_t* labels,
                                float* probs, float* dlogits, void* dl_hilo, float* sample_loss, int32_t* correct, double* accum,
                                void* workspace, size_t workspace_bytes, int B, int K, int n_out, float inv_batch, void* stream) {
    NM_REQUIRE(x && w_hilo && bias && labels && probs && dlogits && sample_loss && correct && workspace, "null pointer");
    NM_REQUIRE(B > 0 && K > 0 && K % 8 == 0 && n_out > 0, "bad dimensions");
    if (n_out > kNO) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_dense_softmax_ce_bf16: needs n_out <= 16");
    DenseFwdParams p{};
    p.B = B; p.kblocks = (K + FWD_BK - 1) / FWD_BK; p.n_out = n_out; p.inv_batch = inv_batch;
    const int splits = fwd_split(B, K, &p.per_split);
    const int mtiles = (B + 127) / 128;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    p.counter = reinterpret_cast<unsigned*>(ws);
    p.block_part = reinterpret_cast<double*>(ws + 256);
    NM_REQUIRE(workspace_bytes >= 256 + (size_t)mtiles * splits * 2 * sizeof(double), "workspace too small");
    p.bias = bias; p.labels = labels; p.probs = probs; p.dlogits = dlogits; p.dl_hilo = static_cast<__nv_bfloat16*>(dl_hilo);
    p.sample_loss = sample_loss; p.correct = correct; p.accum = accum;
    CUtensorMap mx, mw;
    int rc = make_map_2d(&mx, x, (uint64_t)K, (uint64_t)B, (uint64_t)K * 2, FWD_BK, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&mw, w_hilo, (uint64_t)K, kNB, (uint64_t)K * 2, FWD_BK, kNB, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const size_t smem = (size_t)FWD_STAGES * FWD_STAGE + 256;
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(dense_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    return nm::launch_pdl_cluster("dense_head_tc", dense_head_tc_kernel, dim3(mtiles, splits), dim3(192), smem, nm::S(stream),
                                  dim3(1, splits, 1), mx, mw, p);
}

size_t nm_tc_dense_wgrad_workspace(int B, int K) {
    if (B <= 0 || K <= 0) return 0;
    int per;
    const int splits = wgrad_split(B, K, &per);
    return (size_t)splits * K * kNO * sizeof(float);
}

int nm_tc_dense_wgrad_bf16(const void* x, const void* dl_hilo, const float* dlogits, const float* w, float* dw, float* db,
                           void* workspace, size_t workspace_bytes, int B, int K, int n_out, int HW, int C,
                           float scale, float l1, float l2, void* stream) {
    NM_REQUIRE(x && dl_hilo && dlogits && w && dw && db && workspace, "null pointer");
    NM_REQUIRE(B > 0 && K > 0 && K % 8 == 0 && n_out > 0 && n_out <= kNO && HW * C == K, "bad dimensions");
    NM_REQUIRE(workspace_bytes >= nm_tc_dense_wgrad_workspace(B, K), "workspace too small");
    DenseWgradParams p{};
    p.K = K; p.rblocks = (B + 127) / 128;
    const int splits = wgrad_split(B, K, &p.per_split);
    p.partial = static_cast<float*>(workspace);
    CUtensorMap mx, mdl;
    int rc = make_map_2d(&mx, x, (uint64_t)K, (uint64_t)B, (uint64_t)K * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&mdl, dl_hilo, kNB, (uint64_t)B, kNB * 2, kNB, 128, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    const size_t smem = (size_t)WG_STAGES * WG_STAGE + 256;
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(dense_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    rc = nm::launch_pdl("dense_wgrad_tc", dense_wgrad_tc_kernel, dim3((K + 63) / 64, splits), dim3(192), smem, nm::S(stream), mx, mdl, p);
    if (rc) return rc;
    const int dw_blocks = nm_cdiv((size_t)K * kNO, 256);
    return nm::launch_pdl("dense_wgrad_finish", dense_wgrad_finish_kernel, dim3(dw_blocks + n_out), dim3(256), 0, nm::S(stream),
                          (const float*)p.partial, splits, dlogits, w, dw, db, B, K, n_out, HW, C, scale, l1, l2, dw_blocks);
}

}  // extern "C"
