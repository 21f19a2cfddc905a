// bf16 tensor-core 3x3 'same' convolution for sm_100a: dgrad and wgrad (fprop + ReLU + pool: nm_tc_conv_pool.cu).
//
// Implicit GEMM with NO im2col anywhere -- not in HBM, not in shared memory, not in the TMA:
// activations live in HBM as zero-haloed NHWC bf16 [N, Hp=H+2, Wp=W+2, C].  For an M-tile of 128
// consecutive pixels of that flattened grid, ONE TMA box of 128 + 2*Wp + 2 pixel rows is loaded into
// shared memory, and the nine filter taps are nine tcgen05.mma issues whose A-descriptor start address
// is shifted by (r*Wp + s) rows inside that tile (descriptor row shifts validated in
// profiles/r01_tcgen05_probe.txt).  Accumulators live in TMEM (double-buffered), the epilogue warps
// read them with tcgen05.ld while the next tile's MMAs run.
//
// Replaces: conv2d_cpu / conv2d_backward_cpu (src/layers/CPU_kernels/Conv2D.cpp:83-146, :150-240),
// the np.pad in Layer_Conv2D.forward (Conv_wrapepr.py:78-84), the embedded ReLU (:96-98) and the
// following Layer_MaxPooling2D 2x2/2 (maxpooling_backend.cpp:24-133) for the tensor-core mode.
#include "nm_common.cuh"
#include "nm_tc_common.cuh"
#include "nm_tc_host.cuh"
#include "nm_tc_pack.cuh"
#include <cstdlib>

#ifdef NM_ABLATE   // timing experiments only
#define NM_TRACE(P, tile, slot) do { if ((P).trace && blockIdx.x == 0) (P).trace[((tile) / gridDim.x) * 8 + (slot)] = clock64(); } while (0)
#else
#define NM_TRACE(P, tile, slot) do { } while (0)
#endif

namespace {

using namespace tc;

constexpr int TILE_M = 128;

struct ConvTcParams {
    int num_tiles, rows_total;         // M tiles / rows of the flattened padded grid N*Hp*Wp
    int Hp, Wp, H, W;
    int a_rows;                        // 128 + 2*Wp + 2
    uint32_t a_stage_bytes;            // a_rows * row bytes, rounded up to 1024
    int n_stages;
    void* out;                         // bf16 [rows_total][NOUT]
    long long* trace;                  // timing experiments only (NM_ABLATE builds): CTA 0 records clock64() per tile and role
};

// ---------------------------------------------------------------------------------------------
// dgrad (a correlation with the flipped, transposed filter):  D[128 pixels, NOUT] = sum_{tap} A[pixel + shift(tap), 0:CIN] * B[tap][NOUT, CIN]^T
// ---------------------------------------------------------------------------------------------
// FOLD (needs Wp == 16, plain-store epilogue): the three s-taps of a filter row are three N-chunks of ONE MMA
// (N = 3*NOUT, the B blocks of taps 3r..3r+2 are contiguous rows), all reading the A rows shifted by r*Wp only;
// the s-shift is applied to the accumulator instead: out[p] = E[p][0] + E[p+1][1] + E[p+2][2], two warp shuffles
// per value (a warp holds two complete 16-pixel grid rows and the last two pixels of a row are halo, so the shift
// never crosses a warp).  The MMA is bound by the 128 B/cycle read of its 4 KB A slice, so tripling N is free:
// 12 MMAs of 56 cycles per tile instead of 36 of 45 (profiles/r01_mma_rate.txt).
template <int CIN, int NOUT, bool FOLD>
__global__ void __launch_bounds__(384, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, ConvTcParams p) {
    constexpr int ROWB = CIN * 2;                                   // bytes per pixel row: 64 (SW64) or 128 (SW128)
    static_assert(ROWB == 64 || ROWB == 128, "CIN must be 32 or 64");
    constexpr uint32_t LAYOUT = ROWB == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    constexpr uint64_t TMPL = desc_template(16, 8 * ROWB, LAYOUT);  // K-major: SBO = 8 rows
    constexpr int KSTEPS = CIN / 16;
    constexpr int B_TAP_BYTES = NOUT * ROWB;
    constexpr int B_BYTES = 9 * B_TAP_BYTES;
    static_assert(B_BYTES % 1024 == 0, "weight block must keep the A stages 1024-aligned");
    constexpr int NACC = FOLD ? 3 * NOUT : NOUT;                    // accumulator columns per buffer
    // Four accumulator buffers: the hand-off chain commit -> epilogue wake-up -> tcgen05.ld -> arrive -> MMA wake-up is about
    // as long as one tile's MMAs, so with two buffers the tensor pipe idled between tiles (measured: MMA time added on top of
    // the load time instead of hiding under it).
    constexpr int NBUF = 4;
    constexpr uint32_t TMEM_COLS = NBUF * NACC <= 32 ? 32 : NBUF * NACC <= 64 ? 64 : NBUF * NACC <= 128 ? 128 : NBUF * NACC <= 256 ? 256 : 512;
    static_assert(NBUF * NACC <= 512, "accumulators exceed TMEM");
    constexpr uint32_t IDESC = idesc_bf16(TILE_M, NACC, 0, 0);

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* b_smem = smem;
    uint8_t* a_smem = smem + B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + (size_t)p.n_stages * p.a_stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.n_stages;
    uint64_t* tfull = empty + p.n_stages;
    uint64_t* tempty = tfull + NBUF;
    uint64_t* bfull = tempty + NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < NBUF; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_init(bfull, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            mbar_expect_tx(bfull, B_BYTES);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(b_smem + tap * B_TAP_BYTES, &map_b, tap * CIN, 0, bfull);
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&empty[stage], phase ^ 1);
                NM_TRACE(p, tile, 0);
                mbar_expect_tx(&full[stage], (uint32_t)p.a_rows * ROWB);
                tma_load_2d(a_smem + (size_t)stage * p.a_stage_bytes, &map_a, 0, tile * TILE_M, &full[stage]);
                if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ============================== MMA issuers (two warps, alternating tiles) =================
        // The whole warp runs the loop converged and ONE elected lane issues: under `if (lane == 0)` the
        // compiler wraps every UTCHMMA in an ELECT/BRA.U.ANY retry loop fed by R2UR moves (~50 cycles per MMA,
        // measured with scratch/mma_rate.cu); converged + elect.sync issues one every few cycles.
        // UTCHMMA issue is back-pressured by the tensor pipe (a per-tile trace showed issue time == execution
        // time), so one issuer leaves the pipe idle while it waits on the next tile's mbarriers (~500 cycles
        // per tile); a second warp does those waits while the first one's MMAs run.
        mbar_wait(bfull, 0);
        const int mw = warp >> 1;                                        // 0 or 1
        const uint64_t b_desc0 = desc_at(TMPL, smem_u32(b_smem));
        int it = mw;
        for (int tile = blockIdx.x + mw * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int stage = it % p.n_stages, acc = it % NBUF;
            const uint32_t phase = (it / p.n_stages) & 1, acc_phase = (it / NBUF) & 1;
            mbar_wait(&tempty[acc], acc_phase ^ 1);
            if (lane == 0) NM_TRACE(p, tile, 1);
            mbar_wait(&full[stage], phase);
            if (lane == 0) NM_TRACE(p, tile, 2);
            fence_after_sync();
            if (elect_one()) {
                // descriptor = template | (address >> 4): advancing the address is one add on the low word
                const uint64_t a_desc0 = desc_at(TMPL, smem_u32(a_smem + (size_t)stage * p.a_stage_bytes));
                const uint32_t d_addr = tmem_base + acc * NACC;
                if constexpr (FOLD) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const uint32_t shift = (uint32_t)(r * p.Wp) * ROWB;
#pragma unroll
                        for (int kk = 0; kk < KSTEPS; ++kk)
                            mma_bf16(d_addr, a_desc0 + ((shift + kk * 32) >> 4), b_desc0 + ((3 * r * B_TAP_BYTES + kk * 32) >> 4),
                                     IDESC, (r | kk) != 0);
                    }
                } else {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t shift = (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * ROWB;
#pragma unroll
                        for (int kk = 0; kk < KSTEPS; ++kk)
                            mma_bf16(d_addr, a_desc0 + ((shift + kk * 32) >> 4), b_desc0 + ((tap * B_TAP_BYTES + kk * 32) >> 4),
                                     IDESC, (tap | kk) != 0);
                    }
                }
                mma_commit(&empty[stage]);          // A stage reusable once these MMAs have read it
                mma_commit(&tfull[acc]);            // accumulator ready for the epilogue
            }
            __syncwarp();
            if (lane == 0) NM_TRACE(p, tile, 3);
        }
    } else if (warp >= 4) {
        // ============================== epilogue (two groups of 4 warps, alternating tiles) ==========
        const int q = warp & 3;                       // TMEM sub-partition = lanes [32q, 32q+32)
        const int grp = (warp - 4) >> 2;
        int acc = grp; uint32_t acc_phase = 0;        // NBUF is even: a buffer always belongs to the same group
        for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x) {
            mbar_wait(&tfull[acc], acc_phase);
            if (lane == 0 && q == 0) NM_TRACE(p, tile, 4);
            fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NACC;
            uint32_t raw[NOUT];
            if constexpr (FOLD) {
                uint32_t e1[NOUT], e2[NOUT];
#pragma unroll
                for (int c = 0; c < NOUT; c += 16) {
                    tmem_ld_x16(taddr + c, raw + c);
                    tmem_ld_x16(taddr + NOUT + c, e1 + c);
                    tmem_ld_x16(taddr + 2 * NOUT + c, e2 + c);
                }
                tmem_ld_wait();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
#pragma unroll
                for (int c = 0; c < NOUT; ++c)
                    raw[c] = __float_as_uint(__uint_as_float(raw[c]) +
                                             (__uint_as_float(__shfl_down_sync(0xffffffffu, e1[c], 1)) +
                                              __uint_as_float(__shfl_down_sync(0xffffffffu, e2[c], 2))));
            } else {
#pragma unroll
                for (int c = 0; c < NOUT; c += 16) tmem_ld_x16(taddr + c, raw + c);
                tmem_ld_wait();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);   // accumulator drained: MMA may overwrite it
            }
            acc += 2; if (acc >= NBUF) { acc -= NBUF; acc_phase ^= 1; }
            if (lane == 0 && q == 0) NM_TRACE(p, tile, 5);

            {
                // plain bf16 store of the [pixel][NOUT] row (dgrad)
                const long long row = (long long)tile * TILE_M + q * 32 + lane;
                if (row < p.rows_total) {
                    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + row * NOUT);
#pragma unroll
                    for (int c = 0; c < NOUT; c += 8) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(raw[c + 0]), __uint_as_float(raw[c + 1]));
                        o.y = pack_bf16x2(__uint_as_float(raw[c + 2]), __uint_as_float(raw[c + 3]));
                        o.z = pack_bf16x2(__uint_as_float(raw[c + 4]), __uint_as_float(raw[c + 5]));
                        o.w = pack_bf16x2(__uint_as_float(raw[c + 6]), __uint_as_float(raw[c + 7]));
                        dst[c / 8] = o;
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// wgrad:  dW[oc, tap, c] = sum_pixels dY[pixel, oc] * X[pixel + shift(tap), c]   (both operands MN-major)
// Each CTA accumulates its share of the pixel range in nine [COUT x CIN] TMEM accumulators and writes one
// FP32 partial; a second kernel sums the partials in a fixed order (deterministic) into OIHW.
// ---------------------------------------------------------------------------------------------
struct WgradTcParams {
    int num_kblocks, rows_total, Wp;
    int x_rows;                          // 128 + 2*Wp + 2
    uint32_t x_stage_bytes;              // rounded to 1024
    int n_stages;
    float* partial;                      // [gridDim.x][COUT][9*CIN]
};

template <int CIN, int COUT, bool FOLD>
__global__ void __launch_bounds__(256, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, WgradTcParams p) {
    static_assert(COUT == 64 && CIN == 32, "instantiated for the MNIST conv2 shape");
    constexpr int A_ROWB = COUT * 2, B_ROWB = CIN * 2;              // 128 B (SW128), 64 B (SW64)
    constexpr uint64_t TMPL_A = desc_template(16, 8 * A_ROWB, LAYOUT_SW128);   // MN-major: SBO = 8 k-rows
    constexpr uint64_t TMPL_B = desc_template(16, 8 * B_ROWB, LAYOUT_SW64);
    constexpr uint32_t IDESC = idesc_bf16(COUT, CIN, 1, 1);        // M = 64, N = 32, both MN-major
    // FOLD: the three s-taps of a filter row become three N-chunks of ONE MMA (N = 96): for an MN-major operand
    // the leading byte offset is the distance between 32-channel chunks, and 64 B is exactly one pixel row, so
    // chunk j reads the rows shifted by j pixels.  24 MMAs of 48 cycles per k-block instead of 72 of 24.
    constexpr uint64_t TMPL_B3 = desc_template(B_ROWB, 8 * B_ROWB, LAYOUT_SW64);
    constexpr uint32_t IDESC3 = idesc_bf16(COUT, 3 * CIN, 1, 1);
    constexpr int A_BYTES = TILE_M * A_ROWB;                        // 16 KB
    constexpr uint32_t TMEM_COLS = 512;                             // 9 * 32 = 288 columns used

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t stage_bytes = A_BYTES + p.x_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.n_stages;
    uint64_t* done = empty + p.n_stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_dy); tma_prefetch_desc(&map_x); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full[stage], A_BYTES + (uint32_t)p.x_rows * B_ROWB);
                tma_load_2d(st, &map_dy, 0, kb * TILE_M, &full[stage]);
                tma_load_2d(st + A_BYTES, &map_x, 0, kb * TILE_M - (p.Wp + 1), &full[stage]);   // OOB rows read as 0
                if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // One elected lane issues everything.  UTCHMMA issue is back-pressured by the tensor pipe, so the wait for the
        // NEXT k-block's data is done before the last MMAs of the current one are issued: it hides under queued MMAs
        // instead of leaving the pipe idle for an mbarrier round trip per k-block.
        if (elect_one()) {
            int it = 0;
            if ((int)blockIdx.x < p.num_kblocks) mbar_wait(&full[0], 0);
            for (int kb = blockIdx.x; kb < p.num_kblocks; kb += gridDim.x, ++it) {
                const int stage = it % p.n_stages;
                fence_after_sync();
                const uint64_t a_desc0 = desc_at(TMPL_A, smem_u32(smem + (size_t)stage * stage_bytes));
                const uint64_t x_desc0 = desc_at(FOLD ? TMPL_B3 : TMPL_B, smem_u32(smem + (size_t)stage * stage_bytes) + A_BYTES);
#pragma unroll
                for (int ks = 0; ks < TILE_M / 16; ++ks) {
                    if (ks == TILE_M / 16 - 1 && kb + (int)gridDim.x < p.num_kblocks) {
                        const int nit = it + 1;
                        mbar_wait(&full[nit % p.n_stages], (nit / p.n_stages) & 1);
                    }
                    const uint64_t adesc = a_desc0 + ((ks * 16 * A_ROWB) >> 4);
                    if constexpr (FOLD) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const uint32_t shift = (uint32_t)(r * p.Wp + ks * 16) * B_ROWB;
                            mma_bf16(tmem_base + r * 3 * CIN, adesc, x_desc0 + (shift >> 4), IDESC3, !(it == 0 && ks == 0));
                        }
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t shift = (uint32_t)((tap / 3) * p.Wp + (tap % 3) + ks * 16) * B_ROWB;
                            mma_bf16(tmem_base + tap * CIN, adesc, x_desc0 + (shift >> 4), IDESC, !(it == 0 && ks == 0));
                        }
                    }
                }
                mma_commit(&empty[stage]);
            }
            mma_commit(done);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const bool has_work = (int)blockIdx.x < p.num_kblocks;
        if (has_work) { mbar_wait(done, 0); fence_after_sync(); }
        // M = 64 accumulator rows sit in lanes (m%16) + 32*(m/16): lanes 0..15 of each sub-partition
        float* dst = p.partial + ((size_t)blockIdx.x * COUT + 16 * q + lane) * (9 * CIN);
        for (int c = 0; c < 9 * CIN; c += 16) {
            uint32_t raw[16];
            if (has_work) { tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c, raw); tmem_ld_wait(); }
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(dst + c + j) = has_work
                        ? make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Fused backward of the 3x3 conv: dgrad AND wgrad from ONE pass over dY.
//   dX[p, c]      = sum_{tap} dY[p + shift(tap), :] . Wd[tap][c, :]          (K-major A = dY tile, tap-folded N = 96)
//   dW[oc, tap,c] = sum_p dY[p, oc] * X[p + shift(tap) - (Wp+1), c]          (MN-major A = the SAME dY tile)
// The dY tile (128 pixels + halo, 128-byte rows, SWIZZLE_128B) is read by the dgrad MMAs as a K-major operand and by
// the wgrad MMAs as an MN-major one -- the shared-memory image is identical, only the instruction descriptor differs --
// so dY (the largest tensor of the backward pass) crosses L2 once instead of twice and one kernel prologue/tail
// disappears.  TMEM: nine wgrad accumulators (288 columns, resident for the CTA's whole pixel range) + two dgrad
// accumulators of 96 columns.  Wp == 16, C = 32, OC = 64.
// ---------------------------------------------------------------------------------------------
struct ConvBwdParams {
    int num_tiles, rows_total;
    void* dx;                            // bf16 [rows_total][32]  (un-shifted grid)
    float* partial;                      // [gridDim.x][64][288]
    // BUILD variant: dY is never in HBM.  It is rebuilt per tile in shared memory from the ReLU-masked gradient of the
    // POOLED tensor (bf16 [N][PH*PW][64]) and the pool arg-max nibbles ([N][PH*PW][32 bytes]); the arg-max routing of
    // MaxPooling2D.backward (maxpooling_backend.cpp:145-238) is one byte permute per two channels.
    const __nv_bfloat16* dpool;
    const uint8_t* idx;
    int N, H, W;                         // conv output (= dY) size H x W, pooled (H/2) x (W/2)
};

__device__ __forceinline__ uint32_t bw_sw128(uint32_t row, uint32_t byte) { const uint32_t o = row * 128 + byte; return o ^ (((o >> 7) & 7) << 4); }
__device__ __forceinline__ void bw_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int BW_C = 32, BW_OC = 64, BW_WP = 16;
constexpr int BW_ROWS = TILE_M + 2 * BW_WP + 2;                       // 162
constexpr int BW_DY_BYTES = ((BW_ROWS * 128 + 1023) / 1024) * 1024;   // 21 KB
constexpr int BW_X_BYTES = ((BW_ROWS * 64 + 1023) / 1024) * 1024;     // 11 KB
constexpr int BW_STAGE = BW_DY_BYTES + BW_X_BYTES;                    // 32 KB
constexpr int BW_W_BYTES = 9 * BW_C * 128;                            // 36 KB: dgrad weights [tap][c][oc]
constexpr int BW_STAGES = 5;
constexpr int BW_DACC = 3 * BW_C;                                     // 96 dgrad accumulator columns per buffer

template <bool BUILD>
__global__ void __launch_bounds__(BUILD ? 640 : 384, 1)
conv3x3_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                      const __grid_constant__ CUtensorMap map_w, ConvBwdParams p) {
    constexpr uint64_t TMPL_128 = desc_template(16, 1024, LAYOUT_SW128);        // dY (K-major or MN-major) and Wd
    constexpr uint64_t TMPL_X3 = desc_template(64, 512, LAYOUT_SW64);           // X, MN-major, three 32-channel chunks one pixel apart
    constexpr uint32_t IDESC_D = idesc_bf16(TILE_M, BW_DACC, 0, 0);             // dgrad: M = 128 pixels, N = 96
    // wgrad: filter rows r = 1 and r = 0 share ONE M = 128 MMA: the second 64-oc chunk of the MN-major A operand is the dY
    // tile shifted by one grid row (leading byte offset = Wp rows), which pairs it with the filter row above; its pixel
    // range is shifted by Wp too, but the tiles of all CTAs still cover every pixel exactly once (the first Wp rows of
    // the tensor are halo).  Row r = 2 is a separate M = 64 MMA.  Per 16 pixels: 56 + 48 cycles instead of 3 x 48.
    constexpr uint64_t TMPL_DY2 = desc_template(BW_WP * 128, 1024, LAYOUT_SW128);
    constexpr uint32_t IDESC_W2 = idesc_bf16(2 * BW_OC, 3 * BW_C, 1, 1);        // M = 128 = (row pair, oc), N = 96, both MN-major
    constexpr uint32_t IDESC_W1 = idesc_bf16(BW_OC, 3 * BW_C, 1, 1);            // M = 64 oc
    constexpr uint32_t WCOLS = 2 * 3 * BW_C;                                    // 192: [0,96) rows {1,0}, [96,192) row 2

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_smem = smem;
    uint8_t* st_smem = smem + BW_W_BYTES;
    uint32_t* lut = reinterpret_cast<uint32_t*>(st_smem + BW_STAGES * BW_STAGE);     // BUILD: [4 positions][256] PRMT selectors
    uint64_t* full = reinterpret_cast<uint64_t*>(lut + (BUILD ? 4 * 256 : 0));
    uint64_t* empty = full + BW_STAGES;
    uint64_t* tfull = empty + BW_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint64_t* done = wfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (warp == 0 && lane == 0) { if (!BUILD) tma_prefetch_desc(&map_dy); tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
    if constexpr (BUILD) {
        // selector for window position pos: keep a channel's 16 bits where its nibble's position bits == pos, else zero
        // (the ReLU mask is already applied to dpool)
        for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) {
            const int pos = i >> 8, b = i & 255;
            lut[i] = (((b & 3) == pos) ? 0x0010u : 0x0044u) | ((((b >> 4) & 3) == pos) ? 0x3200u : 0x4400u);
        }
    }
    if (warp == 1 && lane == 0) {
        // BUILD: a stage is full when the X tile has landed (TMA tx bytes) AND the four builder warps have written dY
        for (int s = 0; s < BW_STAGES; ++s) { mbar_init(&full[s], BUILD ? 5 : 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_init(wfull, 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            mbar_expect_tx(wfull, BW_W_BYTES);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(w_smem + tap * (BW_C * 128), &map_w, tap * BW_OC, 0, wfull);
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = st_smem + (size_t)stage * BW_STAGE;
                mbar_expect_tx(&full[stage], (BUILD ? 0 : BW_ROWS * 128) + BW_ROWS * 64);
                if (!BUILD) tma_load_2d(st, &map_dy, 0, tile * TILE_M, &full[stage]);
                tma_load_2d(st + BW_DY_BYTES, &map_x, 0, tile * TILE_M - (BW_WP + 1), &full[stage]);   // OOB rows read as 0
                if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ================================
        mbar_wait(wfull, 0);
        if (elect_one()) {
            const uint64_t w_desc0 = desc_at(TMPL_128, smem_u32(w_smem));
            int it = 0;
            if ((int)blockIdx.x < p.num_tiles) mbar_wait(&full[0], 0);          // tempty[0]: fresh barrier, nothing to wait for
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % BW_STAGES, acc = it & 1;
                fence_after_sync();
                const uint32_t st = smem_u32(st_smem + (size_t)stage * BW_STAGE);
                const uint64_t dy_desc0 = desc_at(TMPL_128, st);
                const uint64_t dy_desc2 = desc_at(TMPL_DY2, st);
                const uint64_t x_desc0 = desc_at(TMPL_X3, st + BW_DY_BYTES);
                // ---- dgrad: three filter rows x four k-steps, N = 96 (the three s-taps), accumulator buffer `acc`
                const uint32_t d_addr = tmem_base + WCOLS + acc * BW_DACC;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int kk = 0; kk < BW_OC / 16; ++kk)
                        mma_bf16(d_addr, dy_desc0 + ((r * BW_WP * 128 + kk * 32) >> 4), w_desc0 + ((3 * r * BW_C * 128 + kk * 32) >> 4),
                                 IDESC_D, (r | kk) != 0);
                mma_commit(&tfull[acc]);
                // ---- wgrad: eight k-steps of 16 pixels x three filter rows, N = 96, accumulators resident
#pragma unroll
                for (int ks = 0; ks < TILE_M / 16; ++ks) {
                    if (ks == TILE_M / 16 - 1 && tile + (int)gridDim.x < p.num_tiles) {
                        // next tile: its dgrad accumulator must be drained and its data must have landed; done here the
                        // waits hide under the MMAs still queued
                        const int nit = it + 1;
                        mbar_wait(&tempty[nit & 1], ((nit >> 1) & 1) ^ 1);
                        mbar_wait(&full[nit % BW_STAGES], (nit / BW_STAGES) & 1);
                    }
                    mma_bf16(tmem_base, dy_desc2 + ((ks * 16 * 128) >> 4), x_desc0 + (((1 * BW_WP + ks * 16) * 64) >> 4), IDESC_W2,
                             !(it == 0 && ks == 0));
                    mma_bf16(tmem_base + 3 * BW_C, dy_desc0 + ((ks * 16 * 128) >> 4), x_desc0 + (((2 * BW_WP + ks * 16) * 64) >> 4), IDESC_W1,
                             !(it == 0 && ks == 0));
                }
                mma_commit(&empty[stage]);
            }
            mma_commit(done);
        }
        __syncwarp();
    } else if (BUILD && warp >= 12) {
        // ============================== dY builders: two groups of 4 warps, alternating tiles =========
        // item = (grid pixel row of the tile, half of the 64 channels); 162 rows incl. the halo the filter taps reach.
        const int bg = (warp - 12) >> 2, t = threadIdx.x & 127;
        const int PH = p.H / 2, PW = p.W / 2, Hp = p.H + 2;
        int it = bg;
        for (int tile = blockIdx.x + bg * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int stage = it % BW_STAGES;
            uint4 d[3][4], nb[3];
            int posv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {                   // issue every load of the tile first: one memory round trip
                const int item = t + 128 * k, row = item >> 1, half = item & 1;
                posv[k] = -1;
                if (row < BW_ROWS) {
                    const unsigned R = (unsigned)tile * TILE_M + row;                 // < 2^31: rows_total is an int
                    const int n = (int)(R / (unsigned)(Hp * BW_WP)), rem = (int)(R - (unsigned)n * (Hp * BW_WP)), Y = rem / BW_WP, X = rem % BW_WP;
                    if (n < p.N && Y >= 1 && Y <= p.H && X >= 1 && X <= p.W && !NM_DBG(8)) {
                        const int py = (Y - 1) >> 1, px = (X - 1) >> 1;
                        posv[k] = ((Y - 1) & 1) * 2 + ((X - 1) & 1);
                        const size_t pix = ((size_t)n * PH + py) * PW + px;
                        const uint4* src = reinterpret_cast<const uint4*>(p.dpool + pix * BW_OC + half * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) d[k][j] = __ldg(src + j);
                        nb[k] = __ldg(reinterpret_cast<const uint4*>(p.idx + pix * (BW_OC / 2) + half * 16));
                    }
                }
            }
            mbar_wait(&empty[stage], ((it / BW_STAGES) & 1) ^ 1);
            uint8_t* dyt = st_smem + (size_t)stage * BW_STAGE;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int item = t + 128 * k, row = item >> 1, half = item & 1;
                if (row >= BW_ROWS) continue;
                uint32_t out[16];
                if (posv[k] >= 0) {
                    const uint32_t* sel = lut + posv[k] * 256;
                    const uint32_t dw[16] = {d[k][0].x, d[k][0].y, d[k][0].z, d[k][0].w, d[k][1].x, d[k][1].y, d[k][1].z, d[k][1].w,
                                             d[k][2].x, d[k][2].y, d[k][2].z, d[k][2].w, d[k][3].x, d[k][3].y, d[k][3].z, d[k][3].w};
                    const uint32_t nw[4] = {nb[k].x, nb[k].y, nb[k].z, nb[k].w};
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        out[j] = NM_DBG(9) ? dw[j] ^ nw[j >> 2] : __byte_perm(dw[j], 0u, sel[(nw[j >> 2] >> (8 * (j & 3))) & 0xFFu]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) out[j] = 0u;    // halo pixel (or past the last image)
                }
#pragma unroll
                for (int j = 0; j < (NM_DBG(10) ? 1 : 4); ++j)
                    *reinterpret_cast<uint4*>(dyt + bw_sw128(row, half * 64 + j * 16)) = make_uint4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
            }
            bw_fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    } else if (warp >= 4) {
        // ============================== dgrad epilogue: two groups of 4 warps, alternating tiles ======
        const int q = warp & 3, grp = (warp - 4) >> 2;
        int it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
            mbar_wait(&tfull[grp], (it >> 1) & 1);
            fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + WCOLS + grp * BW_DACC;
            const long long row = (long long)tile * TILE_M + q * 32 + lane;
            uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.dx) + row * BW_C);
            // out[p] = E[p][s=0] + E[p+1][s=1] + E[p+2][s=2]: a warp holds two complete 16-pixel grid rows whose last two
            // pixels are halo, so the shift never crosses the warp.  Two halves of 16 channels keep the register count low.
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t e0[16], e1[16], e2[16];
                tmem_ld_x16(taddr + h * 16, e0);
                tmem_ld_x16(taddr + BW_C + h * 16, e1);
                tmem_ld_x16(taddr + 2 * BW_C + h * 16, e2);
                tmem_ld_wait();
                if (h == 1) {
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[grp]);
                }
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    e0[c] = __float_as_uint(__uint_as_float(e0[c]) + (__uint_as_float(__shfl_down_sync(0xffffffffu, e1[c], 1)) +
                                                                     __uint_as_float(__shfl_down_sync(0xffffffffu, e2[c], 2))));
                if (row < p.rows_total) {
#pragma unroll
                    for (int c = 0; c < 16; c += 8)
                        dst[h * 2 + c / 8] = make_uint4(pack_bf16x2(__uint_as_float(e0[c + 0]), __uint_as_float(e0[c + 1])),
                                                        pack_bf16x2(__uint_as_float(e0[c + 2]), __uint_as_float(e0[c + 3])),
                                                        pack_bf16x2(__uint_as_float(e0[c + 4]), __uint_as_float(e0[c + 5])),
                                                        pack_bf16x2(__uint_as_float(e0[c + 6]), __uint_as_float(e0[c + 7])));
                }
            }
        }
        // ============================== wgrad epilogue: accumulators -> one FP32 partial [64][288] per CTA ===
        if (warp < 8) {
            const bool has_work = (int)blockIdx.x < p.num_tiles;
            if (has_work) { mbar_wait(done, 0); fence_after_sync(); }
            float* part = p.partial + (size_t)blockIdx.x * BW_OC * (9 * BW_C);
            // M = 128 accumulator (columns [0,96)): lane L = 64*j + oc holds filter row r = 1 - j
            {
                const int L = q * 32 + lane, j = L >> 6, oc = L & 63;
                float* dst = part + (size_t)oc * (9 * BW_C) + (1 - j) * 3 * BW_C;
                for (int c = 0; c < 3 * BW_C; c += 16) {
                    uint32_t raw[16];
                    if (has_work) { tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c, raw); tmem_ld_wait(); }
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4)
                        *reinterpret_cast<float4*>(dst + c + jj) = has_work
                            ? make_float4(__uint_as_float(raw[jj]), __uint_as_float(raw[jj + 1]), __uint_as_float(raw[jj + 2]), __uint_as_float(raw[jj + 3]))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // M = 64 accumulator (columns [96,192)), filter row 2: rows sit in lanes (m%16) + 32*(m/16), lanes 0..15 of each sub-partition
            {
                float* dst = part + (size_t)(16 * q + lane) * (9 * BW_C) + 2 * 3 * BW_C;
                for (int c = 0; c < 3 * BW_C; c += 16) {
                    uint32_t raw[16];
                    if (has_work) { tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + 3 * BW_C + c, raw); tmem_ld_wait(); }
                    if (lane < 16) {
#pragma unroll
                        for (int jj = 0; jj < 16; jj += 4)
                            *reinterpret_cast<float4*>(dst + c + jj) = has_work
                                ? make_float4(__uint_as_float(raw[jj]), __uint_as_float(raw[jj + 1]), __uint_as_float(raw[jj + 2]), __uint_as_float(raw[jj + 3]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// dw[oc][c][r][s] = sum_cta partial[cta][oc][(r*3+s)*CIN + c]      (fixed order => deterministic)
// 64 consecutive outputs per block (coalesced 256-byte rows), 16 slices of the CTA range per output (<= 10 partials per
// thread, all loads in flight together: the kernel is latency-, not bandwidth-bound), combined in a fixed order.
constexpr int WR_SLICES = 16, WR_PER = 10;      // covers up to 160 CTAs
__global__ void __launch_bounds__(64 * WR_SLICES) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                      int n_cta, int COUT, int CIN) {
    __shared__ float red[WR_SLICES][64];
    pdl_trigger();
    pdl_wait();
    const int total = COUT * 9 * CIN;
    const int i = blockIdx.x * 64 + (threadIdx.x & 63), slice = threadIdx.x >> 6;
    float s = 0.f;
    if (i < total) {
        float v[WR_PER];
#pragma unroll
        for (int u = 0; u < WR_PER; ++u) {
            const int k = slice * WR_PER + u;
            v[u] = k < n_cta ? __ldg(partial + (size_t)k * total + i) : 0.f;
        }
        s = (((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]))) + (v[8] + v[9]);
    }
    red[slice][threadIdx.x & 63] = s;
    __syncthreads();
    if (slice == 0 && i < total) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < WR_SLICES; ++k) a += red[k][threadIdx.x];
        const int oc = i / (9 * CIN), rem = i - oc * 9 * CIN, tap = rem / CIN, c = rem - tap * CIN;
        dw[((size_t)oc * CIN + c) * 9 + tap] = a;
    }
}

// OIHW f32 -> bf16 [OC][tap][C] (fprop B operand) and [C][tap'][OC] with tap' = 8 - tap (dgrad B operand)
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                         __nv_bfloat16* __restrict__ wd, int OC, int C) {
    pdl_trigger();
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < OC * C * 9) tcpack::conv3x3_elem(i, w, wf, wd, OC, C);
}

template <int CIN, int NOUT, bool FOLD>
int launch_conv(const void* a, const void* b, ConvTcParams p, int N, cudaStream_t st) {
    constexpr int ROWB = CIN * 2;
    CUtensorMap ma, mb;
    const CUtensorMapSwizzle sw = ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    int rc = make_map_2d(&ma, a, CIN, (uint64_t)p.rows_total, ROWB, CIN, (uint32_t)p.a_rows, sw);
    if (rc) return rc;
    rc = make_map_2d(&mb, b, 9 * CIN, NOUT, 9 * ROWB, CIN, NOUT, sw);
    if (rc) return rc;
    p.a_stage_bytes = ((uint32_t)p.a_rows * ROWB + 1023u) & ~1023u;
    {   // as many A stages as fit: the tile period is (load latency + MMA time) / stages when few tiles are in flight
        static int forced = -1;
        if (forced < 0) { const char* e = getenv("NM_TC_CONV_STAGES"); forced = e ? atoi(e) : 0; }
        int fit = (int)((215 * 1024 - 9 * NOUT * ROWB) / p.a_stage_bytes);
        p.n_stages = forced > 0 ? forced : (fit > 12 ? 12 : fit);
    }
    size_t smem = 9 * NOUT * ROWB + (size_t)p.n_stages * p.a_stage_bytes + 512;   // + mbarriers
    auto kern = conv3x3_tc_kernel<CIN, NOUT, FOLD>;
    static bool attr_set = false;
    if (!attr_set) {
        NM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    return nm::launch_pdl("conv3x3_tc", kern, dim3(grid), dim3(384), smem, st, ma, mb, p);
}

}  // namespace

extern "C" {

#ifdef NM_ABLATE
static long long* g_conv_trace = nullptr;
/* timing experiments only: CTA 0 of the next conv3x3 launches records clock64() per tile into buf[tiles][8] */
int nm_tc_debug_set_trace(void* buf) { g_conv_trace = static_cast<long long*>(buf); return NM_OK; }
#endif

int nm_tc_pack_conv_weights_bf16(const float* w_oihw, void* w_fprop, void* w_dgrad, int OC, int C, void* stream) {
    NM_REQUIRE(w_oihw && (w_fprop || w_dgrad), "null pointer");
    NM_REQUIRE(OC > 0 && C > 0, "bad dimensions");
    int total = OC * C * 9;
    return nm::launch_pdl("pack_conv_weights", pack_conv_weights_kernel, dim3(nm_cdiv(total, 256)), dim3(256), 0, nm::S(stream),
                          w_oihw, static_cast<__nv_bfloat16*>(w_fprop), static_cast<__nv_bfloat16*>(w_dgrad), OC, C);
}

int nm_tc_conv3x3_dgrad_bf16(const void* dy_pad, const void* w_packed_dgrad, void* dx_grid,
                             int N, int H, int W, int C, int OC, void* stream) {
    NM_REQUIRE(dy_pad && w_packed_dgrad && dx_grid, "null pointer");
    NM_REQUIRE(N > 0, "bad batch");
    if (!(C == 32 && OC == 64))
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv3x3_dgrad_bf16: only C=32, OC=64 is instantiated");
    ConvTcParams p{};
    p.Hp = H + 2; p.Wp = W + 2; p.H = H; p.W = W;
    NM_REQUIRE(p.Wp <= 60, "image too wide for one TMA box");
    p.rows_total = N * p.Hp * p.Wp;
    p.num_tiles = (p.rows_total + TILE_M - 1) / TILE_M;
    p.a_rows = TILE_M + 2 * p.Wp + 2;
    p.out = dx_grid;
    static int fold = -1;
    if (fold < 0) { const char* e = getenv("NM_TC_DGRAD_FOLD"); fold = e ? atoi(e) : 1; }
#ifdef NM_ABLATE
    p.trace = g_conv_trace;
#endif
    if (fold && p.Wp == 16) return launch_conv<64, 32, true>(dy_pad, w_packed_dgrad, p, N, nm::S(stream));
    return launch_conv<64, 32, false>(dy_pad, w_packed_dgrad, p, N, nm::S(stream));
}

size_t nm_tc_conv3x3_wgrad_workspace(int C, int OC) {
    return (size_t)sm_count() * OC * 9 * C * sizeof(float);
}

int nm_tc_conv3x3_wgrad_bf16(const void* x_pad, const void* dy_pad, float* dw, void* workspace, size_t workspace_bytes,
                             int N, int H, int W, int C, int OC, void* stream) {
    NM_REQUIRE(x_pad && dy_pad && dw && workspace, "null pointer");
    NM_REQUIRE(N > 0, "bad batch");
    if (!(C == 32 && OC == 64))
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv3x3_wgrad_bf16: only C=32, OC=64 is instantiated");
    WgradTcParams p{};
    p.Wp = W + 2;
    NM_REQUIRE(p.Wp <= 60, "image too wide for one TMA box");
    p.rows_total = N * (H + 2) * p.Wp;
    p.num_kblocks = (p.rows_total + TILE_M - 1) / TILE_M;
    p.x_rows = TILE_M + 2 * p.Wp + 2;
    p.x_stage_bytes = ((uint32_t)p.x_rows * 64 + 1023u) & ~1023u;
    {
        static int forced = -1;
        if (forced < 0) { const char* e = getenv("NM_TC_WGRAD_STAGES"); forced = e ? atoi(e) : 0; }
        int fit = (int)((216 * 1024) / (TILE_M * 128 + p.x_stage_bytes));
        p.n_stages = forced > 0 ? forced : (fit > 10 ? 10 : fit);
    }
    int grid = p.num_kblocks < sm_count() ? p.num_kblocks : sm_count();
    NM_REQUIRE(workspace_bytes >= (size_t)grid * OC * 9 * C * sizeof(float), "workspace too small");
    p.partial = static_cast<float*>(workspace);
    CUtensorMap mdy, mx;
    int rc = make_map_2d(&mdy, dy_pad, 64, (uint64_t)p.rows_total, 128, 64, TILE_M, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&mx, x_pad, 32, (uint64_t)p.rows_total, 64, 32, (uint32_t)p.x_rows, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    size_t smem = (size_t)p.n_stages * (TILE_M * 128 + p.x_stage_bytes) + 256;
    static int fold = -1;
    if (fold < 0) { const char* e = getenv("NM_TC_WGRAD_FOLD"); fold = e ? atoi(e) : 1; }
    static bool attr_set = false;
    if (!attr_set) {
        NM_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<32, 64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        NM_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel<32, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    rc = fold ? nm::launch_pdl("conv3x3_wgrad_tc", conv3x3_wgrad_tc_kernel<32, 64, true>, dim3(grid), dim3(256), smem, nm::S(stream), mdy, mx, p)
              : nm::launch_pdl("conv3x3_wgrad_tc", conv3x3_wgrad_tc_kernel<32, 64, false>, dim3(grid), dim3(256), smem, nm::S(stream), mdy, mx, p);
    if (rc) return rc;
    int total = OC * 9 * C;
    NM_REQUIRE(grid <= WR_SLICES * WR_PER, "more CTAs than the reduction covers");
    return nm::launch_pdl("wgrad_reduce", wgrad_reduce_kernel, dim3(nm_cdiv(total, 64)), dim3(64 * WR_SLICES), 0, nm::S(stream),
                          (const float*)p.partial, dw, grid, OC, C);
}

size_t nm_tc_conv3x3_bwd_workspace(int C, int OC) { return (size_t)sm_count() * OC * 9 * C * sizeof(float); }

static int launch_conv_bwd(bool build, const void* x_pad, const void* dy_pad, const void* dpool, const void* idx,
                           const void* w_packed_dgrad, float* dw, void* dx_grid, void* workspace, size_t workspace_bytes,
                           int N, int H, int W, int C, int OC, void* stream) {
    NM_REQUIRE(x_pad && w_packed_dgrad && dx_grid && workspace, "null pointer");      // dw == NULL: leave the per-CTA partials in the workspace
    NM_REQUIRE(N > 0 && H > 0 && H % 2 == 0, "bad dimensions");
    if (!(C == BW_C && OC == BW_OC && W + 2 == BW_WP))
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv3x3_bwd_*: only C=32, OC=64, W=14 is instantiated");
    ConvBwdParams p{};
    p.rows_total = N * (H + 2) * BW_WP;
    p.num_tiles = (p.rows_total + TILE_M - 1) / TILE_M;
    p.dx = dx_grid;
    p.partial = static_cast<float*>(workspace);
    p.dpool = static_cast<const __nv_bfloat16*>(dpool);
    p.idx = static_cast<const uint8_t*>(idx);
    p.N = N; p.H = H; p.W = W;
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    NM_REQUIRE(workspace_bytes >= (size_t)grid * OC * 9 * C * sizeof(float), "workspace too small");
    CUtensorMap mdy, mx, mw;
    int rc = make_map_2d(&mx, x_pad, BW_C, (uint64_t)p.rows_total, 64, BW_C, BW_ROWS, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    mdy = mx;                                        // unused by the BUILD variant
    if (!build) {
        rc = make_map_2d(&mdy, dy_pad, BW_OC, (uint64_t)p.rows_total, 128, BW_OC, BW_ROWS, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    rc = make_map_2d(&mw, w_packed_dgrad, 9 * BW_OC, BW_C, 9 * 128, BW_OC, BW_C, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const size_t smem = (size_t)BW_W_BYTES + (size_t)BW_STAGES * BW_STAGE + (build ? 4 * 256 * 4 : 0) + 512;
    static bool attr_set = false;
    if (!attr_set) {
        NM_CUDA(cudaFuncSetAttribute(conv3x3_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        NM_CUDA(cudaFuncSetAttribute(conv3x3_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    rc = build ? nm::launch_pdl("conv3x3_bwd_tc", conv3x3_bwd_tc_kernel<true>, dim3(grid), dim3(640), smem, nm::S(stream), mdy, mx, mw, p)
               : nm::launch_pdl("conv3x3_bwd_tc", conv3x3_bwd_tc_kernel<false>, dim3(grid), dim3(384), smem, nm::S(stream), mdy, mx, mw, p);
    if (rc) return rc;
    if (!dw) return NM_OK;
    const int total = OC * 9 * C;
    NM_REQUIRE(grid <= WR_SLICES * WR_PER, "more CTAs than the reduction covers");
    return nm::launch_pdl("wgrad_reduce", wgrad_reduce_kernel, dim3(nm_cdiv(total, 64)), dim3(64 * WR_SLICES), 0, nm::S(stream),
                          (const float*)p.partial, dw, grid, OC, C);
}

int nm_tc_conv3x3_bwd_reduce_f32(const void* workspace, float* dw, int N, int H, int W, int C, int OC, void* stream) {
    NM_REQUIRE(workspace && dw, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && C == BW_C && OC == BW_OC && W + 2 == BW_WP, "bad dimensions");
    const int num_tiles = (N * (H + 2) * BW_WP + TILE_M - 1) / TILE_M;
    const int grid = num_tiles < sm_count() ? num_tiles : sm_count();          // the grid nm_tc_conv3x3_bwd_* launched
    NM_REQUIRE(grid <= WR_SLICES * WR_PER, "more CTAs than the reduction covers");
    return nm::launch_pdl("wgrad_reduce", wgrad_reduce_kernel, dim3(nm_cdiv(OC * 9 * C, 64)), dim3(64 * WR_SLICES), 0, nm::S(stream),
                          static_cast<const float*>(workspace), dw, grid, OC, C);
}

int nm_tc_conv3x3_bwd_bf16(const void* x_pad, const void* dy_pad, const void* w_packed_dgrad, float* dw, void* dx_grid,
                           void* workspace, size_t workspace_bytes, int N, int H, int W, int C, int OC, void* stream) {
    NM_REQUIRE(dy_pad, "null pointer");
    return launch_conv_bwd(false, x_pad, dy_pad, nullptr, nullptr, w_packed_dgrad, dw, dx_grid, workspace, workspace_bytes, N, H, W, C, OC, stream);
}

int nm_tc_conv3x3_bwd_pool_bf16(const void* x_pad, const void* dpool, const void* idx, const void* w_packed_dgrad, float* dw,
                                void* dx_grid, void* workspace, size_t workspace_bytes, int N, int H, int W, int C, int OC,
                                void* stream) {
    NM_REQUIRE(dpool && idx, "null pointer");
    return launch_conv_bwd(true, x_pad, nullptr, dpool, idx, w_packed_dgrad, dw, dx_grid, workspace, workspace_bytes, N, H, W, C, OC, stream);
}

}  // extern "C"
