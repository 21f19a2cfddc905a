// bf16 tensor-core 3x3 'same' convolution + ReLU + 2x2/2 max-pool with a SHUFFLE-FREE pool epilogue (sm_100a).
//
// The pool window, not the pixel, is the GEMM row.  The zero-haloed NHWC input [N, 16, 16, C] is viewed as rows of PIXEL
// PAIRS (128 bytes = pixels (Y, 2j) and (Y, 2j+1)), eight pair-rows per image row, and one contiguous TMA box of two
// images (+ the halo rows the largest shift needs) lands in shared memory per tile.  The conv value of window (y,x) at
// window position (dy,dx) and filter tap (r,s) reads input pixel (2y+u, 2x+v), u = dy+r, v = dx+s in 0..3, i.e. pair-row
// (img*16 + 2y + u)*8 + x + v/2 at byte (v&1)*64.  For the 128 windows m = img*64 + y*8 + x of a tile that is: 8-row
// groups (one per (img, y)) at a stride of 16 pair-rows -- the UMMA descriptor's stride-byte-offset -- rows x inside a
// group, a start shift of u*8 + v/2 rows and a K offset of (v&1)*64 bytes: the shifted-descriptor implicit GEMM of
// nm_tc_conv.cu on a stride-2 lattice, no im2col and no re-layout anywhere.  Four accumulators (one per window position)
// sit side by side in TMEM, so each TMEM lane (= epilogue thread) holds all four conv values of ITS window for every
// channel and ReLU + max + arg-max nibble are pure register work: about 3x fewer epilogue instructions per pixel than
// the pixel-row version, whose two warp-shuffle exchange stages made it epilogue-bound (54 us at batch 4096).
//
// M tile = 128 windows = two images x (8 x 8) window grid (7 x 7 valid), 72 tcgen05.mma (M128 N64 K16) per tile.
//
// Replaces: conv2d_cpu (src/layers/CPU_kernels/Conv2D.cpp:83-146) + np.pad (src/layers/Conv_wrapepr.py:78-84) + embedded
// ReLU (:96-98) + Layer_MaxPooling2D(2,2).forward (maxpooling_backend.cpp:24-133) for the tensor-core mode.
#include "nm_common.cuh"
#include "nm_tc_common.cuh"
#include "nm_tc_host.cuh"

#ifdef NM_ABLATE   // timing experiments only
#define NM_TRACE(P, tile, slot) do { if ((P).trace && blockIdx.x == 0) (P).trace[((tile) / gridDim.x) * 8 + (slot)] = clock64(); } while (0)
#else
#define NM_TRACE(P, tile, slot) do { } while (0)
#endif

namespace {

using namespace tc;

constexpr int CIN = 32, COUT = 64;
constexpr int ROWB = CIN * 2;                          // 64 bytes per pixel; weights: 64-byte rows, SWIZZLE_64B
constexpr int PAIRB = 2 * ROWB;                        // 128-byte pixel-pair rows, SWIZZLE_128B
constexpr int A_ROWS = 2 * 128 + 32;                   // two images (256 pair-rows) + the largest shift (3*8 + 1 rows), rounded up
constexpr int STAGE_BYTES = A_ROWS * PAIRB;            // 36 KB
constexpr int B_TAP_BYTES = COUT * ROWB;               // 4 KB
constexpr int B_BYTES = 9 * B_TAP_BYTES;               // 36 KB
constexpr int NSTAGE = 4;
constexpr int NBUF = 2;                                // 2 x (4 positions x 64 channels) = 512 TMEM columns
constexpr int THREADS = 640;                           // warp 0 TMA, 1 MMA issuer, 2 TMEM alloc, 4-19 epilogue (4 groups)

struct PoolConvParams {
    int N, num_tiles;
    __nv_bfloat16* out;                // pooled bf16 NHWC [N, oHp, oWp, 64], written at (+ooff, +ooff)
    uint8_t* idx;                      // nibbles [N, 7, 7, 32]
    int oHp, oWp, ooff;
    long long* trace;                  // timing experiments only
};

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_pool_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, PoolConvParams p) {
    constexpr uint64_t TMPL_A = desc_template(16, 16 * PAIRB, LAYOUT_SW128);   // 8-row groups 16 pair-rows apart
    constexpr uint64_t TMPL_B = desc_template(16, 8 * ROWB, LAYOUT_SW64);
    constexpr uint32_t IDESC = idesc_bf16(128, COUT, 0, 0);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* b_smem = smem;
    uint8_t* a_smem = smem + B_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(a_smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty = full + NSTAGE;
    uint64_t* tfull = empty + NSTAGE;
    uint64_t* tempty = tfull + NBUF;
    uint64_t* bfull = tempty + NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < NBUF; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        mbar_init(bfull, 1);
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
            mbar_expect_tx(bfull, B_BYTES);
            for (int tap = 0; tap < 9; ++tap)          // weight blocks in [r][s descending] order (position stacking, see the issuer)
                tma_load_2d(b_smem + ((tap / 3) * 3 + 2 - tap % 3) * B_TAP_BYTES, &map_w, tap * CIN, 0, bfull);
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&empty[stage], phase ^ 1);
                NM_TRACE(p, tile, 0);
                mbar_expect_tx(&full[stage], STAGE_BYTES);
                // two boxes of 144 pair-rows (a TMA box dimension is at most 256); rows past the tensor read as zero
                tma_load_2d(a_smem + stage * STAGE_BYTES, &map_x, 0, tile * 256, &full[stage]);
                tma_load_2d(a_smem + stage * STAGE_BYTES + (A_ROWS / 2) * PAIRB, &map_x, 0, tile * 256 + A_ROWS / 2, &full[stage]);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ================================
        // ONE issuer: with ~50 MMAs per tile the mbarrier hand-off between tiles is amortised, and with only two
        // accumulator buffers (2 x 256 columns = all of TMEM) two alternating issuers would put the epilogue's drain time
        // on the critical path (traced: 4650 cycles per tile instead of 3456 + hand-off).  The waits for the NEXT tile's
        // barriers are done before the last MMAs of the current tile are issued, so they hide under queued MMAs.
        //
        // Position stacking: the two window positions (dy,0), (dy,1) sit in adjacent accumulator column blocks, and with
        // the weight blocks stored as [r][s descending] the taps they need for one input offset (u, v in {1,2}) --
        // W(r,v) and W(r,v-1) -- are adjacent rows, so ONE N = 128 MMA feeds both.  Per k-step: 12 MMAs of N = 128
        // (64 cycles, at the math floor) + 12 of N = 64 (48 cycles, bound by the 4 KB A read) instead of 36 of N = 64.
        constexpr uint32_t IDESC2 = idesc_bf16(128, 2 * COUT, 0, 0);
        mbar_wait(bfull, 0);
        const uint64_t b_desc0 = desc_at(TMPL_B, smem_u32(b_smem));
        if (elect_one()) {
            int it = 0;
            if ((int)blockIdx.x < p.num_tiles) { mbar_wait(&tempty[0], 1); mbar_wait(&full[0], 0); }
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int stage = it % NSTAGE, acc = it % NBUF;
                fence_after_sync();
                const uint64_t a_desc0 = desc_at(TMPL_A, smem_u32(a_smem + stage * STAGE_BYTES));
                auto A = [&](int u, int v, int kk) { return a_desc0 + (((u * 8 + (v >> 1)) * PAIRB + (v & 1) * ROWB + kk * 32) >> 4); };
                auto B = [&](int blk, int kk) { return b_desc0 + ((blk * B_TAP_BYTES + kk * 32) >> 4); };
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const uint32_t d_pair = tmem_base + acc * (4 * COUT) + dy * (2 * COUT);
#pragma unroll
                    for (int kk = 0; kk < CIN / 16; ++kk) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            const int u = dy + r;
                            if (dy == 1 && kk == CIN / 16 - 1 && r == 2) {
                                // last group of this tile: first make sure the next tile can follow without a bubble
                                const int nit = it + 1;
                                if (tile + (int)gridDim.x < p.num_tiles) {
                                    mbar_wait(&tempty[nit % NBUF], ((nit / NBUF) & 1) ^ 1);
                                    mbar_wait(&full[nit % NSTAGE], (nit / NSTAGE) & 1);
                                }
                            }
                            mma_bf16(d_pair, A(u, 1, kk), B(r * 3 + 1, kk), IDESC2, (kk | r) != 0);          // [W(r,1); W(r,0)]
                            mma_bf16(d_pair, A(u, 2, kk), B(r * 3 + 0, kk), IDESC2, 1);                      // [W(r,2); W(r,1)]
                            mma_bf16(d_pair, A(u, 0, kk), B(r * 3 + 2, kk), IDESC, 1);                       // (dy,0) <- W(r,0)
                            mma_bf16(d_pair + COUT, A(u, 3, kk), B(r * 3 + 0, kk), IDESC, 1);                // (dy,1) <- W(r,2)
                        }
                    }
                }
                mma_commit(&empty[stage]);
                mma_commit(&tfull[acc]);
                NM_TRACE(p, tile, 3);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================== epilogue: four groups of 4 warps ======================
        // group = (tile parity = accumulator buffer, channel half): the accumulator is released only after its last
        // tcgen05.ld, so the time to drain it sits on the MMA issuers' critical path; two groups per tile halve it.
        const int g4 = (warp - 4) >> 2, grp = g4 & 1, half = g4 >> 1, q = warp & 3;   // q = TMEM sub-partition of this warp
        int it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int row = q * 32 + lane;                             // window row of the tile: image (row >> 6), (i, j) = ((row >> 3) & 7, row & 7)
            const int n = tile * 2 + (row >> 6), py = (row >> 3) & 7, px = row & 7;
            const bool valid = n < p.N && py < 7 && px < 7;
            __nv_bfloat16* yo = p.out + (((size_t)n * p.oHp + py + p.ooff) * p.oWp + px + p.ooff) * COUT;
            uint32_t* io = reinterpret_cast<uint32_t*>(p.idx + (((size_t)n * 7 + py) * 7 + px) * (COUT / 2));
            mbar_wait(&tfull[grp], (it / NBUF) & 1);
            if (lane == 0 && q == 0 && half == 0) NM_TRACE(p, tile, 4);
            fence_after_sync();
#pragma unroll
            for (int hh = 0; hh < COUT / 32; ++hh) {
                const int h = half * (COUT / 32) + hh;                 // this group's 16-channel chunks
                uint32_t raw[4][16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * (4 * COUT) + h * 16;
#pragma unroll
                for (int pos = 0; pos < 4; ++pos) tmem_ld_x16(taddr + pos * COUT, raw[pos]);
                tmem_ld_wait();
                if (hh == COUT / 32 - 1) {                             // this group's columns are drained
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[grp]);
                    if (lane == 0 && q == 0 && half == 0) NM_TRACE(p, tile, 5);
                }
                float best[16];
                uint32_t nl = 0, nh = 0;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    // max-pool of ReLU == ReLU of max-pool; first max of the reference scan (0,0),(0,1),(1,0),(1,1) with its
                    // strict '>' (maxpooling_backend.cpp:104); an all-non-positive window keeps position 0 like the reference.
                    float bv = fmaxf(__uint_as_float(raw[0][c]), 0.f); uint32_t bi = 0;
#pragma unroll
                    for (int pos = 1; pos < 4; ++pos) {
                        const float v = __uint_as_float(raw[pos][c]);
                        if (v > bv) { bv = v; bi = pos; }
                    }
                    best[c] = bv;
                    const uint32_t nb = bi | (bv > 0.f ? 4u : 0u);
                    if (c < 8) nl |= nb << (4 * c); else nh |= nb << (4 * (c - 8));
                }
                if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(yo + h * 16);
                    dst[0] = make_uint4(pack_bf16x2(best[0], best[1]), pack_bf16x2(best[2], best[3]),
                                        pack_bf16x2(best[4], best[5]), pack_bf16x2(best[6], best[7]));
                    dst[1] = make_uint4(pack_bf16x2(best[8], best[9]), pack_bf16x2(best[10], best[11]),
                                        pack_bf16x2(best[12], best[13]), pack_bf16x2(best[14], best[15]));
                    *reinterpret_cast<uint2*>(io + 2 * h) = make_uint2(nl, nh);
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" {

#ifdef NM_ABLATE
static long long* g_pool_trace = nullptr;
/* timing experiments only */
int nm_tc_debug_set_trace_pool(void* buf) { g_pool_trace = static_cast<long long*>(buf); return NM_OK; }
#endif

int nm_tc_conv3x3_fprop_pool_bf16(const void* x_pad, const void* w_packed, void* y, void* idx,
                                  int N, int H, int W, int C, int OC, int out_padded, void* stream) {
    NM_REQUIRE(x_pad && w_packed && y && idx, "null pointer");
    NM_REQUIRE(N > 0, "bad batch");
    if (!(C == CIN && OC == COUT && H == 14 && W == 14))
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv3x3_fprop_pool_bf16: only C=32, OC=64, 14x14 is instantiated");
    PoolConvParams p{};
    p.N = N; p.num_tiles = (N + 1) / 2;
    p.out = static_cast<__nv_bfloat16*>(y); p.idx = static_cast<uint8_t*>(idx);
#ifdef NM_ABLATE
    p.trace = g_pool_trace;
#endif
    p.oHp = out_padded ? 9 : 7; p.oWp = p.oHp; p.ooff = out_padded ? 1 : 0;
    // x_pad [N,16,16,32] bf16 as rows of pixel pairs: [N*128][64]
    CUtensorMap mx, mw;
    int rc = make_map_2d(&mx, x_pad, 2 * CIN, (uint64_t)N * 128, PAIRB, 2 * CIN, A_ROWS / 2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_2d(&mw, w_packed, 9 * CIN, COUT, 9 * ROWB, CIN, COUT, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    const size_t smem = (size_t)B_BYTES + (size_t)NSTAGE * STAGE_BYTES + 256;
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(conv3x3_pool_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    return nm::launch_pdl("conv3x3_pool_tc", conv3x3_pool_tc_kernel, dim3(grid), dim3(THREADS), smem, nm::S(stream), mx, mw, p);
}

}  // extern "C"
