// General-shape bf16 tensor-core 3x3 'same' convolution for sm_100a (channel counts that are multiples of 64):
// fprop (+ bias + ReLU), dgrad (+ ReLU' of the layer below) and wgrad (+ bias gradient), for the VGG-style stack of
// BASELINE.json config 4 (C(3->64) C(64->64) P C(64->128) C(128->128) P C(128->256) C(256->256) P).  The kernels of
// nm_tc_conv.cu / nm_tc_conv_pool.cu keep the whole filter bank resident in shared memory, which stops at 32 x 64
// channels; here the filter bank is STREAMED: an implicit GEMM with a K loop over (64-channel chunk, filter tap).
//
// Same data layout and the same no-im2col addressing as nm_tc_conv.cu: activations are zero-haloed NHWC bf16
// [N, Hp = H+2, Wp = W+2, C]; for 256 consecutive pixels of that flattened grid ONE activation tile (256 + 2*Wp + 2 pixel
// rows of one 64-channel chunk, 128-byte rows, SWIZZLE_128B) lands in shared memory per K chunk, and the nine taps are
// tcgen05.mma issues whose A-descriptor start address is shifted by (r*Wp + s) rows inside that tile.  The tile is
// loaded from row (m0 - Wp - 1) -- rows before the tensor read as zero -- so accumulator row p IS grid position m0 + p and
// the epilogue stores straight into the next layer's haloed tensor (halo positions are written as zero).
//
// Replaces: conv2d_cpu / conv2d_backward_cpu (src/layers/CPU_kernels/Conv2D.cpp:83-146, :150-240), the np.pad of
// Layer_Conv2D.forward (Conv_wrapepr.py:78-84), the embedded ReLU (Conv_wrapepr.py:96-98, Relu.py:7-14) and the
// bias gradient of Layer_Conv2D.backward (Conv_wrapepr.py:107-109), for the tensor-core mode.
#include "nm_common.cuh"
#include "nm_tc_common.cuh"
#include "nm_tc_host.cuh"
#include <cstdlib>

#ifdef NM_ABLATE   // timing experiments only: CTA 0 records clock64() per tile and role (scratch/gen_trace.py)
static long long* g_gen_trace = nullptr;
extern "C" int nm_tc_debug_set_trace_gen(void* buf) { g_gen_trace = static_cast<long long*>(buf); return 0; }
#define GEN_TRACE(P, it, slot) do { if ((P).trace && blockIdx.x == 0 && (it) < 48) (P).trace[(it) * 16 + (slot)] = clock64(); } while (0)
#else
#define GEN_TRACE(P, it, slot) do { } while (0)
#endif

namespace {

using namespace tc;

constexpr int G_M = 128;                 // rows of one MMA; a CTA tile is MSUB of them (MSUB accumulators per weight slice)
// Channels per K chunk = one swizzled shared-memory row: 64 channels (128-byte rows, SWIZZLE_128B) in general; 32 channels
// (64-byte rows, SWIZZLE_64B) for a first layer whose few input channels (RGB) are padded to 32 instead of 64 -- half the
// bytes written by the input pack and read by the first conv's fprop / wgrad.

struct ConvGenParams {
    int rows_total, Hp, Wp, H, W;
    int n_tiles, num_tiles;              // tile = (m_tile, n_tile), n fastest: CTAs that run together share the activation tile in L2
    int k_chunks, Cin, Cout;
    int a_boxes, a_box_rows;             // an activation stage = a_boxes TMA boxes of a_box_rows (multiple of 8) rows
    uint32_t a_stage_bytes;
    int a_stages, b_stages;
    __nv_bfloat16* out;                  // [rows_total][Cout], haloed grid
    const __nv_bfloat16* mask;           // optional [rows_total][Cout]: out = mask > 0 ? acc : 0   (ReLU' of the layer below)
    const float* bias;                   // optional [Cout]
    int relu;
    long long* trace;                    // ablation builds only
};

// D[MSUB*128 pixels, NT] = sum_{chunk, tap} A_chunk[pixel + shift(tap), 0:64] * B[tap][n0 : n0+NT, chunk]^T
// MSUB = 2: two M = 128 accumulators share every streamed weight slice; 2 x (2 x NT) TMEM columns double-buffer them against
// the two epilogue groups.  (Measured and dropped, see DESIGN.md experiment log: MSUB = 4 -- single-buffered accumulators
// expose the epilogue -- and a shared-memory-resident filter bank -- 8-15 % slower than streaming it from L2.)
// TG = filter taps per weight-ring slot.  Every slot costs one producer round (wait, expect, TMA) and one issuer round
// (wait, commit) of shared-memory barrier operations in a single thread, ~300 cycles (ablation: with loads, MMAs and epilogue
// switched off the kernel still took 9 x 300 cycles per tile and chunk).  A tap of an N = 64 tile is worth only 190-380
// cycles of MMAs, so those layers take a whole filter ROW per slot (TG = 3); N = 128 tiles (510 cycles per tap, 16 KB per
// slice) keep one tap per slot.
// SPLIT ("bf16x3", FP32-exact): activations, weights and outputs are three bf16 planes each (one tensor map per plane); the
// K loop runs six times over the chunks with the plane pair of the pass (nm_tc_common.cuh), all into the same accumulators,
// and the epilogue splits the FP32 result back into planes (three staged bulk stores per 32-column chunk).
template <bool SPLIT> struct MapSet { CUtensorMap m[SPLIT ? 3 : 1]; };

template <int NT, int MSUB, int G_KC, int TG, bool SPLIT>
__global__ void __launch_bounds__(384, 1)
conv3x3_gen_kernel(const __grid_constant__ MapSet<SPLIT> maps_a, const __grid_constant__ MapSet<SPLIT> maps_b,
                   const __grid_constant__ MapSet<SPLIT> maps_out, ConvGenParams p) {
    constexpr int PASSES = SPLIT ? 6 : 1, NPL = SPLIT ? 3 : 1;
    constexpr int O_WARP = NPL * 2048;                                        // epilogue staging per warp: 2 KB per plane
    constexpr int G_ROWB = G_KC * 2;
    constexpr uint64_t TMPL = desc_template(16, 8 * G_ROWB, G_KC == 64 ? LAYOUT_SW128 : LAYOUT_SW64);     // K-major, SBO = 8 rows
    constexpr uint32_t IDESC = idesc_bf16(G_M, NT, 0, 0);
    constexpr int TILE_ROWS = G_M * MSUB;
    constexpr int B_SLICE = NT * G_ROWB;                                      // one (tap, chunk) weight slice
    constexpr int B_BYTES = TG * B_SLICE;                                     // one ring slot
    constexpr int GROUPS = 9 / TG;
    static_assert(TG == 1 || TG == 3, "a slot holds one tap or one filter row");
    constexpr int NACC = MSUB * NT;                                           // accumulator columns per buffer
    constexpr int NBUF = 2 * NACC <= 512 ? 2 : 1;
    constexpr uint32_t TMEM_COLS = NBUF * NACC <= 256 ? 256 : 512;
    static_assert(NBUF * NACC <= 512 && MSUB % 2 == 0, "accumulators exceed TMEM");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + (size_t)p.a_stages * p.a_stage_bytes;
    // epilogue staging: 32 rows x 32 channels (2 KB, SWIZZLE_64B) per epilogue warp, drained by TMA stores
    uint8_t* o_smem = b_smem + (size_t)p.b_stages * B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(o_smem + 8 * O_WARP);
    uint64_t* afull = bars;
    uint64_t* aempty = afull + p.a_stages;
    uint64_t* bfull = aempty + p.a_stages;
    uint64_t* bempty = bfull + p.b_stages;
    uint64_t* tfull = bempty + p.b_stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) { tma_prefetch_desc(&maps_a.m[pl]); tma_prefetch_desc(&maps_b.m[pl]); tma_prefetch_desc(&maps_out.m[pl]); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < p.b_stages; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }      // one group of 4 warps per buffer
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
            int a_stage = 0, b_stage = 0; uint32_t a_phase = 0, b_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * NT;
                const int row0 = m_tile * TILE_ROWS - (p.Wp + 1);                      // negative / past-the-end rows read as zero
                for (int kcv = 0; kcv < PASSES * p.k_chunks; ++kcv) {
                    int pass = 0, kc = kcv;
                    if constexpr (SPLIT) { pass = kcv / p.k_chunks; kc = kcv - pass * p.k_chunks; }
                    const CUtensorMap* map_a = &maps_a.m[SPLIT ? split_plane_a(pass) : 0];
                    const CUtensorMap* map_b = &maps_b.m[SPLIT ? split_plane_b(pass) : 0];
                    mbar_wait(&aempty[a_stage], a_phase ^ 1);
                    if (NM_DBG(19)) mbar_arrive(&afull[a_stage]);
                    else {
                    mbar_expect_tx(&afull[a_stage], (uint32_t)(p.a_boxes * p.a_box_rows) * G_ROWB);
                    for (int j = 0; j < p.a_boxes; ++j)
                        tma_load_2d(a_smem + (size_t)a_stage * p.a_stage_bytes + (size_t)j * p.a_box_rows * G_ROWB, map_a,
                                    kc * G_KC, row0 + j * p.a_box_rows, &afull[a_stage]);
                    }
                    if (++a_stage == p.a_stages) { a_stage = 0; a_phase ^= 1; }
                    if (kcv == 0) GEN_TRACE(p, tile / (int)gridDim.x, 0);
                    for (int grp = 0; grp < GROUPS; ++grp) {
                        mbar_wait(&bempty[b_stage], b_phase ^ 1);
                        if (NM_DBG(18)) mbar_arrive(&bfull[b_stage]);
                        else {
                        mbar_expect_tx(&bfull[b_stage], B_BYTES);
#pragma unroll
                        for (int t = 0; t < TG; ++t)
                            tma_load_2d(b_smem + (size_t)b_stage * B_BYTES + t * B_SLICE, map_b, (grp * TG + t) * p.Cin + kc * G_KC, n0,
                                        &bfull[b_stage]);
                        }
                        if (++b_stage == p.b_stages) { b_stage = 0; b_phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer (one elected lane runs the whole loop) ==============================
        // The issuing thread is INSTRUCTION-bound, not tensor-bound, unless its loop is lean: ncu (profiles/r02k_*) showed
        // ~130 SASS instructions per (chunk, tap) iteration around 4-8 UTCHMMAs (64-bit descriptor arithmetic, tap / 3 and
        // tap % 3, parameters re-read from constant memory, R2UR moves) -- ~550 cycles per iteration against 130-510 cycles
        // of MMA work, i.e. every layer ran at 0.45-0.65 of its MMA schedule and the epilogue warps sat in mbarrier waits.
        // Now: the nine taps are unrolled, the per-tap row shifts live in registers, descriptors are {constant high word,
        // 32-bit low word} pairs whose low word takes immediate adds, and the accumulate flag is a compile-time constant
        // for every MMA but the first of a chunk.
        // The wait for the NEXT weight slice is issued before the last two MMAs of the current one (UTCHMMA issue is
        // back-pressured, so it hides under the queued MMAs; r01f: tensor pipe 58 % -> 74 % active).
        if (elect_one()) {
            constexpr uint32_t DESC_HI = (uint32_t)(TMPL >> 32);
            constexpr uint32_t DESC_LO = (uint32_t)(TMPL & 0xFFFFFFFFull);
            auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)DESC_HI << 32) | (uint64_t)(DESC_LO | (lo & 0x3FFFu)); };
            const int my_tiles = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            const int k_chunks = PASSES * p.k_chunks, a_stages = p.a_stages, b_stages = p.b_stages;   // SPLIT: six passes, one accumulator
            const uint32_t a_stage16 = p.a_stage_bytes >> 4, a_base16 = smem_u32(a_smem) >> 4, b_base16 = smem_u32(b_smem) >> 4;
            uint32_t shift16[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) shift16[t] = (uint32_t)(((t / 3) * p.Wp + (t % 3)) * G_ROWB) >> 4;
            int remaining_b = my_tiles * k_chunks * GROUPS;                  // weight slots this CTA still has to consume
            int a_stage = 0, b_stage = 0; uint32_t a_phase = 0, b_phase = 0;
            bool b_ready = false;                                            // bfull of the current slot already waited for
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = NBUF == 2 ? (it & 1) : 0;
                const uint32_t use = NBUF == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;   // how often this buffer has been used
                GEN_TRACE(p, it, 2);
                mbar_wait(&tempty[acc], (use & 1) ^ 1);
                GEN_TRACE(p, it, 3);
                fence_after_sync();
                const uint32_t d_addr = tmem_base + acc * NACC;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait(&afull[a_stage], a_phase);
                    if (kc == 0) GEN_TRACE(p, it, 4);
                    const uint32_t a16 = a_base16 + (uint32_t)a_stage * a_stage16;
#pragma unroll
                    for (int grp = 0; grp < GROUPS; ++grp) {
                        if (!b_ready) mbar_wait(&bfull[b_stage], b_phase);
                        b_ready = false;
                        fence_after_sync();
                        const uint32_t b16 = b_base16 + (uint32_t)b_stage * (B_BYTES >> 4);
                        --remaining_b;
#pragma unroll
                        for (int t = 0; t < TG; ++t) {
                            const int tap = grp * TG + t;
                            const uint32_t at16 = a16 + shift16[tap];
#pragma unroll
                            for (int kk = 0; kk < G_KC / 16; ++kk)
#pragma unroll
                                for (int ms = 0; ms < MSUB; ++ms) {
                                    if (t == TG - 1 && kk * MSUB + ms == (G_KC / 16) * MSUB - 2 && remaining_b > 0) {
                                        const int nb = b_stage + 1 == b_stages ? 0 : b_stage + 1;
                                        mbar_wait(&bfull[nb], nb == 0 ? b_phase ^ 1 : b_phase);
                                        b_ready = true;
                                    }
                                    const uint32_t accum = (tap | kk) != 0 ? 1u : (uint32_t)(kc != 0);
                                    if (!NM_DBG(17)) mma_bf16(d_addr + ms * NT, mk(at16 + ((ms * G_M * G_ROWB + kk * 32) >> 4)),
                                                              mk(b16 + ((t * B_SLICE + kk * 32) >> 4)), IDESC, accum);
                                }
                        }
                        mma_commit(&bempty[b_stage]);                               // weight slot reusable
                        if (grp == GROUPS - 1) {
                            mma_commit(&aempty[a_stage]);                           // activation tile reusable
                            if (kc == k_chunks - 1) { mma_commit(&tfull[acc]); GEN_TRACE(p, it, 5); }   // accumulators ready for the epilogue
                        }
                        if (++b_stage == b_stages) { b_stage = 0; b_phase ^= 1; }
                    }
                    if (++a_stage == a_stages) { a_stage = 0; a_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ============================== epilogue (two groups of 4 warps) ==============================
        // NBUF == 2: group g drains accumulator buffer g, i.e. every second tile, all its sub-tiles.
        // NBUF == 1: both groups drain every tile, MSUB / 2 sub-tiles each.
        const int q = warp & 3;                       // TMEM sub-partition = lanes [32q, 32q+32)
        // Two groups of 4 warps: group g drains accumulator buffer g, i.e. every second tile, both 128-row sub-tiles.
        // (Four groups -- one per (buffer, sub-tile) -- were measured: no gain, the tile period of the narrow layers is set
        // by the refill latency of the weight-slice ring, see the host-side stage count.)
        static_assert(NBUF == 2 && MSUB == 2, "epilogue groups are laid out for two buffers x two sub-tiles");
        const int grp = (warp - 4) >> 2;
        const int plane = p.Hp * p.Wp;
        constexpr int IT_STEP = 2;
        const int acc = grp;
        const int ms_lo = 0, ms_hi = MSUB;
        int it = acc;
        for (int tile = blockIdx.x + it * gridDim.x; tile < p.num_tiles; tile += IT_STEP * gridDim.x, it += IT_STEP) {
            const uint32_t use = (uint32_t)(it >> 1);
            if (q == 0 && lane == 0) GEN_TRACE(p, it, 6);
            mbar_wait(&tfull[acc], use & 1);
            if (q == 0 && lane == 0) GEN_TRACE(p, it, 7);
            fence_after_sync();
            const int m_tile = tile / p.n_tiles, n0 = (tile - m_tile * p.n_tiles) * NT;
#pragma unroll 1
            for (int ms = ms_lo; ms < ms_hi; ++ms) {
                const int row = m_tile * TILE_ROWS + ms * G_M + q * 32 + lane;        // rows_total < 2^31 (checked on the host)
                const int rem = row % plane, hh = rem / p.Wp, ww = rem - hh * p.Wp;
                const bool in_range = row < p.rows_total;
                const bool interior = in_range && hh >= 1 && hh <= p.H && ww >= 1 && ww <= p.W;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NACC + ms * NT;
                // Stores go through shared memory and TMA: a thread owns a ROW of the tile, so direct 16-byte global stores
                // make every warp instruction touch 32 different 128-byte lines with half a sector each (ncu: 303 MB of
                // L1->L2 write traffic for 151 MB of output).  The warp's 32 rows x 32 channels are staged in the swizzled
                // layout and leave as one bulk store (UTMASTG) per 32-column chunk.
                uint8_t* stage = o_smem + (size_t)(warp - 4) * O_WARP;
                const int row_base = m_tile * TILE_ROWS + ms * G_M + q * 32;
                const bool use_mask = p.mask != nullptr && interior;
                const __nv_bfloat16* msk = p.mask + (size_t)row * p.Cout + n0;
                // Software pipeline over 32-column chunks: the TMEM load of chunk i+1 and the mask read of chunk i are in
                // flight while chunk i is converted and stored (measured: with one load -> wait -> store chain per chunk the
                // epilogue, not the MMAs, set the tile period).
                uint32_t v[2][32];
                tmem_ld_x16(taddr, v[0]);
                tmem_ld_x16(taddr + 16, v[0] + 16);
#pragma unroll
                for (int ci = 0; ci < NT / 32; ++ci) {
                    const int c0 = ci * 32;
                    uint4 mk[4];
                    if (use_mask) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) mk[g] = __ldg(reinterpret_cast<const uint4*>(msk + c0 + g * 8));
                    }
                    tmem_ld_wait();
                    if (ci + 1 < NT / 32) {
                        tmem_ld_x16(taddr + c0 + 32, v[(ci + 1) & 1]);
                        tmem_ld_x16(taddr + c0 + 48, v[(ci + 1) & 1] + 16);
                    }
                    if (lane == 0) tma_store_wait_read();             // the previous bulk store has finished reading the staging buffer
                    __syncwarp();
                    if (in_range && !NM_DBG(16)) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float f[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = interior ? __uint_as_float(v[ci & 1][g * 8 + j]) : 0.f;
                            if (interior) {
                                if (p.bias) {
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + g * 8));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + g * 8 + 4));
                                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                                    f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                                }
                                if (p.relu) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                                }
                                if (use_mask) {
                                    const uint32_t mw[4] = {mk[g].x, mk[g].y, mk[g].z, mk[g].w};
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const uint32_t h = (mw[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;   // bf16 > 0: sign clear, not zero
                                        if ((h & 0x8000u) || !(h & 0x7FFFu)) f[j] = 0.f;
                                    }
                                }
                            }
                            uint8_t* dst = stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);                   // SWIZZLE_64B
                            if constexpr (SPLIT) {
                                float h[8], m[8], l[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) split3(f[j], h[j], m[j], l[j]);
                                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                                            pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
                                *reinterpret_cast<uint4*>(dst + 2048) = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]),
                                                                                   pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
                                *reinterpret_cast<uint4*>(dst + 4096) = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]),
                                                                                   pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
                            } else {
                                uint4 o;
                                o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                                o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                                *reinterpret_cast<uint4*>(dst) = o;
                            }
                        }
                    }
                    tma_store_fence();                                // 32 channels staged: hand them to the TMA
                    __syncwarp();
                    if (lane == 0 && row_base < p.rows_total && !NM_DBG(16)) {
#pragma unroll
                        for (int pl = 0; pl < NPL; ++pl)
                            tma_store_2d(&maps_out.m[pl], stage + pl * 2048, n0 + c0, row_base);     // rows past the tensor are clipped
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);        // accumulator buffer drained: the MMA warp may overwrite it
            if (q == 0 && lane == 0) GEN_TRACE(p, it, 8);
        }
        if (lane == 0) tma_store_wait_all();                 // every bulk store of this warp has landed
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// wgrad:  dW[oc, tap, c] = sum_pixels dY[pixel, oc] * X[pixel + shift(tap) - (Wp+1), c]      (both operands MN-major)
// A CTA owns (MOC output channels) x (32 input channels) x all nine taps -- three accumulators of N = 96, the three
// s-taps of a filter row folded into N (for an MN-major operand the leading byte offset is the distance between
// 32-channel chunks, and 64 B is exactly one pixel row of the 32-channel X slice, so chunk j reads the rows shifted
// by j pixels) -- and every `splits`-th 128-pixel block of the grid; accumulators stay in TMEM for the CTA's whole
// pixel range; one FP32 partial per CTA and a fixed-order reduction keep the result deterministic.
// ---------------------------------------------------------------------------------------------
struct WgradGenParams {
    int num_kblocks, rows_total, Wp;
    int x_rows;                          // 128 + 2*Wp + 2
    uint32_t x_stage_bytes;              // rounded to 1024
    int n_stages;
    int splits, c_chunks;                // gridDim.x = (OCp / MOC) * c_chunks, gridDim.y = splits
    int OCp, Cp;
    float* partial;                      // [splits][OCp][9][Cp]
    float* bias_partial;                 // optional [splits * c_chunks][OCp]: column sums of dY (the bias gradient)
};

// SPLIT ("bf16x3"): dY and X are three bf16 planes each; the CTA walks its pixel blocks six times, once per plane pair, into
// the same accumulators.  The dY plane of passes 0, 1, 2 is l, h, m: the bias-gradient MMAs run in exactly those passes.
template <int MOC, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
conv3x3_wgrad_gen_kernel(const __grid_constant__ MapSet<SPLIT> maps_dy, const __grid_constant__ MapSet<SPLIT> maps_x, WgradGenParams p) {
    static_assert(MOC == 64 || MOC == 128, "M = 64 or 128 output channels per CTA");
    constexpr int PASSES = SPLIT ? 6 : 1, NPL = SPLIT ? 3 : 1;
    constexpr int CB = 32;                                              // input channels per CTA: 64-byte rows, SWIZZLE_64B
    constexpr int A_BOX = G_M * 128;                                    // one [128 pixels][64 oc] box, 16 KB
    constexpr int A_BYTES = (MOC / 64) * A_BOX;
    // MN-major A: 64-oc chunks A_BOX bytes apart (leading byte offset), 8 pixel rows = 1024 B (stride byte offset)
    constexpr uint64_t TMPL_A = desc_template(MOC == 128 ? A_BOX : 16, 1024, LAYOUT_SW128);
    constexpr uint64_t TMPL_B3 = desc_template(CB * 2, 8 * CB * 2, LAYOUT_SW64);
    constexpr uint32_t IDESC3 = idesc_bf16(MOC, 3 * CB, 1, 1);
    constexpr uint32_t IDESC_ONES = idesc_bf16(MOC, CB, 1, 1);
    constexpr uint32_t TMEM_COLS = 512;                                 // 9 * 32 = 288 columns + 32 for the bias gradient
    constexpr int ONES_BYTES = G_M * CB * 2;                            // [128 pixels][32] bf16 of 1.0: the "X" of the bias gradient

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t stage_bytes = A_BYTES + p.x_stage_bytes;
    uint8_t* ones = smem + (size_t)p.n_stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + ONES_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.n_stages;
    uint64_t* done = empty + p.n_stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // item is the fastest grid index: the CTAs that are resident together walk the SAME pixel blocks (one dY / X slice
    // each), so a block of dY and X crosses HBM once and is served to the other items from L2
    const int item = blockIdx.x, split = blockIdx.y;
    const int cq = item % p.c_chunks;
    const int oc0 = (item / p.c_chunks) * MOC, c0 = cq * CB;
    // Bias gradient (Conv_wrapepr.py:107-109: sum of dvalues over batch and positions) as one more MMA against a tile of
    // ones: the CTAs that share (oc block, split) walk the same pixel blocks, CTA `cq` of them takes every c_chunks-th
    // block, so each block's column sums are formed exactly once and the extra MMA (~24 % of a block's MMA time) is spread
    // evenly.  dY is not read a second time (the separate column-sum kernels cost 166 us of the 2.08 ms VGG step).
    const bool want_bias = p.bias_partial != nullptr;
    pdl_trigger();
    if (want_bias) {
        for (int i = threadIdx.x; i < ONES_BYTES / 16; i += blockDim.x)
            reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl) { tma_prefetch_desc(&maps_dy.m[pl]); tma_prefetch_desc(&maps_x.m[pl]); }
    }
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
            for (int pass = 0; pass < PASSES; ++pass) {
                const CUtensorMap* map_dy = &maps_dy.m[SPLIT ? split_plane_a(pass) : 0];
                const CUtensorMap* map_x = &maps_x.m[SPLIT ? split_plane_b(pass) : 0];
                for (int kb = split; kb < p.num_kblocks; kb += p.splits) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full[stage], A_BYTES + (uint32_t)p.x_rows * CB * 2);
#pragma unroll
                    for (int j = 0; j < MOC / 64; ++j) tma_load_2d(st + j * A_BOX, map_dy, oc0 + j * 64, kb * G_M, &full[stage]);
                    tma_load_2d(st + A_BYTES, map_x, c0, kb * G_M - (p.Wp + 1), &full[stage]);      // OOB rows read as 0
                    if (++stage == p.n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // One elected lane issues everything.  UTCHMMA issue is back-pressured by the tensor pipe, so the wait for the NEXT
        // k-block's data is done before the last MMAs of the current one are issued: it hides under queued MMAs.
        if (elect_one()) {
            const int its_pass = split < p.num_kblocks ? (p.num_kblocks - 1 - split) / p.splits + 1 : 0;    // blocks per pass
            const int total = PASSES * its_pass;
            if (total > 0) mbar_wait(&full[0], 0);
            for (int it = 0; it < total; ++it) {
                const int stage = it % p.n_stages;
                const int pass = SPLIT ? it / its_pass : 0, bi = SPLIT ? it - pass * its_pass : it;
                fence_after_sync();
                const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t a_desc0 = desc_at(TMPL_A, st);
                const uint64_t x_desc0 = desc_at(TMPL_B3, st + A_BYTES);
                const uint64_t ones_desc0 = desc_at(TMPL_B3, smem_u32(ones));
                const bool bias_here = want_bias && (bi % p.c_chunks) == cq && pass < 3;
#pragma unroll
                for (int ks = 0; ks < G_M / 16; ++ks) {
                    if (ks == G_M / 16 - 1 && it + 1 < total) {
                        const int nit = it + 1;
                        mbar_wait(&full[nit % p.n_stages], (nit / p.n_stages) & 1);
                    }
                    const uint64_t adesc = a_desc0 + ((ks * 16 * 128) >> 4);
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const uint32_t shift = (uint32_t)(r * p.Wp + ks * 16) * (CB * 2);
                        mma_bf16(tmem_base + r * 3 * CB, adesc, x_desc0 + (shift >> 4), IDESC3, !(it == 0 && ks == 0));
                    }
                }
                if (bias_here) {            // one branch per block (the issuing thread is instruction-bound); first one: it == cq
#pragma unroll
                    for (int ks = 0; ks < G_M / 16; ++ks)
                        mma_bf16(tmem_base + 9 * CB, a_desc0 + ((ks * 16 * 128) >> 4), ones_desc0 + ((uint32_t)(ks * 16 * CB * 2) >> 4),
                                 IDESC_ONES, !(it == cq && ks == 0));
                }
                mma_commit(&empty[stage]);
            }
            mma_commit(done);
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int q = warp & 3;
        const bool has_work = split < p.num_kblocks;
        if (has_work) { mbar_wait(done, 0); fence_after_sync(); }
        // M = 128: accumulator row m is TMEM lane m.  M = 64: rows sit in lanes (m % 16) + 32 * (m / 16).
        const int m = MOC == 128 ? 32 * q + lane : 16 * q + lane;
        const bool row_ok = MOC == 128 || lane < 16;
        float* dst = p.partial + ((size_t)split * p.OCp + oc0 + m) * (size_t)(9 * p.Cp) + c0;
        for (int c = 0; c < 9 * CB; c += 16) {
            uint32_t raw[16];
            if (has_work) { tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + c, raw); tmem_ld_wait(); }
            if (row_ok) {
                const int tap = c / CB, ch = c % CB;                    // column = (r*3 + s) * 32 + channel
                float* d = dst + (size_t)tap * p.Cp + ch;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(d + j) = has_work
                        ? make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (want_bias) {
            // this CTA formed column sums iff it walked more than cq blocks; all 32 ones-columns hold the same sum
            const int n_its = split < p.num_kblocks ? (p.num_kblocks - 1 - split) / p.splits + 1 : 0;
            const bool did = n_its > cq;
            uint32_t raw[16];
            if (did) { tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + 9 * CB, raw); tmem_ld_wait(); }
            if (row_ok) p.bias_partial[((size_t)split * p.c_chunks + cq) * p.OCp + oc0 + m] = did ? __uint_as_float(raw[0]) : 0.f;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// partial [splits][OCp][9][Cp] -> OIHW [OC][C][3][3].  A block owns 64 work items x SL slices: slice j sums splits j, j + SL,
// ... in order, the slice sums are then added in order -- a fixed summation tree, so the result is bitwise repeatable.
// VEC = 4: a work item is four consecutive input channels (one 16-byte load per split; C % 4 == 0); VEC = 1: one element
// (the 3-channel first layer).  (The scalar 8-slice version took 131 us per VGG step over the six layers: 4-byte loads and,
// for the first layer, 27 blocks summing 296 splits.)
template <int VEC, int SL>
__global__ void __launch_bounds__(64 * SL)
wgrad_gen_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int OC, int C, int OCp, int Cp) {
    __shared__ float red[SL][64][VEC];
    pdl_trigger();
    pdl_wait();
    const int slice = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int cg = C / VEC;                                      // work items per (oc, tap)
    const int i = blockIdx.x * 64 + t;                           // i = (oc * 9 + tap) * cg + c / VEC : coalesced reads
    const bool ok = i < OC * 9 * cg;
    int oc = 0, tap = 0, c = 0;
    float a[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) a[v] = 0.f;
    if (ok) {
        oc = i / (9 * cg); const int rem = i - oc * 9 * cg; tap = rem / cg; c = (rem - tap * cg) * VEC;
        const size_t src = ((size_t)oc * 9 + tap) * Cp + c, plane = (size_t)OCp * 9 * Cp;
#pragma unroll 4
        for (int s = slice; s < splits; s += SL) {               // unrolled: loads in flight together
            if (VEC == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(partial + s * plane + src));
                a[0] += v.x; a[1 % VEC] += v.y; a[2 % VEC] += v.z; a[3 % VEC] += v.w;
            } else {
                a[0] += __ldg(partial + s * plane + src);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[slice][t][v] = a[v];
    __syncthreads();
    if (slice == 0 && ok) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float tsum = 0.f;
            for (int k = 0; k < SL; ++k) tsum += red[k][t][v];
            dw[((size_t)oc * C + c + v) * 9 + tap] = tsum;
        }
    }
}

// bias gradient, final stage: db[oc] = sum of the per-CTA column sums of dY that the wgrad kernel formed (fixed order).
constexpr int DB_BLOCKS = 296;
__global__ void dbias_final_kernel(const float* __restrict__ partial, float* __restrict__ db, int blocks, int OC, int OCp) {
    pdl_trigger();
    pdl_wait();
    const int oc = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;      // one warp per channel
    if (oc >= OC) return;
    float a = 0.f;
    for (int b = lane; b < blocks; b += 32) a += partial[(size_t)b * OCp + oc];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);                   // fixed tree: deterministic
    if (lane == 0) db[oc] = a;
}

// OIHW f32 [OC][C][3][3] -> bf16 [OCp][tap][Cp] (fprop B operand) and [Cp][8 - tap][OCp] (dgrad B operand), zero-padded channels
__global__ void pack_conv_weights_gen_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                             __nv_bfloat16* __restrict__ wd, int OC, int C, int OCp, int Cp) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OCp * Cp * 9) return;
    const int oc = i / (Cp * 9), rem = i - oc * Cp * 9, c = rem / 9, tap = rem - c * 9;
    const __nv_bfloat16 v = __float2bfloat16((oc < OC && c < C) ? w[((size_t)oc * C + c) * 9 + tap] : 0.f);
    if (wf) wf[((size_t)oc * 9 + tap) * Cp + c] = v;
    if (wd) wd[((size_t)c * 9 + (8 - tap)) * OCp + oc] = v;
}

// f32 NCHW [N][C][H][W] -> zero-haloed bf16 NHWC [N][H+2][W+2][Cp]; every element of the destination is written.
// One thread per grid pixel: reads are coalesced along w, and a thread writes its pixel's own Cp*2 contiguous bytes.
__device__ __forceinline__ float in_val(const float* p, float) { return __ldg(p); }
__device__ __forceinline__ float in_val(const uint8_t* p, float scale) { return (float)__ldg(p) * scale; }   // raw pixels: X / 255 done here

template <typename XT>
__global__ void pack_input_nhwc_kernel(const XT* __restrict__ x, float x_scale, __nv_bfloat16* __restrict__ xp,
                                       int N, int C, int H, int W, int Cp) {
    pdl_trigger();
    pdl_wait();
    const int groups = Cp / 8, Hp = H + 2, Wp = W + 2;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long long)N * Hp * Wp) return;
    const int ww = (int)(pix % Wp);
    const int hh = (int)((pix / Wp) % Hp);
    const int n = (int)(pix / ((long long)Wp * Hp));
    const bool interior = hh >= 1 && hh <= H && ww >= 1 && ww <= W;
    const XT* src = x + ((size_t)n * C * H + (hh - 1)) * W + (ww - 1);
    uint4* dst = reinterpret_cast<uint4*>(xp) + pix * groups;
    for (int g = 0; g < groups; ++g) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            f[j] = (interior && c < C) ? in_val(src + (size_t)c * H * W, x_scale) : 0.f;
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
        dst[g] = o;
    }
}

// the same two packs as three bf16 planes h, m, l (FP32-exact mode); weight planes are OCp * 9 * Cp elements apart
__global__ void pack_conv_weights_gen_x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                                __nv_bfloat16* __restrict__ wd, int OC, int C, int OCp, int Cp) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OCp * Cp * 9) return;
    const int oc = i / (Cp * 9), rem = i - oc * Cp * 9, c = rem / 9, tap = rem - c * 9;
    float v[3];
    split3((oc < OC && c < C) ? w[((size_t)oc * C + c) * 9 + tap] : 0.f, v[0], v[1], v[2]);
    const size_t plane = (size_t)OCp * 9 * Cp;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const __nv_bfloat16 b = __float2bfloat16(v[pl]);
        if (wf) wf[pl * plane + ((size_t)oc * 9 + tap) * Cp + c] = b;
        if (wd) wd[pl * plane + ((size_t)c * 9 + (8 - tap)) * OCp + oc] = b;
    }
}

// raw uint8 pixels are DIVIDED by denom here (the IEEE division of the FP32 path, nm_u8_to_f32_div)
__device__ __forceinline__ float in_val_div(const float* p, float) { return __ldg(p); }
__device__ __forceinline__ float in_val_div(const uint8_t* p, float denom) { return (float)__ldg(p) / denom; }

template <typename XT>
__global__ void pack_input_nhwc_x3_kernel(const XT* __restrict__ x, float denom, __nv_bfloat16* __restrict__ xp, size_t plane,
                                          int N, int C, int H, int W, int Cp) {
    pdl_trigger();
    pdl_wait();
    const int groups = Cp / 8, Hp = H + 2, Wp = W + 2;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (long long)N * Hp * Wp) return;
    const int ww = (int)(pix % Wp);
    const int hh = (int)((pix / Wp) % Hp);
    const int n = (int)(pix / ((long long)Wp * Hp));
    const bool interior = hh >= 1 && hh <= H && ww >= 1 && ww <= W;
    const XT* src = x + ((size_t)n * C * H + (hh - 1)) * W + (ww - 1);
    for (int g = 0; g < groups; ++g) {
        float h[8], m[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            split3((interior && c < C) ? in_val_div(src + (size_t)c * H * W, denom) : 0.f, h[j], m[j], l[j]);
        }
        uint4* dst = reinterpret_cast<uint4*>(xp) + pix * groups + g;
        dst[0] = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]), pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
        dst[plane / 8] = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
        dst[2 * (plane / 8)] = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
    }
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// tensor maps of the planes of a bf16 [rows][inner] matrix (`plane` elements apart)
template <bool SPLIT>
int make_mapset(MapSet<SPLIT>* m, const void* base, size_t plane, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows,
                CUtensorMapSwizzle swz) {
    for (int pl = 0; pl < (SPLIT ? 3 : 1); ++pl) {
        const int rc = make_map_2d(&m->m[pl], static_cast<const __nv_bfloat16*>(base) + (size_t)pl * plane, inner, rows, inner * 2,
                                   box_inner, box_rows, swz);
        if (rc) return rc;
    }
    return NM_OK;
}

constexpr int G_SMEM_BUDGET = 212 * 1024;

// activation stage geometry for a tile of msub * 128 pixels
void conv_gen_tile(ConvGenParams& p, int msub, int kc) {
    const int tile_rows = G_M * msub, a_rows = tile_rows + 2 * p.Wp + 2;
    p.num_tiles = ((p.rows_total + tile_rows - 1) / tile_rows) * p.n_tiles;
    p.a_boxes = (a_rows + 255) / 256;
    p.a_box_rows = round_up((a_rows + p.a_boxes - 1) / p.a_boxes, 8);
    p.a_stage_bytes = (uint32_t)round_up(p.a_boxes * p.a_box_rows * kc * 2, 1024);      // swizzled tiles start on 1024-byte boundaries
}

template <int NT, int MSUB, int G_KC, bool SPLIT>
int launch_conv_gen(const void* x_pad, size_t x_plane, const void* w_packed, size_t y_plane, ConvGenParams p, cudaStream_t st) {
    constexpr int G_ROWB = G_KC * 2;
    constexpr CUtensorMapSwizzle SWZ = G_KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    NM_REQUIRE(p.a_box_rows <= 256, "image too wide");
    MapSet<SPLIT> ma, mb, mo;
    int rc = make_mapset<SPLIT>(&ma, x_pad, x_plane, (uint64_t)p.Cin, (uint64_t)p.rows_total, G_KC, (uint32_t)p.a_box_rows, SWZ);
    if (rc) return rc;
    rc = make_mapset<SPLIT>(&mb, w_packed, (size_t)p.Cout * 9 * p.Cin, (uint64_t)9 * p.Cin, (uint64_t)p.Cout, G_KC, NT, SWZ);
    if (rc) return rc;
    rc = make_mapset<SPLIT>(&mo, p.out, y_plane, (uint64_t)p.Cout, (uint64_t)p.rows_total, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    constexpr int TG = NT == 64 ? 3 : 1;                        // taps per weight-ring slot (see the kernel)
    constexpr int B_BYTES = TG * NT * G_ROWB;
    constexpr int O_BYTES = 8 * (SPLIT ? 3 : 1) * 2048;         // epilogue staging, 2 KB per epilogue warp and plane
    p.a_stages = 2;
    {
        const int budget = G_SMEM_BUDGET - O_BYTES - p.a_stages * (int)p.a_stage_bytes;
        // The weight-slice ring must cover the refill latency of a slot (MMAs complete -> commit -> producer -> TMA from L2 ->
        // mbarrier: ~4000 cycles measured) with MMA work: a slice is worth only 190-510 cycles, so the narrow layers (N tile
        // of 64, one or two K chunks) need far more than one tile's nine slices in flight -- with nine, c1 / c2 ran at the
        // ring's latency (~4500 cycles per tile) instead of their MMA schedule (1700 / 3500).
        p.b_stages = budget / B_BYTES > 24 ? 24 : budget / B_BYTES;
    }
    NM_REQUIRE(p.a_stages >= 2 && p.b_stages >= 2, "image too wide for the shared-memory budget");
    NM_REQUIRE(2 * (p.a_stages + p.b_stages) + 4 <= 62, "too many pipeline stages for the barrier block");
    const size_t smem = (size_t)p.a_stages * p.a_stage_bytes + (size_t)p.b_stages * B_BYTES + O_BYTES + 512;   // + mbarriers
    auto kern = conv3x3_gen_kernel<NT, MSUB, G_KC, TG, SPLIT>;
    static bool attr_set = false;
    if (!attr_set) {
        NM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    return nm::launch_pdl(SPLIT ? "conv3x3_gen_x3" : "conv3x3_gen", kern, dim3(grid), dim3(384), smem, st, ma, mb, mo, p);
}

struct WgradGenPlan { int moc, c_chunks, items, splits, num_kblocks, rows_total; };

WgradGenPlan wgrad_gen_plan(int N, int H, int W, int C, int OCp) {
    WgradGenPlan g;
    g.moc = OCp % 128 == 0 ? 128 : 64;
    g.c_chunks = (C + 31) / 32;                         // all-zero padded input-channel chunks are skipped
    g.items = (OCp / g.moc) * g.c_chunks;
    g.rows_total = N * (H + 2) * (W + 2);
    g.num_kblocks = (g.rows_total + G_M - 1) / G_M;
    int s = (2 * sm_count()) / g.items;                 // about two waves: the tail of the first hides under the second
    if (s * g.items > sm_count() && s * g.items < 2 * sm_count()) s = sm_count() / g.items;   // ... or exactly one
    if (s < 1) s = 1;
    if (s > g.num_kblocks) s = g.num_kblocks;
    g.splits = s;
    return g;
}

size_t wgrad_gen_workspace(int N, int H, int W, int C, int Cp, int OCp) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || Cp < C || OCp <= 0 || (Cp % 64 && Cp != 32) || OCp % 64) return 0;
    const WgradGenPlan g = wgrad_gen_plan(N, H, W, C, OCp);
    return (size_t)g.splits * OCp * 9 * Cp * sizeof(float) + (size_t)DB_BLOCKS * OCp * sizeof(float);
}

template <bool SPLIT>
int conv_gen(const void* x_pad, size_t x_plane, const void* w_packed, const float* bias, const void* relu_mask, void* y_pad, size_t y_plane,
             int N, int H, int W, int Cin, int Cout, int relu, cudaStream_t st) {
    NM_REQUIRE(x_pad && w_packed && y_pad, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0, "bad dimensions");
    NM_REQUIRE(Cin > 0 && Cout > 0 && (Cin % 64 == 0 || Cin == 32) && Cout % 64 == 0,
               "channel counts must be multiples of 64 (pad with zero channels); Cin = 32 is accepted for a padded first layer");
    ConvGenParams p{};
    p.Hp = H + 2; p.Wp = W + 2; p.H = H; p.W = W;
    NM_REQUIRE((long long)N * p.Hp * p.Wp < (1ll << 31) - 1024, "batch too large for 32-bit row coordinates");
    p.rows_total = N * p.Hp * p.Wp;
    p.Cin = Cin; p.Cout = Cout;
    const int kc = Cin == 32 ? 32 : 64;
    p.k_chunks = Cin / kc;
    const int nt = Cout % 128 == 0 ? 128 : 64;
    p.n_tiles = Cout / nt;
    p.out = static_cast<__nv_bfloat16*>(y_pad);
    p.mask = static_cast<const __nv_bfloat16*>(relu_mask);
    p.bias = bias;
    p.relu = relu;
#ifdef NM_ABLATE
    p.trace = g_gen_trace;
#endif
    conv_gen_tile(p, 2, kc);
    if (kc == 32)
        return nt == 128 ? launch_conv_gen<128, 2, 32, SPLIT>(x_pad, x_plane, w_packed, y_plane, p, st)
                         : launch_conv_gen<64, 2, 32, SPLIT>(x_pad, x_plane, w_packed, y_plane, p, st);
    return nt == 128 ? launch_conv_gen<128, 2, 64, SPLIT>(x_pad, x_plane, w_packed, y_plane, p, st)
                     : launch_conv_gen<64, 2, 64, SPLIT>(x_pad, x_plane, w_packed, y_plane, p, st);
}

template <int MOC, bool SPLIT>
int launch_wgrad_gen(const MapSet<SPLIT>& mdy, const MapSet<SPLIT>& mx, const WgradGenParams& p, int items, size_t smem, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        NM_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_gen_kernel<MOC, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_set = true;
    }
    return nm::launch_pdl(SPLIT ? "conv3x3_wgrad_gen_x3" : "conv3x3_wgrad_gen", conv3x3_wgrad_gen_kernel<MOC, SPLIT>, dim3(items, p.splits),
                          dim3(256), smem, st, mdy, mx, p);
}

template <bool SPLIT>
int conv_wgrad_gen(const void* x_pad, size_t x_plane, const void* dy_pad, size_t dy_plane, float* dw, float* dbias, void* workspace,
                   size_t workspace_bytes, int N, int H, int W, int C, int OC, int Cp, int OCp, cudaStream_t st) {
    NM_REQUIRE(x_pad && dy_pad && dw && workspace, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0, "bad dimensions");
    NM_REQUIRE(C > 0 && OC > 0 && Cp >= C && OCp >= OC && (Cp % 64 == 0 || Cp == 32) && OCp % 64 == 0,
               "padded channel counts must be multiples of 64 (input channels: or 32)");
    NM_REQUIRE((long long)N * (H + 2) * (W + 2) < (1ll << 31) - 1024, "batch too large for 32-bit row coordinates");
    const WgradGenPlan g = wgrad_gen_plan(N, H, W, C, OCp);
    NM_REQUIRE(workspace_bytes >= wgrad_gen_workspace(N, H, W, C, Cp, OCp), "workspace too small");
    WgradGenParams p{};
    p.Wp = W + 2;
    NM_REQUIRE(p.Wp <= 60, "image too wide for one TMA box");
    p.rows_total = g.rows_total;
    p.num_kblocks = g.num_kblocks;
    p.x_rows = G_M + 2 * p.Wp + 2;
    p.x_stage_bytes = ((uint32_t)p.x_rows * 64 + 1023u) & ~1023u;
    const int a_bytes = (g.moc / 64) * G_M * 128;
    {
        const int fit = (int)((212 * 1024 - G_M * 64) / (a_bytes + p.x_stage_bytes));
        p.n_stages = fit > 8 ? 8 : fit;
    }
    p.splits = g.splits;
    p.c_chunks = g.c_chunks;
    p.OCp = OCp; p.Cp = Cp;
    p.partial = static_cast<float*>(workspace);
    p.bias_partial = dbias ? p.partial + (size_t)g.splits * OCp * 9 * Cp : nullptr;
    NM_REQUIRE(g.splits * g.c_chunks <= DB_BLOCKS, "bias-gradient partials exceed the workspace");
    MapSet<SPLIT> mdy, mx;
    int rc = make_mapset<SPLIT>(&mdy, dy_pad, dy_plane, (uint64_t)OCp, (uint64_t)p.rows_total, 64, G_M, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_mapset<SPLIT>(&mx, x_pad, x_plane, (uint64_t)Cp, (uint64_t)p.rows_total, 32, (uint32_t)p.x_rows, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    const size_t smem = (size_t)p.n_stages * (a_bytes + p.x_stage_bytes) + G_M * 64 + 256;
    rc = g.moc == 128 ? launch_wgrad_gen<128, SPLIT>(mdy, mx, p, g.items, smem, st) : launch_wgrad_gen<64, SPLIT>(mdy, mx, p, g.items, smem, st);
    if (rc) return rc;
    if (C % 4 == 0)
        rc = nm::launch_pdl("wgrad_gen_reduce", wgrad_gen_reduce_kernel<4, 8>, dim3(nm_cdiv((size_t)OC * (C / 4) * 9, 64)), dim3(64 * 8), 0,
                            st, (const float*)p.partial, dw, p.splits, OC, C, OCp, Cp);
    else
        rc = nm::launch_pdl("wgrad_gen_reduce", wgrad_gen_reduce_kernel<1, 16>, dim3(nm_cdiv((size_t)OC * C * 9, 64)), dim3(64 * 16), 0,
                            st, (const float*)p.partial, dw, p.splits, OC, C, OCp, Cp);
    if (rc || !dbias) return rc;
    // column sums of dY were formed inside the wgrad kernel (one partial per CTA): fixed-order final sum
    return nm::launch_pdl("dbias_final", dbias_final_kernel, dim3(nm_cdiv((size_t)OC * 32, 256)), dim3(256), 0, st,
                          (const float*)p.bias_partial, dbias, g.splits * g.c_chunks, OC, OCp);
}

}  // namespace

extern "C" {

int nm_tc_pack_conv_weights_gen_bf16(const float* w_oihw, void* w_fprop, void* w_dgrad, int OC, int C, int OCp, int Cp, void* stream) {
    NM_REQUIRE(w_oihw && (w_fprop || w_dgrad), "null pointer");
    NM_REQUIRE(OC > 0 && C > 0 && OCp >= OC && Cp >= C && OCp % 64 == 0 && (Cp % 64 == 0 || Cp == 32),
               "padded channel counts must be multiples of 64 (input channels: or 32)");
    const int total = OCp * Cp * 9;
    return nm::launch_pdl("pack_conv_weights_gen", pack_conv_weights_gen_kernel, dim3(nm_cdiv(total, 256)), dim3(256), 0, nm::S(stream),
                          w_oihw, static_cast<__nv_bfloat16*>(w_fprop), static_cast<__nv_bfloat16*>(w_dgrad), OC, C, OCp, Cp);
}

int nm_tc_pack_input_nhwc_bf16(const float* x_nchw, void* x_pad, int N, int C, int H, int W, int Cp, void* stream) {
    NM_REQUIRE(x_nchw && x_pad, "null pointer");
    NM_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C && Cp % 8 == 0, "bad dimensions");
    const long long total = (long long)N * (H + 2) * (W + 2);
    return nm::launch_pdl("pack_input_nhwc", pack_input_nhwc_kernel<float>, dim3(nm_cdiv((size_t)total, 128)), dim3(128), 0, nm::S(stream),
                          x_nchw, 1.f, static_cast<__nv_bfloat16*>(x_pad), N, C, H, W, Cp);
}
int nm_tc_pack_input_nhwc_u8_bf16(const uint8_t* x_nchw, float scale, void* x_pad, int N, int C, int H, int W, int Cp, void* stream) {
    NM_REQUIRE(x_nchw && x_pad, "null pointer");
    NM_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C && Cp % 8 == 0, "bad dimensions");
    const long long total = (long long)N * (H + 2) * (W + 2);
    return nm::launch_pdl("pack_input_nhwc_u8", pack_input_nhwc_kernel<uint8_t>, dim3(nm_cdiv((size_t)total, 128)), dim3(128), 0, nm::S(stream),
                          x_nchw, scale, static_cast<__nv_bfloat16*>(x_pad), N, C, H, W, Cp);
}

int nm_tc_conv3x3_gen_bf16(const void* x_pad, const void* w_packed, const float* bias, const void* relu_mask, void* y_pad,
                           int N, int H, int W, int Cin, int Cout, int relu, void* stream) {
    return conv_gen<false>(x_pad, 0, w_packed, bias, relu_mask, y_pad, 0, N, H, W, Cin, Cout, relu, nm::S(stream));
}
// FP32-exact variant: x_pad3 / y_pad3 are three bf16 planes x_plane / y_plane elements apart, w_packed3 the three planes of
// nm_tc_pack_conv_weights_gen_x3; relu_mask = the h plane of the activation below
int nm_tc_conv3x3_gen_x3(const void* x_pad3, size_t x_plane, const void* w_packed3, const float* bias, const void* relu_mask, void* y_pad3,
                         size_t y_plane, int N, int H, int W, int Cin, int Cout, int relu, void* stream) {
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad dimensions");
    NM_REQUIRE(x_plane >= (size_t)N * (H + 2) * (W + 2) * Cin && y_plane >= (size_t)N * (H + 2) * (W + 2) * Cout && x_plane % 8 == 0 &&
               y_plane % 8 == 0, "plane strides must cover a haloed NHWC plane and be multiples of 8 elements");
    return conv_gen<true>(x_pad3, x_plane, w_packed3, bias, relu_mask, y_pad3, y_plane, N, H, W, Cin, Cout, relu, nm::S(stream));
}

size_t nm_tc_conv3x3_wgrad_gen_workspace(int N, int H, int W, int C, int Cp, int OCp) { return wgrad_gen_workspace(N, H, W, C, Cp, OCp); }

int nm_tc_conv3x3_wgrad_gen_bf16(const void* x_pad, const void* dy_pad, float* dw, float* dbias, void* workspace, size_t workspace_bytes,
                                 int N, int H, int W, int C, int OC, int Cp, int OCp, void* stream) {
    return conv_wgrad_gen<false>(x_pad, 0, dy_pad, 0, dw, dbias, workspace, workspace_bytes, N, H, W, C, OC, Cp, OCp, nm::S(stream));
}
// FP32-exact variant on planes (same workspace size)
int nm_tc_conv3x3_wgrad_gen_x3(const void* x_pad3, size_t x_plane, const void* dy_pad3, size_t dy_plane, float* dw, float* dbias,
                               void* workspace, size_t workspace_bytes, int N, int H, int W, int C, int OC, int Cp, int OCp, void* stream) {
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && Cp > 0 && OCp > 0, "bad dimensions");
    NM_REQUIRE(x_plane >= (size_t)N * (H + 2) * (W + 2) * Cp && dy_plane >= (size_t)N * (H + 2) * (W + 2) * OCp && x_plane % 8 == 0 &&
               dy_plane % 8 == 0, "plane strides must cover a haloed NHWC plane and be multiples of 8 elements");
    return conv_wgrad_gen<true>(x_pad3, x_plane, dy_pad3, dy_plane, dw, dbias, workspace, workspace_bytes, N, H, W, C, OC, Cp, OCp,
                                nm::S(stream));
}

int nm_tc_pack_conv_weights_gen_x3(const float* w_oihw, void* w_fprop3, void* w_dgrad3, int OC, int C, int OCp, int Cp, void* stream) {
    NM_REQUIRE(w_oihw && (w_fprop3 || w_dgrad3), "null pointer");
    NM_REQUIRE(OC > 0 && C > 0 && OCp >= OC && Cp >= C && OCp % 64 == 0 && (Cp % 64 == 0 || Cp == 32),
               "padded channel counts must be multiples of 64 (input channels: or 32)");
    const int total = OCp * Cp * 9;
    return nm::launch_pdl("pack_conv_weights_gen_x3", pack_conv_weights_gen_x3_kernel, dim3(nm_cdiv(total, 256)), dim3(256), 0, nm::S(stream),
                          w_oihw, static_cast<__nv_bfloat16*>(w_fprop3), static_cast<__nv_bfloat16*>(w_dgrad3), OC, C, OCp, Cp);
}

int nm_tc_pack_input_nhwc_x3(const float* x_nchw, void* x_pad3, size_t plane, int N, int C, int H, int W, int Cp, void* stream) {
    NM_REQUIRE(x_nchw && x_pad3, "null pointer");
    NM_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C && Cp % 8 == 0, "bad dimensions");
    const long long total = (long long)N * (H + 2) * (W + 2);
    NM_REQUIRE(plane >= (size_t)total * Cp && plane % 8 == 0, "plane stride must cover a haloed NHWC plane and be a multiple of 8 elements");
    return nm::launch_pdl("pack_input_nhwc_x3", pack_input_nhwc_x3_kernel<float>, dim3(nm_cdiv((size_t)total, 128)), dim3(128), 0, nm::S(stream),
                          x_nchw, 1.f, static_cast<__nv_bfloat16*>(x_pad3), plane, N, C, H, W, Cp);
}
int nm_tc_pack_input_nhwc_u8_x3(const uint8_t* x_nchw, float denom, void* x_pad3, size_t plane, int N, int C, int H, int W, int Cp, void* stream) {
    NM_REQUIRE(x_nchw && x_pad3, "null pointer");
    NM_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && Cp >= C && Cp % 8 == 0 && denom != 0.f, "bad dimensions");
    const long long total = (long long)N * (H + 2) * (W + 2);
    NM_REQUIRE(plane >= (size_t)total * Cp && plane % 8 == 0, "plane stride must cover a haloed NHWC plane and be a multiple of 8 elements");
    return nm::launch_pdl("pack_input_nhwc_u8_x3", pack_input_nhwc_x3_kernel<uint8_t>, dim3(nm_cdiv((size_t)total, 128)), dim3(128), 0, nm::S(stream),
                          x_nchw, denom, static_cast<__nv_bfloat16*>(x_pad3), plane, N, C, H, W, Cp);
}

}  // extern "C"
