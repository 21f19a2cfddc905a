// First convolution of the net (C_in = 1, 3x3 'same', ReLU, no bias) fused with the following 2x2/2 max-pool,
// forward and backward, on the tcgen05 tensor cores.
//
// With one input channel the GEMM K dimension is only 9, so instead of nine shifted taps the kernels use the
// 4x4 INPUT PATCH of a pool window as the K = 16 dimension:
//
//   forward   D[window, (pos, oc)] = sum_{a,b} patch[window][a][b] * Wext[(pos, oc)][a][b]
//             Wext[(dy,dx,oc)][a][b] = W[oc][a-dy][b-dx] (zero outside the 3x3 support)
//             -> ONE tcgen05.mma (M = 128 windows, N = 128 = 4 window positions x 32 channels, K = 16) per tile;
//             every TMEM lane (= epilogue thread) then holds all four conv values of its pool window for all
//             channels: ReLU + max-pool + arg-max nibble need no cross-lane traffic at all.
//   backward  E[(pos, oc), (a,b)] = sum_windows G[window][(pos, oc)] * patch[window][a][b]
//             G = pooled gradient routed to its arg-max position and masked by ReLU' (the nibble), so
//             dW[oc][r][s] = sum_pos E[(pos,oc)][(dy+r, dx+s)], and a constant-one 17th patch column gives
//             db[oc] = sum_pos E[(pos,oc)][16].  Nothing of the 28x28x32 activation gradient is materialised.
//
// The operand tiles are written by builder warps with ordinary stores in the UMMA swizzled layouts
// (no TMA: the rows are gathered 4x4 patches) and published to the async proxy with fence.proxy.async.
//
// Replaces: conv2d_cpu / conv2d_backward_cpu for the first layer (src/layers/CPU_kernels/Conv2D.cpp:83-146, :150-240),
// Layer_Conv2D.forward/backward incl. np.pad and the bias gradient (src/layers/Conv_wrapepr.py:73-115), Activation_ReLU
// (src/activations/Relu.py:7-14) and Layer_MaxPooling2D 2x2/2 forward/backward (maxpooling_backend.cpp:24-133, :145-238).
#include "nm_common.cuh"
#include "nm_tc_common.cuh"
#include "nm_tc_host.cuh"
#include "nm_tc_pack.cuh"
#include <cstdlib>

#ifdef NM_ABLATE   // timing experiments only: CTA 0 records clock64() per tile and role
#define NM_TRACE(G, tile, slot) do { if ((G).g.trace && blockIdx.x == 0) (G).g.trace[((tile) / gridDim.x) * 8 + (slot)] = clock64(); } while (0)
static long long* g_conv1_trace = nullptr;
extern "C" int nm_tc_debug_set_trace_conv1(void* buf) { g_conv1_trace = static_cast<long long*>(buf); return 0; }
#else
#define NM_TRACE(G, tile, slot) do { } while (0)
#endif

namespace {

using namespace tc;

constexpr int OC = 32;

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ __nv_bfloat162 u2bf(uint32_t u) { return *reinterpret_cast<__nv_bfloat162*>(&u); }
__device__ __forceinline__ uint32_t bf2u(__nv_bfloat162 b) { return *reinterpret_cast<uint32_t*>(&b); }
// byte offset of (row, byte-in-row) inside a tile of 64-byte / 128-byte rows in the SWIZZLE_64B / SWIZZLE_128B layout
// (16-byte chunk index XOR address bits [7,9) / [7,10)); the tile base is 1024-byte aligned.
__device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t byte) { const uint32_t o = row * 64 + byte; return o ^ (((o >> 7) & 3) << 4); }
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte) { const uint32_t o = row * 128 + byte; return o ^ (((o >> 7) & 7) << 4); }

struct Conv1Geom { int N, H, W, PH, PW, PHW; unsigned total; int num_tiles; long long* trace; };   // total pool windows < 2^31

// Geometry as the kernels read it.  HC = 28: H = W = 28 (MNIST, BASELINE configs[1]) known at COMPILE time, so the divisions
// by the pooled width / image size in the per-tile index math fold into multiply-shifts (ncu: the run-time divisions --
// I2F / MUFU.RCP / F2I chains -- and their fix-ups were ~15 % of the instructions of these issue-bound kernels);
// HC = 0: any even H, W from the run-time struct.
template <int HC> struct GeomView {
    const Conv1Geom& g;
    __device__ __forceinline__ int H() const { return HC ? HC : g.H; }
    __device__ __forceinline__ int W() const { return HC ? HC : g.W; }
    __device__ __forceinline__ int PH() const { return HC ? HC / 2 : g.PH; }
    __device__ __forceinline__ int PW() const { return HC ? HC / 2 : g.PW; }
    __device__ __forceinline__ int PHW() const { return HC ? (HC / 2) * (HC / 2) : g.PHW; }
    unsigned total;
    int num_tiles;
};

// Images [n0, n1] touched by the 128 windows of a tile; the bulk-copy producer stages them (contiguous in x).
struct TileSpan { unsigned g0, g1; int n0, n1; };
template <int HC>
__device__ __forceinline__ TileSpan tile_span(const GeomView<HC>& G, int tile) {
    TileSpan t;
    t.g0 = (unsigned)tile * 128u;
    t.g1 = min(t.g0 + 127u, G.total - 1u);
    t.n0 = (int)(t.g0 / (unsigned)G.PHW());
    t.n1 = t.n0 + ((t.g1 - (unsigned)t.n0 * G.PHW()) >= (unsigned)G.PHW() ? 1 : 0);     // a tile touches at most two images
    return t;
}

// Input pixels are either FP32 (the reference's `X.astype(float32) / 255`, examples/test_Conv2D_MNIST.py:27, done on the host)
// or the raw uint8 pixels with that normalisation done here: float(u) * scale.  With scale = 1/255 the product rounds to
// the same bf16 operand as the reference's float32 division for all 256 pixel values (tests/test_tc_layout_host.py).
__device__ __forceinline__ float px_val(const float* img, int i, float) { return img[i]; }
__device__ __forceinline__ float px_val(const uint8_t* img, int i, float scale) { return (float)img[i] * scale; }

// the 4x4 input patch of pool window (n, py, px) from the staged images (zero outside the image: the 'same'
// padding of Conv_wrapepr.py:78-84)
template <typename XT, int HC>
__device__ __forceinline__ void load_patch(const XT* __restrict__ xs, const GeomView<HC>& G, float scale, bool valid, int n_local, int py, int px,
                                           float p[16]) {
    const XT* img = xs + n_local * G.H() * G.W();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int yy = 2 * py - 1 + a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int xx = 2 * px - 1 + b;
            p[a * 4 + b] = (valid && (unsigned)yy < (unsigned)G.H() && (unsigned)xx < (unsigned)G.W()) ? px_val(img, yy * G.W() + xx, scale) : 0.f;
        }
    }
}

// rows a0, a0 + 1 of that patch only (the backward builders split a window's patch between two threads)
template <typename XT, int HC>
__device__ __forceinline__ void load_patch_rows(const XT* __restrict__ xs, const GeomView<HC>& G, float scale, bool valid, int n_local, int py, int px,
                                                int a0, float p[8]) {
    const XT* img = xs + n_local * G.H() * G.W();
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int yy = 2 * py - 1 + a0 + a;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int xx = 2 * px - 1 + b;
            p[a * 4 + b] = (valid && (unsigned)yy < (unsigned)G.H() && (unsigned)xx < (unsigned)G.W()) ? px_val(img, yy * G.W() + xx, scale) : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int F_TILE_BYTES = 128 * 64;          // 128 rows x 64 B (K = 16 bf16 used, SWIZZLE_64B)
constexpr int F_STAGES = 4;                     // A-tile ring (builders -> MMA)
constexpr int F_NBUF = 4;                       // TMEM accumulators (4 x 128 columns): the MMA runs two tiles ahead of each epilogue
                                                // group, which hides the ~900-cycle mbarrier hand-off chain seen in the role trace
constexpr int F_XSTAGES = 6;                    // staged-image ring (bulk-copy producer -> builders): hides the HBM latency
constexpr int F_THREADS = 832;                  // warps 0-7 builders (2 groups), 8 MMA, 9-24 epilogue (4 groups), 25 bulk-copy producer

// PACK: the step's weight re-pack rides in this kernel instead of a launch of its own in front of it (~4 us of every step's
// critical path): the B tile is built from the FP32 conv1 weights directly, and two spare warps refresh the packed copies of
// conv2's and the Dense head's weights (69 k elements over the grid) while the first tiles are computed.  Their consumers are
// later kernels of the stream; their previous readers belong to the step before.
struct Conv1Pack {
    const float* w1;                                                     // conv1 [32][1][3][3]
    const float* w2; __nv_bfloat16* wf; __nv_bfloat16* wd; int OC2, C2;  // conv2 OIHW -> fprop / dgrad operands
    const float* w3; float* wp; __nv_bfloat16* whl; int K, n_out, HW, C3;    // Dense (c,h,w) rows -> packed FP32 + hi/lo
};
constexpr int F_PACK_WARPS = 2;

template <typename XT, int HC, bool PACK>
__global__ void __launch_bounds__(F_THREADS + (PACK ? 32 * F_PACK_WARPS : 0), 1)
conv1_tc_fwd_kernel(const XT* __restrict__ x, float x_scale, const __nv_bfloat16* __restrict__ wext, __nv_bfloat16* __restrict__ y_pad,
                    uint8_t* __restrict__ idx, Conv1Geom Gp, int x_stage_bytes, Conv1Pack pk) {
    const GeomView<HC> G{Gp, Gp.total, Gp.num_tiles};
    constexpr uint64_t TMPL = desc_template(16, 512, LAYOUT_SW64);
    constexpr uint32_t IDESC = idesc_bf16(128, 128, 0, 0);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* b_smem = smem;
    uint8_t* a_smem = smem + F_TILE_BYTES;
    uint8_t* x_smem = a_smem + F_STAGES * F_TILE_BYTES;
    uint64_t* afull = reinterpret_cast<uint64_t*>(x_smem + (size_t)F_XSTAGES * x_stage_bytes);
    uint64_t* aempty = afull + F_STAGES;
    uint64_t* xfull = aempty + F_STAGES;
    uint64_t* xempty = xfull + F_XSTAGES;
    uint64_t* tfull = xempty + F_XSTAGES;
    uint64_t* tempty = tfull + F_NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + F_NBUF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        // builder barriers count WARPS: 128 per-thread mbarrier arrivals per tile would serialise on one shared-memory word
        for (int s = 0; s < F_STAGES; ++s) { mbar_init(&afull[s], 4); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < F_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], 4); }
        for (int a = 0; a < F_NBUF; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, F_NBUF * 128);
    pdl_wait();
    // Wext [128 (pos,oc)][16] bf16 -> B tile
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int row = i >> 1, c = i & 1;
        if constexpr (PACK) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = tcpack::conv1_value(row * 16 + c * 8 + j, pk.w1);
            *reinterpret_cast<uint4*>(b_smem + sw64(row, c * 16)) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        } else {
            *reinterpret_cast<uint4*>(b_smem + sw64(row, c * 16)) = __ldg(reinterpret_cast<const uint4*>(wext + row * 16 + c * 8));
        }
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 25) {
        // ============================== bulk-copy producer: the images a tile touches =========
        if (lane == 0) {
            int xs = 0; uint32_t xphase = 0;
            for (int tile = blockIdx.x; tile < G.num_tiles; tile += gridDim.x) {
                const TileSpan t = tile_span(G, tile);
                const uint32_t bytes = (uint32_t)(t.n1 - t.n0 + 1) * G.H() * G.W() * (uint32_t)sizeof(XT);
                mbar_wait(&xempty[xs], xphase ^ 1);
                mbar_expect_tx(&xfull[xs], bytes);
                bulk_load(x_smem + (size_t)xs * x_stage_bytes, x + (size_t)t.n0 * G.H() * G.W(), bytes, &xfull[xs]);
                if (++xs == F_XSTAGES) { xs = 0; xphase ^= 1; }
            }
        }
    } else if (warp < 8) {
        // ============================== builders: one pool window per thread =================
        // two groups of 4 warps alternate tiles (a tile's build is a chain of mbarrier / shared-memory round trips)
        const int bg = warp >> 2, row = threadIdx.x & 127;
        int it = bg;
        for (int tile = blockIdx.x + bg * gridDim.x; tile < G.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int stage = it % F_STAGES, xs = it % F_XSTAGES;
            const uint32_t phase = (it / F_STAGES) & 1, xphase = (it / F_XSTAGES) & 1;
            const TileSpan t = tile_span(G, tile);
            const unsigned g = t.g0 + row;
            const bool valid = g < G.total;
            int n = t.n0, py = 0, px = 0;
            if (valid) {
                unsigned w = g - (unsigned)t.n0 * G.PHW();
                if (w >= (unsigned)G.PHW()) { w -= G.PHW(); ++n; }
                py = (int)(w / (unsigned)G.PW()); px = (int)w - py * G.PW();
            }
            float p[16];
            if (row == 0) NM_TRACE(G, tile, 0);
            mbar_wait(&xfull[xs], xphase);
            if (row == 0) NM_TRACE(G, tile, 1);
            load_patch(reinterpret_cast<const XT*>(x_smem + (size_t)xs * x_stage_bytes), G, x_scale, valid, n - t.n0, py, px, p);
            // The slot is handed back only once the pixels are IN REGISTERS: the packed words below consume every load, and
            // the empty asm makes the arrive depend on them.  Issuing the loads is not enough -- with the arrive right behind
            // sixteen LDS the bulk copy of a later tile overwrote pixels that had not been read yet (1 launch in 10 at batch
            // 4096 once the index math got cheaper; never seen with the slower run-time divisions in between).
            const uint4 plo = make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
            const uint4 phi = make_uint4(pack_bf16x2(p[8], p[9]), pack_bf16x2(p[10], p[11]), pack_bf16x2(p[12], p[13]), pack_bf16x2(p[14], p[15]));
            asm volatile("" :: "r"(plo.x), "r"(plo.y), "r"(plo.z), "r"(plo.w), "r"(phi.x), "r"(phi.y), "r"(phi.z), "r"(phi.w) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&xempty[xs]);          // the patches are in registers: the image slot may be refilled
            mbar_wait(&aempty[stage], phase ^ 1);
            if (row == 0) NM_TRACE(G, tile, 2);
            uint8_t* at = a_smem + stage * F_TILE_BYTES;
            *reinterpret_cast<uint4*>(at + sw64(row, 0)) = plo;
            *reinterpret_cast<uint4*>(at + sw64(row, 16)) = phi;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[stage]);
            if (row == 0) NM_TRACE(G, tile, 3);
        }
    } else if (warp == 8) {
        // ============================== MMA issuer ===========================================
        const uint32_t b_base = smem_u32(b_smem);
        int stage = 0; uint32_t phase = 0; int it = 0;
        for (int tile = blockIdx.x; tile < G.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it % F_NBUF;
            mbar_wait(&tempty[acc], ((it / F_NBUF) & 1) ^ 1);
            mbar_wait(&afull[stage], phase);
            if (lane == 0) NM_TRACE(G, tile, 4);
            fence_after_sync();
            if (elect_one()) {
                mma_bf16(tmem_base + acc * 128, desc_at(TMPL, smem_u32(a_smem + stage * F_TILE_BYTES)), desc_at(TMPL, b_base), IDESC, 0);
                mma_commit(&aempty[stage]);
                mma_commit(&tfull[acc]);
            }
            __syncwarp();
            if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp < 25) {
        // ============================== epilogue: four groups of 4 warps ======================
        // group = (tile parity, channel half); a thread owns 16 channels of its pool window, in two chunks of 8.  More,
        // lighter warps: the per-window work is a few hundred dependent ALU instructions, and with two 4-warp groups the
        // issue slots sat idle on latency (role trace: ~2400 cycles per tile and group).
        const int g4 = (warp - 9) >> 2, grp = g4 & 1, half = g4 >> 1, q = warp & 3;   // q = TMEM sub-partition of this warp
        int it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < G.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const unsigned g = (unsigned)tile * 128u + q * 32 + lane;
            const bool valid = g < G.total;
            int n = 0, py = 0, px = 0, w = 0;
            if (valid) { n = (int)(g / (unsigned)G.PHW()); w = (int)(g - (unsigned)n * G.PHW()); py = w / G.PW(); px = w - py * G.PW(); }
            __nv_bfloat16* yo = y_pad + (((size_t)n * (G.PH() + 2) + py + 1) * (G.PW() + 2) + px + 1) * OC;
            uint32_t* io = reinterpret_cast<uint32_t*>(idx + ((size_t)n * G.PHW() + w) * (OC / 2));
            const int acc = it % F_NBUF;                                   // F_NBUF is even: buffers {grp, grp + 2} belong to this parity
            mbar_wait(&tfull[acc], (it / F_NBUF) & 1);
            if (lane == 0 && q == 0 && half == 0) NM_TRACE(G, tile, 5);
            fence_after_sync();
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int chunk = half * 2 + h2;                           // channels [8*chunk, 8*chunk + 8)
                uint32_t raw[4][8];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + chunk * 8;
#pragma unroll
                for (int pos = 0; pos < 4; ++pos) tmem_ld_x8(taddr + pos * 32, raw[pos]);
                tmem_ld_wait();
                if (h2 == 1) {                                             // this group's columns are drained
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
                // ReLU + 2x2 max-pool + arg-max on PACKED bf16 pairs (two channels per instruction): the kernel is bound by the
                // half-rate ALU pipe, and the exact-FP32 version cost ~15 ALU instructions per channel (3 x compare / select
                // value / select index, ReLU, nibble assembly, pack) against ~9 here.  The pooled VALUE is unchanged --
                // rounding is monotonic, so bf16(max(relu(v))) == max(relu(bf16(v))) bit for bit.  The arg-max follows the
                // reference's scan (0,0),(0,1),(1,0),(1,1) with its strict '>' (maxpooling_backend.cpp:104: the first maximum
                // wins) on the bf16-ROUNDED conv values, i.e. two positions that differ only below bf16 resolution count as
                // a tie -- the values the rest of the tensor-core plan sees.  A window with no positive value keeps position 0
                // and a cleared ReLU bit, like the reference's pool over the ReLU output.
                uint32_t outw[4], n4[4];
                const __nv_bfloat162 zero2 = u2bf(0u);
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const __nv_bfloat162 h0 = u2bf(pack_bf16x2(__uint_as_float(raw[0][2 * pr]), __uint_as_float(raw[0][2 * pr + 1]))),
                                         h1 = u2bf(pack_bf16x2(__uint_as_float(raw[1][2 * pr]), __uint_as_float(raw[1][2 * pr + 1]))),
                                         h2 = u2bf(pack_bf16x2(__uint_as_float(raw[2][2 * pr]), __uint_as_float(raw[2][2 * pr + 1]))),
                                         h3 = u2bf(pack_bf16x2(__uint_as_float(raw[3][2 * pr]), __uint_as_float(raw[3][2 * pr + 1])));
                    const __nv_bfloat162 a = __hmax2(h0, h1), b = __hmax2(h2, h3), m = __hmax2(a, b);
                    // 0xFFFF per half where the later position is strictly greater: ties keep the earlier one
                    const uint32_t g01 = __hgt2_mask(h1, h0), g23 = __hgt2_mask(h3, h2), gab = __hgt2_mask(b, a), gp = __hgt2_mask(m, zero2);
                    outw[pr] = bf2u(__hmax2(m, zero2));
                    const uint32_t b0 = (gab & g23) | (~gab & g01);              // bit 0 of the winner's position, bit 1 is gab
                    n4[pr] = (b0 & gp & 0x00010001u) | (gab & gp & 0x00020002u) | (gp & 0x00040004u);   // nibble per half
                }
                // n4[pr] holds the nibbles of channels 2pr (bits 0-2) and 2pr+1 (bits 16-18): interleave them into bytes
                const uint32_t w01 = n4[0] | (n4[1] << 8), w23 = n4[2] | (n4[3] << 8);
                const uint32_t nibs = ((w01 | (w01 >> 12)) & 0xFFFFu) | (((w23 | (w23 >> 12)) & 0xFFFFu) << 16);
                if (valid && !NM_DBG(3)) {
                    *reinterpret_cast<uint4*>(yo + chunk * 8) = make_uint4(outw[0], outw[1], outw[2], outw[3]);
                    io[chunk] = nibs;
                }
            }
            if (lane == 0 && q == 0 && half == 0) NM_TRACE(G, tile, 6);
        }
    } else if (PACK && warp >= 26) {
        // ============================== weight re-pack for the later kernels of the step ========
        const int t = blockIdx.x * (32 * F_PACK_WARPS) + ((int)threadIdx.x - 26 * 32), T = gridDim.x * (32 * F_PACK_WARPS);
        for (int i = t; i < pk.OC2 * pk.C2 * 9; i += T) tcpack::conv3x3_elem(i, pk.w2, pk.wf, pk.wd, pk.OC2, pk.C2);
        for (int i = t; i < pk.K * 16; i += T) tcpack::dense_elem(i, pk.w3, pk.wp, pk.whl, pk.K, pk.n_out, pk.HW, pk.C3);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, F_NBUF * 128);
}

// Wext[(dy,dx,oc)][a*4+b] = W[oc][a-dy][b-dx] inside the 3x3 support, else 0
__global__ void pack_conv1_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wext) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 128 * 16) tcpack::conv1_elem(i, w, wext);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int B_G_BYTES = 2 * 128 * 128;        // G: two SW128 column blocks [(pos>>1)] of [128 windows][64 (pos&1, oc)] bf16
constexpr int B_P_BYTES = 128 * 64;             // P: [128 windows][32 = 16 patch + 1 ones + 15 zero] bf16, SW64
constexpr int B_STAGE = B_G_BYTES + B_P_BYTES;  // 40 KB
constexpr int B_STAGES = 2;                     // operand ring (builders -> MMA)
constexpr int B_LSTAGES = 6;                    // staged-input ring (bulk-copy producer -> builders)
constexpr int B_IDX_BYTES = 128 * (OC / 2);     // 2 KB of nibbles per tile
constexpr int B_MAX_GRID = 148;
constexpr int B_THREADS = 576;                  // warps 0-15 builders in two groups alternating tiles (0-3 also the epilogue), 16 MMA, 17 bulk-copy producer

struct Conv1BwdSmem { int x_bytes, dp_bytes, lstage_bytes; };   // per load stage: [x images | dp rows | nibbles]

template <typename XT, int HC>
__global__ void __launch_bounds__(B_THREADS, 1)
conv1_tc_bwd_kernel(const XT* __restrict__ x, float x_scale, const __nv_bfloat16* __restrict__ dp_grid, const uint8_t* __restrict__ idx,
                    float* __restrict__ partial, Conv1Geom Gp, Conv1BwdSmem L) {
    const GeomView<HC> G{Gp, Gp.total, Gp.num_tiles};
    constexpr uint64_t TMPL_A = desc_template(16, 1024, LAYOUT_SW128);     // MN-major: SBO = 8 k-rows of 128 B
    constexpr uint64_t TMPL_B = desc_template(16, 512, LAYOUT_SW64);       //           SBO = 8 k-rows of 64 B
    constexpr uint32_t IDESC = idesc_bf16(64, 32, 1, 1);
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* l_smem = smem + B_STAGES * B_STAGE;
    uint64_t* afull = reinterpret_cast<uint64_t*>(l_smem + (size_t)B_LSTAGES * L.lstage_bytes);
    uint64_t* aempty = afull + B_STAGES;
    uint64_t* lfull = aempty + B_STAGES;
    uint64_t* lempty = lfull + B_LSTAGES;
    uint64_t* done = lempty + B_LSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row_bytes = (G.PW() + 2) * OC * 2;                              // one row of the dP grid
    pdl_trigger();

    // constant half of every P row: column 16 = 1 (bias gradient), columns 17..31 = 0
    for (int i = threadIdx.x; i < B_STAGES * 128; i += blockDim.x) {
        uint8_t* pt = smem + (i / 128) * B_STAGE + B_G_BYTES;
        const int row = i % 128;
        *reinterpret_cast<uint4*>(pt + sw64(row, 32)) = make_uint4(0x00003F80u, 0u, 0u, 0u);   // bf16(1.0) = 0x3F80
        *reinterpret_cast<uint4*>(pt + sw64(row, 48)) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(&afull[s], 8); mbar_init(&aempty[s], 1); }      // one arrival per builder warp
        for (int s = 0; s < B_LSTAGES; ++s) { mbar_init(&lfull[s], 1); mbar_init(&lempty[s], 8); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc(tmem_slot, 64);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const bool has_work = (int)blockIdx.x < G.num_tiles;
    pdl_wait();

    if (warp == 17) {
        // ============================== bulk-copy producer ===================================
        // per tile: the images it touches (contiguous in x), the dP grid rows from its first to its last window
        // (contiguous, incl. the don't-care halo rows when the tile crosses an image boundary) and its 2 KB of nibbles
        if (lane == 0) {
            int ls = 0; uint32_t lphase = 0;
            for (int tile = blockIdx.x; tile < G.num_tiles; tile += gridDim.x) {
                const TileSpan t = tile_span(G, tile);
                const int w0 = (int)(t.g0 - (unsigned)t.n0 * G.PHW()), w1 = (int)(t.g1 - (unsigned)t.n1 * G.PHW());
                const int r0 = t.n0 * (G.PH() + 2) + w0 / G.PW(), r1 = t.n1 * (G.PH() + 2) + w1 / G.PW();
                const uint32_t xb = (uint32_t)(t.n1 - t.n0 + 1) * G.H() * G.W() * (uint32_t)sizeof(XT);
                const uint32_t db = (uint32_t)(r1 - r0 + 1) * row_bytes;
                const uint32_t ib = (uint32_t)(t.g1 - t.g0 + 1) * (OC / 2);
                uint8_t* st = l_smem + (size_t)ls * L.lstage_bytes;
                mbar_wait(&lempty[ls], lphase ^ 1);
                NM_TRACE(G, tile, 0);
                mbar_expect_tx(&lfull[ls], xb + db + ib);
                bulk_load(st, x + (size_t)t.n0 * G.H() * G.W(), xb, &lfull[ls]);
                bulk_load(st + L.x_bytes, reinterpret_cast<const uint8_t*>(dp_grid) + (size_t)r0 * row_bytes, db, &lfull[ls]);
                bulk_load(st + L.x_bytes + L.dp_bytes, idx + (size_t)t.g0 * (OC / 2), ib, &lfull[ls]);
                if (++ls == B_LSTAGES) { ls = 0; lphase ^= 1; }
            }
        }
    } else if (warp < 16) {
        // ============================== builders: (window, half of the channels) per thread ==
        // two groups of 8 warps alternate tiles (group g owns operand stage g): a tile's build is a chain of dependent
        // shared-memory round trips, so one group alone left the issue slots mostly idle
        // Thread -> (window row, channel half hh): a warp owns 16 rows x 2 halves, laid out so that every quarter-warp of a
        // 128-bit shared-memory access touches eight different 16-byte bank groups.  Lane bits: b0 = hh, (b1, b3, b2, b4) =
        // row bits 0..3 -- a quarter-warp holds rows {r, r+1, r+4, r+5} -- and rows with bit 2 set take the two 16-byte
        // pieces of their 32 bytes in the opposite order ("flip"), for the staged dP reads as well as the G-tile writes.
        const int grp = warp >> 3, hh = lane & 1, flip = (lane >> 2) & 1;
        const int row = (warp & 7) * 16 + (((lane >> 1) & 1) | (((lane >> 3) & 1) << 1) | (flip << 2) | (((lane >> 4) & 1) << 3));
        const int stage = grp;
        int it = grp;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < G.num_tiles; tile += 2 * gridDim.x, it += 2) {
            const int ls = it % B_LSTAGES;
            const uint32_t phase = (it >> 1) & 1, lphase = (it / B_LSTAGES) & 1;
            const TileSpan t = tile_span(G, tile);
            const unsigned g = t.g0 + row;
            const bool valid = g < G.total;
            int n = t.n0, py = 0, px = 0;
            if (valid) {
                unsigned w = g - (unsigned)t.n0 * G.PHW();
                if (w >= (unsigned)G.PHW()) { w -= G.PHW(); ++n; }
                py = (int)(w / (unsigned)G.PW()); px = (int)w - py * G.PW();
            }
            const int w0 = (int)(t.g0 - (unsigned)t.n0 * G.PHW()), r0 = t.n0 * (G.PH() + 2) + w0 / G.PW();
            const uint8_t* st = l_smem + (size_t)ls * L.lstage_bytes;
            uint4 da = make_uint4(0, 0, 0, 0), db = da;          // pieces `flip` and `flip ^ 1` of this thread's 16 channels
            uint2 nb = make_uint2(0, 0);
            float p[8];
            if ((threadIdx.x & 255) == 0) NM_TRACE(G, tile, 1);
            mbar_wait(&lfull[ls], lphase);
            if ((threadIdx.x & 255) == 0) NM_TRACE(G, tile, 2);
            if (valid) {
                const uint4* src = reinterpret_cast<const uint4*>(st + L.x_bytes + (size_t)(n * (G.PH() + 2) + py - r0) * row_bytes + px * (OC * 2) + hh * 32);
                da = src[flip]; db = src[flip ^ 1];
                nb = *reinterpret_cast<const uint2*>(st + L.x_bytes + L.dp_bytes + row * (OC / 2) + hh * 8);
            }
            load_patch_rows(reinterpret_cast<const XT*>(st), G, x_scale, valid, n - t.n0, py, px, 2 * hh, p);
            // hand the slot back only once every staged value is in registers (see the forward builder)
            const uint4 pk = make_uint4(pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
            asm volatile("" :: "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w), "r"(da.x), "r"(da.y), "r"(da.z), "r"(da.w),
                               "r"(db.x), "r"(db.y), "r"(db.z), "r"(db.w), "r"(nb.x), "r"(nb.y) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&lempty[ls]);          // everything is in registers: the slot may be refilled
            const uint32_t dw[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};   // 16 channels as bf16 pairs, piece order
            uint32_t out[4][8];
            // arg-max routing + ReLU mask with no table: a nibble is (relu << 2 | position), so for window position pos
            // the byte (nibble ^ (4 | pos) ^ 0x7F) + 1 has bit 7 set exactly when the nibble matches (values stay below
            // 0x80: no carries between bytes); PRMT's sign-replicate mode widens those bits to one 16-bit mask per channel.
            // (A 256-entry selector table in shared memory did the same with one LDS.128 per channel pair, but the
            // kernel is bound by shared-memory wavefronts and those random 16-byte reads were ~40% of them.)
#pragma unroll
            for (int wd = 0; wd < 2; ++wd) {
                const uint32_t nbw = (wd ^ flip) ? nb.y : nb.x;                              // nibbles of that piece's 8 channels
                const uint32_t lo = nbw & 0x07070707u, hi = (nbw >> 4) & 0x07070707u;      // even / odd channels of 4 pairs
#pragma unroll
                for (int pos = 0; pos < 4; ++pos) {
                    const uint32_t k = (0x04040404u | (0x01010101u * pos)) ^ 0x7F7F7F7Fu;
                    const uint32_t sl = (lo ^ k) + 0x01010101u, sh = (hi ^ k) + 0x01010101u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t mask;
                        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(mask) : "r"(sl), "r"(sh), "r"(0xCC88u + 0x1111u * j));
                        out[pos][wd * 4 + j] = dw[wd * 4 + j] & mask;
                    }
                }
            }
            mbar_wait(&aempty[stage], phase ^ 1);
            if ((threadIdx.x & 255) == 0) NM_TRACE(G, tile, 3);
            uint8_t* gt = smem + stage * B_STAGE;
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                uint8_t* blk = gt + (pos >> 1) * (128 * 128);
                const uint32_t base = (pos & 1) * 64 + hh * 32;
                *reinterpret_cast<uint4*>(blk + sw128(row, base + 16 * flip)) = make_uint4(out[pos][0], out[pos][1], out[pos][2], out[pos][3]);
                *reinterpret_cast<uint4*>(blk + sw128(row, base + 16 * (flip ^ 1))) = make_uint4(out[pos][4], out[pos][5], out[pos][6], out[pos][7]);
            }
            *reinterpret_cast<uint4*>(gt + B_G_BYTES + sw64(row, 16 * hh)) = pk;              // patch rows 2hh, 2hh+1 of this window
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&afull[stage]);
            if ((threadIdx.x & 255) == 0) NM_TRACE(G, tile, 4);
        }
        // ============================== epilogue: E -> one FP32 partial per CTA ==============
        if (warp < 4) {
            if (has_work) { mbar_wait(done, 0); fence_after_sync(); }
            // M = 64 accumulator rows sit in TMEM lanes (m % 16) + 32 * (m / 16): lanes 0..15 of each sub-partition
            const int q = warp, m = 16 * q + lane;
#pragma unroll
            for (int mh = 0; mh < 2; ++mh) {
                uint32_t raw[32];
                if (has_work) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + mh * 32;
                    tmem_ld_x16(taddr, raw);
                    tmem_ld_x16(taddr + 16, raw + 16);
                    tmem_ld_wait();
                }
                if (lane < 16) {
                    float4* dst = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * 128 + mh * 64 + m) * 32);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        dst[j / 4] = has_work ? make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]),
                                                            __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    } else {
        // ============================== MMA issuer ===========================================
        int stage = 0; uint32_t phase = 0; uint32_t first = 1;
        for (int tile = blockIdx.x; tile < G.num_tiles; tile += gridDim.x) {
            mbar_wait(&afull[stage], phase);
            if (lane == 0) NM_TRACE(G, tile, 5);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t g_base = smem_u32(smem + stage * B_STAGE), p_base = g_base + B_G_BYTES;
#pragma unroll
                for (int ks = 0; ks < (NM_DBG(6) ? 1 : 8); ++ks) {
                    const uint64_t bdesc = desc_at(TMPL_B, p_base + ks * 16 * 64);
#pragma unroll
                    for (int mh = 0; mh < 2; ++mh)
                        mma_bf16(tmem_base + mh * 32, desc_at(TMPL_A, g_base + mh * (128 * 128) + ks * 16 * 128), bdesc, IDESC,
                                 !(first && ks == 0));
                }
                mma_commit(&aempty[stage]);
            }
            __syncwarp();
            if (lane == 0) NM_TRACE(G, tile, 6);
            first = 0;
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
        }
        if (has_work && elect_one()) mma_commit(done);
        __syncwarp();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 64);
}

// dW[oc][r][s] = sum_cta sum_pos E[cta][(pos,oc)][(dy+r)*4 + dx+s],  db[oc] = sum_cta sum_pos E[cta][(pos,oc)][16]
// 32 outputs x 32 slices of the CTA range per block (<= 5 CTAs per thread, all their loads in flight at once: the kernel
// is a chain of L2 round trips, not bandwidth), slices combined in a fixed order (deterministic).
constexpr int R_SLICES = 32, R_PER = (B_MAX_GRID + R_SLICES - 1) / R_SLICES;
__global__ void __launch_bounds__(32 * R_SLICES)
conv1_tc_bwd_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db, int n_cta) {
    __shared__ float red[R_SLICES][32];
    pdl_trigger();
    pdl_wait();
    const int total = OC * 10;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
    float s = 0.f;
    if (i < total) {
        const int oc = i / 10, t = i % 10, r = t / 3, sx = t % 3;
        float v[R_PER][4];
#pragma unroll
        for (int u = 0; u < R_PER; ++u) {
            const int c = slice * R_PER + u;
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                const int col = t < 9 ? ((pos >> 1) + r) * 4 + (pos & 1) + sx : 16;
                v[u][pos] = c < n_cta ? __ldg(partial + ((size_t)c * 128 + pos * 32 + oc) * 32 + col) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < R_PER; ++u) s += (v[u][0] + v[u][1]) + (v[u][2] + v[u][3]);
    }
    red[slice][threadIdx.x & 31] = s;
    __syncthreads();
    if (slice == 0 && i < total) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < R_SLICES; ++k) a += red[k][threadIdx.x];
        const int oc = i / 10, t = i % 10;
        if (t < 9) dw[oc * 9 + t] = a; else db[oc] = a;
    }
}

int geom(Conv1Geom* g, int N, int H, int W) {
    g->N = N; g->H = H; g->W = W; g->PH = H / 2; g->PW = W / 2; g->PHW = g->PH * g->PW;
    if ((long long)N * g->PHW > 0x7FFFFF00LL) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "conv1: more than 2^31 pool windows");
    g->total = (unsigned)N * (unsigned)g->PHW;
    g->num_tiles = (int)((g->total + 127u) / 128u);
    g->trace = nullptr;
#ifdef NM_ABLATE
    g->trace = g_conv1_trace;
#endif
    return NM_OK;
}

}  // namespace

template <typename XT, int HC, bool PACK>
static int conv1_fwd_launch_k(const XT* x, float scale, const void* w_ext, void* y_pad, void* idx, const Conv1Geom& g, int x_stage, size_t smem,
                              const Conv1Pack& pk, void* stream) {
    static bool attr = false;
    if (!attr) { NM_CUDA(cudaFuncSetAttribute(conv1_tc_fwd_kernel<XT, HC, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    const int grid = g.num_tiles < sm_count() ? g.num_tiles : sm_count();
    return nm::launch_pdl(PACK ? "conv1_tc_fwd_pack" : "conv1_tc_fwd", conv1_tc_fwd_kernel<XT, HC, PACK>, dim3(grid),
                          dim3(F_THREADS + (PACK ? 32 * F_PACK_WARPS : 0)), smem, nm::S(stream), x, scale,
                          static_cast<const __nv_bfloat16*>(w_ext), static_cast<__nv_bfloat16*>(y_pad), static_cast<uint8_t*>(idx), g, x_stage, pk);
}

template <typename XT>
static int conv1_fwd_launch(const XT* x, float scale, const void* w_ext, void* y_pad, void* idx, int N, int H, int W, int OCn, void* stream,
                            const Conv1Pack* pack = nullptr) {
    NM_REQUIRE(x && (w_ext || pack) && y_pad && idx, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad dimensions");
    NM_REQUIRE((H * W * sizeof(XT)) % 16 == 0, "an image must be a multiple of 16 bytes (bulk copies)");
    if (OCn != OC) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv1_fwd_pool: only OC=32 is instantiated");
    Conv1Geom g;
    { int rc = geom(&g, N, H, W); if (rc) return rc; }
    NM_REQUIRE(g.PHW >= 128, "image too small: a 128-window tile must touch at most two images");
    const int x_stage = (2 * H * W * (int)sizeof(XT) + 127) & ~127;
    const size_t smem = (size_t)(1 + F_STAGES) * F_TILE_BYTES + (size_t)F_XSTAGES * x_stage + 256;
    if (smem > 200 * 1024) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv1_fwd_pool: image too large for the staged-image ring");
    const bool k28 = H == 28 && W == 28 && !getenv("NM_CONV1_GENERIC");      // MNIST geometry: compile-time index math (the variable: A/B debugging)
    const Conv1Pack none{};
    if (pack)
        return k28 ? conv1_fwd_launch_k<XT, 28, true>(x, scale, w_ext, y_pad, idx, g, x_stage, smem, *pack, stream)
                   : conv1_fwd_launch_k<XT, 0, true>(x, scale, w_ext, y_pad, idx, g, x_stage, smem, *pack, stream);
    return k28 ? conv1_fwd_launch_k<XT, 28, false>(x, scale, w_ext, y_pad, idx, g, x_stage, smem, none, stream)
               : conv1_fwd_launch_k<XT, 0, false>(x, scale, w_ext, y_pad, idx, g, x_stage, smem, none, stream);
}

template <typename XT>
static int conv1_bwd_launch(const XT* x, float scale, const void* dp_grid, const void* idx, float* dw, float* db,
                            void* workspace, size_t workspace_bytes, int N, int H, int W, int OCn, void* stream) {
    NM_REQUIRE(x && dp_grid && idx && dw && db && workspace, "null pointer");
    NM_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad dimensions");
    NM_REQUIRE((H * W * sizeof(XT)) % 16 == 0, "an image must be a multiple of 16 bytes (bulk copies)");
    if (OCn != OC) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv1_bwd: only OC=32 is instantiated");
    Conv1Geom g;
    { int rc = geom(&g, N, H, W); if (rc) return rc; }
    int grid = sm_count() < B_MAX_GRID ? sm_count() : B_MAX_GRID;
    if (g.num_tiles < grid) grid = g.num_tiles;
    NM_REQUIRE(workspace_bytes >= (size_t)grid * 128 * 32 * sizeof(float), "workspace too small");
    NM_REQUIRE(g.PHW >= 128, "image too small: a 128-window tile must touch at most two images");
    Conv1BwdSmem L;
    L.x_bytes = (2 * H * W * (int)sizeof(XT) + 127) & ~127;
    const int max_rows = (128 + g.PW - 1) / g.PW + 1 + 2;          // pooled rows a tile can span + the 2 halo rows at an image boundary
    L.dp_bytes = max_rows * (g.PW + 2) * OC * 2;
    L.lstage_bytes = L.x_bytes + L.dp_bytes + B_IDX_BYTES;
    const size_t smem = (size_t)B_STAGES * B_STAGE + (size_t)B_LSTAGES * L.lstage_bytes + 256;
    if (smem > 220 * 1024) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_conv1_bwd: image too large for the staged-input ring");
    static bool attr = false;
    if (!attr) {
        NM_CUDA(cudaFuncSetAttribute(conv1_tc_bwd_kernel<XT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        NM_CUDA(cudaFuncSetAttribute(conv1_tc_bwd_kernel<XT, 28>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    int rc = (H == 28 && W == 28 && !getenv("NM_CONV1_GENERIC"))
        ? nm::launch_pdl("conv1_tc_bwd", conv1_tc_bwd_kernel<XT, 28>, dim3(grid), dim3(B_THREADS), smem, nm::S(stream), x, scale,
                         static_cast<const __nv_bfloat16*>(dp_grid), static_cast<const uint8_t*>(idx), static_cast<float*>(workspace), g, L)
        : nm::launch_pdl("conv1_tc_bwd", conv1_tc_bwd_kernel<XT, 0>, dim3(grid), dim3(B_THREADS), smem, nm::S(stream), x, scale,
                         static_cast<const __nv_bfloat16*>(dp_grid), static_cast<const uint8_t*>(idx), static_cast<float*>(workspace), g, L);
    if (rc) return rc;
    return nm::launch_pdl("conv1_tc_bwd_reduce", conv1_tc_bwd_reduce_kernel, dim3(nm_cdiv(OC * 10, 32)), dim3(32 * R_SLICES), 0, nm::S(stream),
                          static_cast<const float*>(workspace), dw, db, grid);
}

extern "C" {

int nm_tc_pack_conv1_weights_bf16(const float* w, void* w_ext, int OCn, void* stream) {
    NM_REQUIRE(w && w_ext, "null pointer");
    if (OCn != OC) return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_pack_conv1_weights_bf16: only OC=32 is instantiated");
    return nm::launch_pdl("pack_conv1_weights", pack_conv1_weights_kernel, dim3(8), dim3(256), 0, nm::S(stream), w, static_cast<__nv_bfloat16*>(w_ext));
}

int nm_tc_conv1_fwd_pool_bf16(const float* x, const void* w_ext, void* y_pad, void* idx, int N, int H, int W, int OCn, void* stream) {
    return conv1_fwd_launch<float>(x, 1.f, w_ext, y_pad, idx, N, H, W, OCn, stream);
}
int nm_tc_conv1_fwd_pool_u8_bf16(const uint8_t* x, float scale, const void* w_ext, void* y_pad, void* idx, int N, int H, int W, int OCn,
                                 void* stream) {
    return conv1_fwd_launch<uint8_t>(x, scale, w_ext, y_pad, idx, N, H, W, OCn, stream);
}

// conv1 forward + the step's weight re-pack in ONE launch (x: float32, or raw uint8 pixels * u8_scale when x_is_u8): the B tile
// comes from the FP32 conv1 weights w1, and the packed copies of conv2's (w2 -> w2_fprop / w2_dgrad) and the Dense head's
// (w3 -> w3_packed / w3_hilo) weights are refreshed as by nm_tc_pack_cnn_weights_bf16
int nm_tc_conv1_fwd_pool_pack_bf16(const void* x, int x_is_u8, float u8_scale, const float* w1, void* y_pad, void* idx, int N, int H, int W, int OCn,
                                   const float* w2, void* w2_fprop, void* w2_dgrad, int OC2, int C2,
                                   const float* w3, float* w3_packed, void* w3_hilo, int K, int n_out, int HW, int C3, void* stream) {
    NM_REQUIRE(w1 && w2 && (w2_fprop || w2_dgrad) && w3 && (w3_packed || w3_hilo), "null pointer");
    NM_REQUIRE(OC2 > 0 && C2 > 0 && K > 0 && n_out > 0 && n_out <= 16 && HW * C3 == K, "bad dimensions");
    const Conv1Pack pk{w1, w2, static_cast<__nv_bfloat16*>(w2_fprop), static_cast<__nv_bfloat16*>(w2_dgrad), OC2, C2,
                       w3, w3_packed, static_cast<__nv_bfloat16*>(w3_hilo), K, n_out, HW, C3};
    if (x_is_u8) return conv1_fwd_launch<uint8_t>(static_cast<const uint8_t*>(x), u8_scale, nullptr, y_pad, idx, N, H, W, OCn, stream, &pk);
    return conv1_fwd_launch<float>(static_cast<const float*>(x), 1.f, nullptr, y_pad, idx, N, H, W, OCn, stream, &pk);
}

size_t nm_tc_conv1_bwd_workspace(int OCn) { (void)OCn; return (size_t)B_MAX_GRID * 128 * 32 * sizeof(float); }

int nm_tc_conv1_bwd_bf16(const float* x, const void* dp_grid, const void* idx, float* dw, float* db,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int OCn, void* stream) {
    return conv1_bwd_launch<float>(x, 1.f, dp_grid, idx, dw, db, workspace, workspace_bytes, N, H, W, OCn, stream);
}
int nm_tc_conv1_bwd_u8_bf16(const uint8_t* x, float scale, const void* dp_grid, const void* idx, float* dw, float* db,
                            void* workspace, size_t workspace_bytes, int N, int H, int W, int OCn, void* stream) {
    return conv1_bwd_launch<uint8_t>(x, scale, dp_grid, idx, dw, db, workspace, workspace_bytes, N, H, W, OCn, stream);
}

}  // extern "C"
