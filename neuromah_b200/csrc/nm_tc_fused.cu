// CUDA-core fused kernel that completes the bf16 tensor-core training path around the tcgen05 kernels:
//   Dense dgrad + max-pool backward (arg-max routing) + ReLU backward -> conv2's dY in one pass (+ conv2's bias gradient).
// HBM-bound: coalesced 16-byte accesses, nothing materialised that a consumer can recompute from a nibble.
#include "nm_common.cuh"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// ------------------------------------------------------------------------------------------------
// Dense dgrad + pool backward + ReLU backward: dY of the last conv, straight into its zero-haloed bf16
// layout [B, 2PH+2, 2PW+2, C].  Also the conv bias gradient (sum of the routed values), one partial per CTA.
// A thread owns ONE (pooled pixel, 8-channel group) for every image its CTA processes and keeps that
// [n_out x 8] slice of the Dense weights in registers: the weights are read once per CTA instead of once
// per image (the per-image version re-read 125 KB x 4096 images from L2 and sat at 10% of HBM).
// The next image's nibbles are prefetched while the current one is routed and stored.
// ------------------------------------------------------------------------------------------------
// POOLED = true: only the ReLU-masked gradient of the POOLED tensor is written (bf16 [B, PH*PW, C]); the routing to the
// arg-max position is left to the consumer (nm_tc_conv3x3_bwd_pool_bf16 builds the dY tile in shared memory from it),
// which removes the 4x larger dY tensor and its HBM round trip.
template <int NO, bool POOLED>
__global__ void __launch_bounds__(416, 1)
dense_bwd_unpool_kernel(const float* __restrict__ dlogits, const float* __restrict__ wp, const uint8_t* __restrict__ idx,
                        __nv_bfloat16* __restrict__ dy_pad, float* __restrict__ dbias_cta, float* __restrict__ dbias,
                        unsigned* __restrict__ ticket, int B, int PH, int PW, int C, int n_out) {
    extern __shared__ float red[];                       // [items][8] for the bias-gradient reduction
    pdl_trigger();
    pdl_wait();
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
    // dlogits of the next DLC images of this CTA are staged in shared memory once per chunk: read per image from global
    // memory they were a dependent L2 round trip at the top of every iteration (28% of the stall samples)
    constexpr int DLC = 32;                      // even
    __shared__ float dls[DLC][NO];
    int n = blockIdx.x;
    uint32_t nibs_next = (active && n < B) ? __ldg(idx32 + (size_t)n * items + item) : 0u;
    uint32_t nibs_next_b = (active && n + (int)gridDim.x < B) ? __ldg(idx32 + (size_t)(n + gridDim.x) * items + item) : 0u;
    for (int n0 = blockIdx.x; n0 < B; n0 += DLC * gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < DLC * NO; i += blockDim.x) {
            const int k = i / NO, j = i % NO, nn = n0 + k * gridDim.x;
            dls[k][j] = (nn < B && j < n_out) ? __ldg(dlogits + (size_t)nn * n_out + j) : 0.f;
        }
        __syncthreads();
        // two images per iteration: 16 independent FMA chains per thread instead of 8 (the kernel is latency-bound at
        // 13 warps per SM), the weight registers are read once for both
        auto emit = [&](int n, uint32_t nibs, const float (&d)[8]) {
            if constexpr (POOLED) {
                float val[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    val[c] = ((nibs >> (4 * c)) & 4u) ? d[c] : 0.f;          // ReLU mask
                    bsum[c] += val[c];
                }
                *reinterpret_cast<uint4*>(dy_pad + ((size_t)n * items + item) * 8) =
                    make_uint4(pack2(val[0], val[1]), pack2(val[2], val[3]), pack2(val[4], val[5]), pack2(val[6], val[7]));
            } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) bsum[c] += ((nibs >> (4 * c)) & 4u) ? d[c] : 0.f;   // ReLU mask (bias gradient)
            // ReLU mask + arg-max routing on the packed bf16 pairs: a nibble is (relu << 2 | position), so the byte
            // (nibble ^ (4 | pos) ^ 0x7F) + 1 has bit 7 set exactly where the channel is live AND its maximum sat at pos;
            // PRMT's sign-replicate mode turns those bits into one 16-bit mask per channel (nm_tc_conv1.cu uses the same trick)
            const uint32_t pw[4] = {pack2(d[0], d[1]), pack2(d[2], d[3]), pack2(d[4], d[5]), pack2(d[6], d[7])};
            const uint32_t lo = nibs & 0x07070707u, hi = (nibs >> 4) & 0x07070707u;
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                const int Y = 2 * py + (pos >> 1) + 1, X = 2 * px + (pos & 1) + 1;
                const uint32_t kk = (0x04040404u | (0x01010101u * pos)) ^ 0x7F7F7F7Fu;
                const uint32_t sl = (lo ^ kk) + 0x01010101u, sh = (hi ^ kk) + 0x01010101u;
                uint32_t m[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m[j]) : "r"(sl), "r"(sh), "r"(0xCC88u + 0x1111u * j));
                const uint4 o = make_uint4(pw[0] & m[0], pw[1] & m[1], pw[2] & m[2], pw[3] & m[3]);
                if (!NM_DBG(1)) *reinterpret_cast<uint4*>(dy_pad + (((size_t)n * H2 + Y) * W2 + X) * C + g * 8) = o;
                else if (o.x == 0x12345678u) dbias_cta[0] = 1.f;
            }
            }
        };
        for (int k = 0; k < DLC && n < B; k += 2, n += 2 * gridDim.x) {
            const uint32_t nibsA = nibs_next, nibsB = nibs_next_b;
            const int nB = n + gridDim.x;
            const bool hasB = nB < B;                          // DLC is even: a pair never straddles a dlogits chunk
            if (active) {
                const int na = n + 2 * gridDim.x, nb = n + 3 * gridDim.x;
                if (na < B) nibs_next = __ldg(idx32 + (size_t)na * items + item);
                if (nb < B) nibs_next_b = __ldg(idx32 + (size_t)nb * items + item);
            }
            if (!active) continue;
            float dA[8], dB[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) { dA[c] = 0.f; dB[c] = 0.f; }
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                if (NM_DBG(0)) break;
                const float sa = dls[k][j], sb = dls[k + 1][j];               // same address for the whole CTA: broadcast
#pragma unroll
                for (int c = 0; c < 8; ++c) { dA[c] = fmaf(sa, wv[j][c], dA[c]); dB[c] = fmaf(sb, wv[j][c], dB[c]); }
            }
            emit(n, nibsA, dA);
            if (hasB) emit(nB, nibsB, dB);
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
    // the last CTA to finish sums the per-CTA partials in CTA order (deterministic) -- no separate reduction launch
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence();
        if ((int)threadIdx.x < C) {
            float s = 0.f;
            for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(dbias_cta + (size_t)b * C + threadIdx.x);
            dbias[threadIdx.x] = s;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

constexpr int kNO = 16;

}  // namespace

extern "C" {

size_t nm_tc_dense_bwd_unpool_workspace(int C) { return (size_t)148 * 2 * C * sizeof(float); }

static int launch_dense_bwd(bool pooled, const float* dlogits, const float* w_packed, const void* idx, void* out,
                            float* dbias, void* workspace, size_t workspace_bytes,
                            int B, int PH, int PW, int C, int n_out, void* stream) {
    NM_REQUIRE(dlogits && w_packed && idx && out && dbias && workspace, "null pointer");
    NM_REQUIRE(B > 0 && PH > 0 && PW > 0 && C % 8 == 0 && n_out > 0 && n_out <= kNO, "bad dimensions");
    const int items = PH * PW * (C / 8);
    if (items > 416 || C > 416)
        return nm::fail(NM_ERR_UNSUPPORTED, "%s%s", "nm_tc_dense_bwd_*: needs PH*PW*C/8 <= 416 (one thread per item)");
    int grid = B < 148 ? B : 148;
    NM_REQUIRE(workspace_bytes >= (size_t)grid * C * sizeof(float), "workspace too small");
    const size_t smem = (size_t)items * 8 * sizeof(float);
    const uint8_t* ip = static_cast<const uint8_t*>(idx);
    __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out);
    float* part = static_cast<float*>(workspace);
    static unsigned* ticket = nullptr;             // last-block ticket, left at zero by every launch (one device per process)
    if (!ticket) { NM_CUDA(cudaMalloc(&ticket, sizeof(unsigned))); NM_CUDA(cudaMemset(ticket, 0, sizeof(unsigned))); }
    // the [n_out x 8] weight slice lives in registers: instantiate the common head width exactly
#define NM_LAUNCH_DB(NO_, P_) nm::launch_pdl("dense_bwd_unpool", dense_bwd_unpool_kernel<NO_, P_>, dim3(grid), dim3(416), smem, nm::S(stream), \
                                            dlogits, w_packed, ip, op, part, dbias, ticket, B, PH, PW, C, n_out)
    if (pooled) return n_out <= 10 ? NM_LAUNCH_DB(10, true) : NM_LAUNCH_DB(kNO, true);
    return n_out <= 10 ? NM_LAUNCH_DB(10, false) : NM_LAUNCH_DB(kNO, false);
#undef NM_LAUNCH_DB
}

int nm_tc_dense_bwd_unpool_bf16(const float* dlogits, const float* w_packed, const void* idx, void* dy_pad,
                                float* dbias, void* workspace, size_t workspace_bytes,
                                int B, int PH, int PW, int C, int n_out, void* stream) {
    return launch_dense_bwd(false, dlogits, w_packed, idx, dy_pad, dbias, workspace, workspace_bytes, B, PH, PW, C, n_out, stream);
}

int nm_tc_dense_bwd_pool_bf16(const float* dlogits, const float* w_packed, const void* idx, void* dpool,
                              float* dbias, void* workspace, size_t workspace_bytes,
                              int B, int PH, int PW, int C, int n_out, void* stream) {
    return launch_dense_bwd(true, dlogits, w_packed, idx, dpool, dbias, workspace, workspace_bytes, B, PH, PW, C, n_out, stream);
}

}  // extern "C"
