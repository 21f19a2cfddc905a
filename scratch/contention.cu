// Do tcgen05.mma operand reads (smem) and TMA writes (into other smem) contend?  Three modes per config:
//   loads only / MMAs only / both concurrently.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lcuda
#include "../neuromah_b200/csrc/nm_tc_common.cuh"
#include "../neuromah_b200/csrc/nm_tc_host.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace nm { thread_local char g_err[512] = ""; std::atomic<unsigned long long> g_launches{0}; }
using namespace tc;

constexpr int STAGES = 6;
template <int N, int ROWB, int NTAPS, int KS>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int n_tiles, int box_rows, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[STAGES], mbar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage_bytes = ((uint32_t)box_rows * ROWB + 1023u) & ~1023u;
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); mbar_init(&mbar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    long long t0 = clock64();
    if (warp == 1 && lane == 0 && (mode & 1)) {
        // loads: keep STAGES boxes in flight, re-using a stage as soon as its own load has landed
        uint32_t phase[STAGES] = {0};
        for (int t = 0; t < n_tiles + STAGES; ++t) {
            const int s = t % STAGES;
            if (t >= STAGES) { mbar_wait(&full[s], phase[s]); phase[s] ^= 1; }
            if (t < n_tiles) {
                mbar_expect_tx(&full[s], (uint32_t)box_rows * ROWB);
                tma_load_2d(smem + 65536 + (size_t)s * stage_bytes, &map, 0, (blockIdx.x * n_tiles + t) * 128, &full[s]);
            }
        }
    }
    if (warp == 2 && (mode & 2)) {
        constexpr uint32_t LAYOUT = ROWB == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
        constexpr uint64_t T = desc_template(16, 8 * ROWB, LAYOUT);
        constexpr uint32_t IDESC = idesc_bf16(128, N, 0, 0);
        const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
        if (elect_one()) {
            for (int t = 0; t < n_tiles; ++t) {
#pragma unroll
                for (int tap = 0; tap < NTAPS; ++tap)
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk)
                        mma_bf16(tmem_base, desc_at(T, a + ((tap / 3) * 16 + tap % 3) * ROWB + kk * 32), desc_at(T, b + tap * 1024 + kk * 32), IDESC, 1);
            }
            mma_commit(&mbar);
            mbar_wait(&mbar, 0);
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, int ROWB, int NTAPS, int KS>
void run(const char* name, void* gbuf, size_t rows_total) {
    const int box_rows = 162, n_tiles = 55, grid = 148;
    CUtensorMap m;
    make_map_2d(&m, gbuf, ROWB / 2, rows_total, ROWB, ROWB / 2, box_rows, ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    long long* d; cudaMalloc(&d, grid * sizeof(long long));
    auto kern = k<N, ROWB, NTAPS, KS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int mode = 1; mode <= 3; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            kern<<<grid, 128, 65536 + STAGES * 24 * 1024>>>(m, n_tiles, box_rows, mode, d);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s FAILED %s\n", name, cudaGetErrorString(e)); exit(1); }
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(grid); cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("%-44s mode %d (%s)  %7.1f us  max-CTA %8lld cyc  = %6.0f cyc/tile\n", name, mode,
               mode == 1 ? "loads" : mode == 2 ? "MMAs " : "both ", ms * 1e3, mx, (double)mx / n_tiles);
    }
    cudaFree(d);
}

int main() {
    cudaSetDevice(0); cudaFree(0);
    size_t rows = (size_t)148 * 55 * 128 + 4096;
    void* g; cudaMalloc(&g, rows * 128); cudaMemset(g, 0, rows * 128);
    run<64, 64, 9, 2>("fprop: N64, 64B rows, 18 MMA/tile, 10KB box", g, rows);
    run<32, 128, 9, 4>("dgrad: N32, 128B rows, 36 MMA/tile, 20KB box", g, rows);
    run<96, 128, 3, 4>("dgrad fold: N96, 128B rows, 12 MMA/tile", g, rows);
    run<256, 128, 1, 4>("gemm: N256, 128B rows, 4 MMA/tile", g, rows);
    return 0;
}
