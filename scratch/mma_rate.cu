// tcgen05.mma issue-rate microbenchmark: cycles per MMA for the operand shapes / layouts the conv kernels use.
// smem content is irrelevant (zeros); one thread issues N MMAs back to back, commits, waits, reads clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 scratch/mma_rate.cu -o scratch/mma_rate
#include "../neuromah_b200/csrc/nm_tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace tc;



// compile-time pattern: PAT 0 = conv taps (TR x TS taps, KS k-steps, shifts of (r*16+s) rows of ROWB bytes, B advances per tap)
//                       PAT 1 = wgrad: 9 accumulators, A fixed, B shifted by tap rows
template <int M, int N, int AMN, int BMN, uint32_t AL, uint32_t BL, uint32_t ALBO, uint32_t ASBO, uint32_t BLBO, uint32_t BSBO,
          int PAT, int TR, int TS, int KS, int ROWB, int BTAP, int DSTEP>
__global__ void __launch_bounds__(128, 1) rate_kernel(int rounds, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && elect_one()) {      // converged warp + elect: UTCHMMA issues without the compiler's ELECT retry loop
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 96 * 1024);
        constexpr uint64_t ta = desc_template(ALBO, ASBO, AL), tb = desc_template(BLBO, BSBO, BL);
        constexpr uint32_t idesc = idesc_bf16(M, N, AMN, BMN);
        long long t0 = 0;
        for (int r = -1; r < rounds; ++r) {
            if (r == 0) { mma_commit(&bar); mbar_wait(&bar, 0); t0 = clock64(); }
#pragma unroll
            for (int tr = 0; tr < TR; ++tr)
#pragma unroll
                for (int ts = 0; ts < TS; ++ts)
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) {
                        const int tap = tr * TS + ts;
                        if constexpr (PAT == 0)
                            mma_bf16(tmem_base, desc_at(ta, a_base + (tr * 16 + ts) * ROWB + kk * 32),
                                     desc_at(tb, b_base + tap * BTAP + kk * 32), idesc, 1);
                        else
                            mma_bf16(tmem_base + tap * DSTEP, desc_at(ta, a_base + kk * 16 * 128),
                                     desc_at(tb, b_base + ((tr * 16 + ts) + kk * 16) * ROWB), idesc, 1);
                    }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 1);
        out[blockIdx.x] = clock64() - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <typename K>
static void run(K kern, const char* name, int M, int N, int n_pat, int grid) {
    long long* d; cudaMalloc(&d, grid * sizeof(long long));
    const int rounds = 64;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    kern<<<grid, 128, 192 * 1024>>>(rounds, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-64s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    double per = (double)mx / (rounds * n_pat);
    double floor_c = (double)(M > 128 ? M : 128) * N / 256.0;
    double bytes = M * 32.0 + N * 32.0;
    printf("%-66s grid %3d  cyc/MMA %6.1f  floor %5.1f  smem %5.0f B/MMA -> %5.1f B/cyc  %5.1f%% of MMA peak\n", name, grid, per,
           floor_c, bytes, bytes / per, 100.0 * (M * N / 256.0) / per);
    cudaFree(d);
}

int main() {
    cudaSetDevice(0); cudaFree(0);
    constexpr uint32_t S128 = LAYOUT_SW128, S64 = LAYOUT_SW64;
    for (int grid : {1, 148}) {
        run(rate_kernel<128, 64, 0, 0, S64, S64, 16, 512, 16, 512, 0, 3, 3, 2, 64, 64 * 64, 0>, "fprop  M128 N64  K-major SW64 (C=32), 9 taps x 2 k", 128, 64, 18, grid);
        run(rate_kernel<128, 192, 0, 0, S64, S64, 16, 512, 16, 512, 0, 3, 1, 2, 64, 192 * 64, 0>, "fprop' M128 N192 K-major SW64 (3 s-taps folded into N), 3 x 2", 128, 192, 6, grid);
        run(rate_kernel<128, 32, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 3, 3, 4, 128, 32 * 128, 0>, "dgrad  M128 N32  K-major SW128 (OC=64), 9 taps x 4 k", 128, 32, 36, grid);
        run(rate_kernel<128, 96, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 3, 1, 4, 128, 96 * 128, 0>, "dgrad' M128 N96  K-major SW128 (3 s-taps folded into N), 3 x 4", 128, 96, 12, grid);
        run(rate_kernel<128, 256, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 1, 1, 4, 128, 0, 0>, "gemm   M128 N256 K-major SW128", 128, 256, 4, grid);
        run(rate_kernel<128, 128, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 1, 1, 4, 128, 0, 0>, "gemm   M128 N128 K-major SW128", 128, 128, 4, grid);
        run(rate_kernel<128, 64, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 1, 1, 4, 128, 0, 0>, "gemm   M128 N64  K-major SW128", 128, 64, 4, grid);
        run(rate_kernel<128, 32, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 1, 1, 4, 128, 0, 0>, "gemm   M128 N32  K-major SW128", 128, 32, 4, grid);
        run(rate_kernel<128, 16, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 1, 1, 4, 128, 0, 0>, "dense  M128 N16  K-major SW128", 128, 16, 4, grid);
        run(rate_kernel<64, 64, 0, 0, S128, S128, 16, 1024, 16, 1024, 0, 1, 1, 4, 128, 0, 0>, "gemm   M64  N64  K-major SW128", 64, 64, 4, grid);
        run(rate_kernel<64, 32, 1, 1, S128, S64, 16, 1024, 16, 512, 1, 3, 3, 8, 64, 0, 32>, "wgrad  M64  N32  MN-major A SW128 / B SW64, 9 acc x 8 k", 64, 32, 72, grid);
        run(rate_kernel<64, 96, 1, 1, S128, S64, 16, 1024, 64, 512, 1, 3, 1, 8, 64, 0, 96>, "wgrad' M64  N96  MN-major, B LBO=64 B (3 s-taps as N chunks)", 64, 96, 24, grid);
        run(rate_kernel<64, 32, 1, 0, S128, S64, 16, 1024, 16, 512, 1, 3, 3, 8, 64, 0, 32>, "wgrad* M64  N32  A MN-major, B K-major(dummy), 9 acc x 8 k", 64, 32, 72, grid);
        run(rate_kernel<64, 32, 0, 1, S128, S64, 16, 1024, 16, 512, 1, 3, 3, 8, 64, 0, 32>, "wgrad* M64  N32  A K-major(dummy), B MN-major, 9 acc x 8 k", 64, 32, 72, grid);
    }
    return 0;
}
