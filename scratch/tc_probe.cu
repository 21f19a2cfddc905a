// tcgen05 / TMA probe: validates the smem-descriptor, instruction-descriptor and TMEM-layout
// assumptions the tensor-core kernels rely on, one hypothesis per launch, exact integer data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 scratch/tc_probe.cu -o scratch/tc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct ProbeOp { uint32_t a_off, b_off; };
struct ProbeParams {
    uint32_t a_bytes, b_bytes;
    int a_c0, a_c1, b_c0, b_c1;
    unsigned long long a_desc, b_desc;     // descriptor without the start address
    uint32_t idesc;
    int n_ops;
    ProbeOp ops[32];
    int ncols;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             ProbeParams P, float* out, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* a_tile = smem;
    uint8_t* b_tile = smem + 65536;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_load)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    bool ok = true;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar_load)), "r"(P.a_bytes + P.b_bytes));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(smem_u32(a_tile)), "l"(&map_a), "r"(P.a_c0), "r"(P.a_c1), "r"(smem_u32(&bar_load)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(smem_u32(b_tile)), "l"(&map_b), "r"(P.b_c0), "r"(P.b_c1), "r"(smem_u32(&bar_load)) : "memory");
        // wait for the loads
        uint32_t done = 0; int spins = 0;
        while (!done && spins < 2000000) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar_load)), "r"(0));
            ++spins;
        }
        if (!done) { ok = false; status[0] = 1; }
        if (ok) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int i = 0; i < P.n_ops; ++i) {
                unsigned long long ad = P.a_desc | (unsigned long long)(((smem_u32(a_tile) + P.ops[i].a_off) >> 4) & 0x3FFF);
                unsigned long long bd = P.b_desc | (unsigned long long)(((smem_u32(b_tile) + P.ops[i].b_off) >> 4) & 0x3FFF);
                uint32_t acc = i > 0;
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                             :: "r"(tmem_base), "l"(ad), "l"(bd), "r"(P.idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar_mma)) : "memory");
        }
    }
    __syncthreads();
    if (status[0] == 0) {
        uint32_t done = 0; int spins = 0;
        while (!done && spins < 2000000) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar_mma)), "r"(0));
            ++spins;
        }
        if (!done) status[0] = 2;
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (done) {
            for (int c = 0; c < P.ncols; c += 8) {
                uint32_t v[8];
                uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * P.ncols + c + j] = __uint_as_float(v[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256));
}

// ------------------------------------------------------------------ host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

static CUtensorMap make_map(void* gptr, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_inner,
                            uint32_t box_rows, CUtensorMapSwizzle sw) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

static unsigned long long make_desc(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type, uint32_t base_offset) {
    unsigned long long d = 0;
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ULL << 46;                                   // version = 1 (Blackwell)
    d |= (unsigned long long)(base_offset & 7) << 49;
    d |= (unsigned long long)(layout_type & 7) << 61;
    return d;
}

static uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;              // c_format = F32
    d |= 1u << 7;              // a_format = BF16
    d |= 1u << 10;             // b_format = BF16
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

struct Ctx {
    std::vector<float> Ah, Bh;       // host copies (float) of the global matrices
    int a_inner, a_rows, b_inner, b_rows;
    __nv_bfloat16 *Ad, *Bd;
    float* out_d; int* status_d;
};

static void fill(std::vector<float>& h, __nv_bfloat16** d, int n, unsigned seed) {
    h.resize(n);
    std::vector<__nv_bfloat16> t(n);
    srand(seed);
    for (int i = 0; i < n; ++i) { int v = (rand() % 7) - 3; h[i] = (float)v; t[i] = __float2bfloat16((float)v); }
    CK(cudaMalloc(d, n * sizeof(__nv_bfloat16)));
    CK(cudaMemcpy(*d, t.data(), n * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
}

// run one experiment; returns the dumped [128 x ncols] TMEM block
static std::vector<float> run(Ctx& c, const CUtensorMap& ma, const CUtensorMap& mb, ProbeParams P, int* status) {
    CK(cudaMemset(c.out_d, 0xFF, 128 * 256 * sizeof(float)));
    CK(cudaMemset(c.status_d, 0, sizeof(int)));
    probe_kernel<<<1, 128, 131072>>>(ma, mb, P, c.out_d, c.status_d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(status, c.status_d, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float> o(128 * P.ncols);
    CK(cudaMemcpy(o.data(), c.out_d, o.size() * sizeof(float), cudaMemcpyDeviceToHost));
    return o;
}

// compare TMEM dump with expected D[M x N] under a lane mapping
static void report(const char* name, const std::vector<float>& o, int ncols, const std::vector<float>& D, int M, int N,
                   int status, bool m64_layout) {
    if (status) { printf("%-58s STATUS=%d (barrier timeout)\n", name, status); return; }
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m) {
        int lane = m64_layout ? (m % 16) + 32 * (m / 16) : m;
        for (int n = 0; n < N; ++n) {
            double e = fabs((double)o[lane * ncols + n] - (double)D[m * N + n]);
            if (!(e <= 1e-3)) ++bad;
            if (e > maxerr || e != e) maxerr = e;
        }
    }
    printf("%-58s %s  (bad %d / %d, max err %.3g)\n", name, bad == 0 ? "PASS" : "FAIL", bad, M * N, maxerr);
}

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&g_encode, cudaEnableDefault, &qres));
    if (!g_encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    Ctx c;
    CK(cudaMalloc(&c.out_d, 128 * 256 * sizeof(float)));
    CK(cudaMalloc(&c.status_d, sizeof(int)));
    int status;

    // ===================== E1: K-major SW128, A[256 x 64], B[64 x 64], M=128, N=64, K=64 =====================
    {
        const int AR = 256, K = 64, N = 64;
        fill(c.Ah, &c.Ad, AR * K, 1); fill(c.Bh, &c.Bd, N * K, 2);
        CUtensorMap ma = make_map(c.Ad, K, AR, K * 2, 64, 192, CU_TENSOR_MAP_SWIZZLE_128B);   // super tile of 192 rows
        CUtensorMap mb = make_map(c.Bd, K, N, K * 2, 64, N, CU_TENSOR_MAP_SWIZZLE_128B);
        auto expect = [&](int r0) {
            std::vector<float> D(128 * N);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                float s = 0; for (int k = 0; k < K; ++k) s += c.Ah[(r0 + m) * K + k] * c.Bh[n * K + k];
                D[m * N + n] = s;
            }
            return D;
        };
        ProbeParams P{}; P.a_bytes = 192 * 128; P.b_bytes = N * 128; P.a_c0 = 0; P.a_c1 = 0; P.b_c0 = 0; P.b_c1 = 0;
        P.idesc = make_idesc(128, N, 0, 0); P.ncols = N; P.n_ops = 4;
        P.b_desc = make_desc(16, 1024, 2, 0);
        for (int shift : {0, 1, 3, 8, 17, 33}) {
            for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                if (shift % 8 == 0 && bo_mode == 1) continue;
                P.a_desc = make_desc(16, 1024, 2, bo_mode ? (shift % 8) : 0);
                for (int i = 0; i < 4; ++i) { P.ops[i].a_off = shift * 128 + i * 32; P.ops[i].b_off = i * 32; }
                auto o = run(c, ma, mb, P, &status);
                char nm[128]; snprintf(nm, sizeof nm, "E1 K-major SW128 M128 N64 K64 row-shift %d base_off=%d", shift, bo_mode ? shift % 8 : 0);
                report(nm, o, N, expect(shift), 128, N, status, false);
            }
        }
        // M = 64 accumulator layout with the same operands
        P.idesc = make_idesc(64, N, 0, 0); P.a_desc = make_desc(16, 1024, 2, 0);
        for (int i = 0; i < 4; ++i) { P.ops[i].a_off = i * 32; P.ops[i].b_off = i * 32; }
        auto o = run(c, ma, mb, P, &status);
        report("E1b K-major SW128 M64 (lanes m%16+32*(m/16))", o, N, expect(0), 64, N, status, true);
        report("E1b K-major SW128 M64 (lanes = m)           ", o, N, expect(0), 64, N, status, false);
        cudaFree(c.Ad); cudaFree(c.Bd);
    }

    // ===================== E2: K-major SW64, A[256 x 32], B[64 x 32], M=128, N=64, K=32 ======================
    {
        const int AR = 256, K = 32, N = 64;
        fill(c.Ah, &c.Ad, AR * K, 3); fill(c.Bh, &c.Bd, N * K, 4);
        CUtensorMap ma = make_map(c.Ad, K, AR, K * 2, 32, 192, CU_TENSOR_MAP_SWIZZLE_64B);
        CUtensorMap mb = make_map(c.Bd, K, N, K * 2, 32, N, CU_TENSOR_MAP_SWIZZLE_64B);
        auto expect = [&](int r0) {
            std::vector<float> D(128 * N);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                float s = 0; for (int k = 0; k < K; ++k) s += c.Ah[(r0 + m) * K + k] * c.Bh[n * K + k];
                D[m * N + n] = s;
            }
            return D;
        };
        ProbeParams P{}; P.a_bytes = 192 * 64; P.b_bytes = N * 64; P.idesc = make_idesc(128, N, 0, 0); P.ncols = N; P.n_ops = 2;
        P.b_desc = make_desc(16, 512, 4, 0);
        for (int shift : {0, 1, 2, 5, 16, 18, 34}) {
            for (int bo_mode = 0; bo_mode < 3; ++bo_mode) {
                if (shift % 8 == 0 && bo_mode) continue;
                int bo = bo_mode == 0 ? 0 : (bo_mode == 1 ? ((shift * 64) >> 7) & 7 : shift % 8);
                P.a_desc = make_desc(16, 512, 4, bo);
                for (int i = 0; i < 2; ++i) { P.ops[i].a_off = shift * 64 + i * 32; P.ops[i].b_off = i * 32; }
                auto o = run(c, ma, mb, P, &status);
                char nm[128]; snprintf(nm, sizeof nm, "E2 K-major SW64 M128 N64 K32 row-shift %d base_off=%d", shift, bo);
                report(nm, o, N, expect(shift), 128, N, status, false);
            }
        }
        cudaFree(c.Ad); cudaFree(c.Bd);
    }

    // ===================== E3: N=16 and N=32 tiles (Dense head, dgrad) ======================================
    {
        const int AR = 128, K = 64;
        for (int N : {16, 32}) {
            fill(c.Ah, &c.Ad, AR * K, 5); fill(c.Bh, &c.Bd, N * K, 6);
            CUtensorMap ma = make_map(c.Ad, K, AR, K * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
            CUtensorMap mb = make_map(c.Bd, K, N, K * 2, 64, N, CU_TENSOR_MAP_SWIZZLE_128B);
            std::vector<float> D(128 * N);
            for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
                float s = 0; for (int k = 0; k < K; ++k) s += c.Ah[m * K + k] * c.Bh[n * K + k];
                D[m * N + n] = s;
            }
            ProbeParams P{}; P.a_bytes = 128 * 128; P.b_bytes = N * 128; P.idesc = make_idesc(128, N, 0, 0); P.ncols = N; P.n_ops = 4;
            P.a_desc = make_desc(16, 1024, 2, 0); P.b_desc = make_desc(16, 1024, 2, 0);
            for (int i = 0; i < 4; ++i) { P.ops[i].a_off = i * 32; P.ops[i].b_off = i * 32; }
            auto o = run(c, ma, mb, P, &status);
            char nm[128]; snprintf(nm, sizeof nm, "E3 K-major SW128 M128 N%d K64", N);
            report(nm, o, N, D, 128, N, status, false);
            cudaFree(c.Ad); cudaFree(c.Bd);
        }
    }

    // ===================== E4: MN-major operands (wgrad): A = dY[pix][64 oc] SW128, B = X[pix][32 c] SW64 ====
    // D[oc (M=64), c (N=32)] = sum_pix dY[pix, oc] * X[pix + shift, c]
    {
        const int PIX = 256, OC = 64, CC = 32;
        fill(c.Ah, &c.Ad, PIX * OC, 7); fill(c.Bh, &c.Bd, PIX * CC, 8);
        CUtensorMap ma = make_map(c.Ad, OC, PIX, OC * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);    // 64 pixels x 128 B
        CUtensorMap mb = make_map(c.Bd, CC, PIX, CC * 2, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);    // 128 pixels x 64 B
        auto expect = [&](int shift, int KP) {
            std::vector<float> D(OC * CC);
            for (int m = 0; m < OC; ++m) for (int n = 0; n < CC; ++n) {
                float s = 0; for (int p = 0; p < KP; ++p) s += c.Ah[p * OC + m] * c.Bh[(p + shift) * CC + n];
                D[m * CC + n] = s;
            }
            return D;
        };
        ProbeParams P{}; P.a_bytes = 64 * 128; P.b_bytes = 128 * 64; P.idesc = make_idesc(64, CC, 1, 1); P.ncols = CC;
        const int KP = 64; P.n_ops = KP / 16;
        // candidate (LBO, SBO) encodings for the MN-major descriptors
        struct Cand { uint32_t a_lbo, a_sbo, b_lbo, b_sbo; const char* tag; };
        Cand cands[] = {
            {8192, 1024, 8192, 512, "A(LBO=tile,SBO=1024) B(LBO=tile,SBO=512)"},
            {1024, 8192, 512, 8192, "A(LBO=1024,SBO=tile) B(LBO=512,SBO=tile)"},
            {16, 1024, 16, 512, "A(LBO=16,SBO=1024) B(LBO=16,SBO=512)"},
        };
        for (auto& cd : cands) {
            for (int shift : {0, 1, 5, 17}) {
                for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                    if (shift % 8 == 0 && bo_mode) continue;
                    int bo = bo_mode ? ((shift * 64) >> 7) & 7 : 0;
                    P.a_desc = make_desc(cd.a_lbo, cd.a_sbo, 2, 0);
                    P.b_desc = make_desc(cd.b_lbo, cd.b_sbo, 4, bo);
                    for (int i = 0; i < P.n_ops; ++i) { P.ops[i].a_off = i * 16 * 128; P.ops[i].b_off = shift * 64 + i * 16 * 64; }
                    auto o = run(c, ma, mb, P, &status);
                    char nm[160]; snprintf(nm, sizeof nm, "E4 MN-major M64 N32 %s k-shift %d bo=%d", cd.tag, shift, bo);
                    report(nm, o, CC, expect(shift, KP), OC, CC, status, true);
                }
            }
        }
        cudaFree(c.Ad); cudaFree(c.Bd);
    }
    printf("probe done\n");
    return 0;
}
