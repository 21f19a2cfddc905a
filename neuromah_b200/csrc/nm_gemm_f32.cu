// FP32-exact contractions on CUDA cores: one register-tiled GEMM kernel with
// pluggable operand loaders covers
//   conv fprop / dgrad  (implicit GEMM, im2col gathered on the fly -- never materialised)
//   conv wgrad          (split-K over N*OH*OW, fixed-order reduction -> deterministic)
//   Dense fprop / dgrad / wgrad (+bias, +ReLU, +1/B scale, +L1/L2)
// Layouts are the reference's: NCHW activations, OIHW weights, (in,out) Dense weights.
//
// GEMM view  C[M,N] = sum_k A[m,k] * B[k,n]:
//   conv fprop : M = N_img*OH*OW, N = OC,        K = C*kH*kW   (Conv2D.cpp:113-134 intent)
//   conv dgrad : M = N_img*H*W,   N = C,         K = OC*kH*kW  (Conv2D.cpp:218-222 + col2im)
//   conv wgrad : M = OC,          N = C*kH*kW(+1 ones column = bias grad), K = N_img*OH*OW
//   dense fprop: M = B, N = out, K = in;  dgrad: M = B, N = in, K = out;
//   dense wgrad: M = in (+1 ones row = bias grad), N = out, K = B
#include "nm_common.cuh"

namespace {

constexpr int BK = 16;
constexpr int NT = 256;

// ---------------------------------------------------------------------------
// Operand loaders.  Each thread loads CNT elements per tile.  For "fast-MN"
// loaders the thread's m (or n) is constant and k steps; for "fast-K" loaders
// the thread's k is constant within a tile and m (or n) steps.
//   St prep(int c)                       state for the thread-constant coordinate
//   load<CNT>(st, cvalid, vfirst, vstep, vend, out)   out[j] for v = vfirst + j*vstep
// ---------------------------------------------------------------------------

// A[m,k] = src[n_img, c, y + sgn*p + oy, x + sgn*q + ox], m=(n_img,y,x), k=(c,p,q); zero outside.
struct ConvGatherA {
    static constexpr bool kFastMN = true;
    const float* src; int Cs, IH, IW, Hd, Wd, kH, kW, oy, ox, sgn;
    struct St { size_t base; int y, x; };
    __device__ St prep(int m) const {
        int hw = Hd * Wd; int n = m / hw; int r = m - n * hw; int y = r / Wd;
        return St{(size_t)n * Cs * IH * IW, y, r - y * Wd};
    }
    template <int CNT>
    __device__ void load(const St& st, bool cvalid, int kfirst, int kstep, int kend, float* out) const {
        int khw = kH * kW; int c = kfirst / khw; int r = kfirst - c * khw; int p = r / kW; int q = r - p * kW;
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            float v = 0.f;
            if (cvalid && kfirst + j * kstep < kend) {
                int yy = st.y + sgn * p + oy, xx = st.x + sgn * q + ox;
                if ((unsigned)yy < (unsigned)IH && (unsigned)xx < (unsigned)IW)
                    v = __ldg(src + st.base + ((size_t)c * IH + yy) * IW + xx);
            }
            out[j] = v;
            q += kstep;
            while (q >= kW) { q -= kW; ++p; }
            while (p >= kH) { p -= kH; ++c; }
        }
    }
};

// B[k,n] = w[koff(k) + n*sn];  koff(k) = k (fprop) or (k/khw)*cmul + k%khw (dgrad).
struct ConvWeightB {
    static constexpr bool kFastMN = false;
    const float* w; int sn, khw, cmul;
    struct St { int koff; };
    __device__ St prep(int k) const {
        if (cmul == 0) return St{k};
        int o = k / khw; return St{o * cmul + (k - o * khw)};
    }
    template <int CNT>
    __device__ void load(const St& st, bool cvalid, int nfirst, int nstep, int nend, float* out) const {
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            int n = nfirst + j * nstep;
            out[j] = (cvalid && n < nend) ? __ldg(w + st.koff + (size_t)n * sn) : 0.f;
        }
    }
};

// Generic strided operand: v[c*sc + v*sv]; optional "ones" index on the varying/const axes.
template <bool FAST_MN>
struct Strided {
    static constexpr bool kFastMN = FAST_MN;
    const float* p; size_t s_mn, s_k; int ones_mn;    // element (mn,k) = p[mn*s_mn + k*s_k]; mn==ones_mn -> 1
    struct St { size_t off; bool one; };
    __device__ St prep(int c) const {
        if (FAST_MN) return St{(size_t)c * s_mn, c == ones_mn};
        return St{(size_t)c * s_k, false};
    }
    template <int CNT>
    __device__ void load(const St& st, bool cvalid, int vfirst, int vstep, int vend, float* out) const {
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            int v = vfirst + j * vstep;
            float r = 0.f;
            if (cvalid && v < vend) {
                if (FAST_MN) r = st.one ? 1.f : __ldg(p + st.off + (size_t)v * s_k);
                else         r = (v == ones_mn) ? 1.f : __ldg(p + st.off + (size_t)v * s_mn);
            }
            out[j] = r;
        }
    }
};

// conv wgrad A[m=oc, k=pix] = dy[(n_img*OC + oc)*OHW + r]
struct WgradDyA {
    static constexpr bool kFastMN = false;
    const float* dy; int OC, OHW;
    struct St { size_t base; };
    __device__ St prep(int k) const { int n = k / OHW; return St{(size_t)n * OC * OHW + (k - n * OHW)}; }
    template <int CNT>
    __device__ void load(const St& st, bool cvalid, int mfirst, int mstep, int mend, float* out) const {
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            int m = mfirst + j * mstep;
            out[j] = (cvalid && m < mend) ? __ldg(dy + st.base + (size_t)m * OHW) : 0.f;
        }
    }
};

// conv wgrad B[k=pix, n=(c,p,q)] = x[n_img, c, y+p-pt, x+q-pl];  n == ones_col -> 1 (bias grad)
struct WgradXB {
    static constexpr bool kFastMN = false;
    const float* x; int C, H, W, OH, OW, kH, kW, pt, pl, ones_col;
    struct St { size_t base; int y, x; };
    __device__ St prep(int k) const {
        int hw = OH * OW; int n = k / hw; int r = k - n * hw; int y = r / OW;
        return St{(size_t)n * C * H * W, y, r - y * OW};
    }
    template <int CNT>
    __device__ void load(const St& st, bool cvalid, int nfirst, int nstep, int nend, float* out) const {
        int khw = kH * kW; int c = nfirst / khw; int r = nfirst - c * khw; int p = r / kW; int q = r - p * kW;
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            int n = nfirst + j * nstep;
            float v = 0.f;
            if (cvalid && n < nend) {
                if (n == ones_col) v = 1.f;
                else {
                    int yy = st.y + p - pt, xx = st.x + q - pl;
                    if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)
                        v = __ldg(x + st.base + ((size_t)c * H + yy) * W + xx);
                }
            }
            out[j] = v;
            q += nstep;
            while (q >= kW) { q -= kW; ++p; }
            while (p >= kH) { p -= kH; ++c; }
        }
    }
};

// ---------------------------------------------------------------------------
// Epilogues
// ---------------------------------------------------------------------------
struct EpiNCHW {            // y[(n_img*Cout + n)*HW + r] = act(acc + bias[n]),  m = n_img*HW + r
    float* y; const float* bias; int Cout, HW, relu;
    __device__ void store(int m, int n, float acc, int) const {
        int img = m / HW; int r = m - img * HW;
        float v = acc + (bias ? __ldg(bias + n) : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        y[((size_t)img * Cout + n) * HW + r] = v;
    }
};
struct EpiRowMajor {        // y[m*ld + n] = act(acc + bias[n])
    float* y; const float* bias; int ld, relu;
    __device__ void store(int m, int n, float acc, int) const {
        float v = acc + (bias ? __ldg(bias + n) : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        y[(size_t)m * ld + n] = v;
    }
};
struct EpiSplitK {          // ws[(z*M + m)*N + n] = acc
    float* ws; int M, N;
    __device__ void store(int m, int n, float acc, int z) const {
        ws[((size_t)z * M + m) * N + n] = acc;
    }
};

// ---------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN, class AL, class BL, class EP>
__global__ void __launch_bounds__(NT)
gemm_f32_kernel(AL al, BL bl, EP ep, int M, int N, int K, int k_per_split) {
    static_assert((BM / TM) * (BN / TN) == NT, "thread tile mismatch");
    constexpr int ACNT = BM * BK / NT;
    constexpr int BCNT = BN * BK / NT;
    static_assert(ACNT >= 1 && BCNT >= 1, "tile too small");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);

    // A mapping
    const int a_c  = AL::kFastMN ? tid % BM : tid % BK;          // thread-constant coord in tile
    const int a_v0 = AL::kFastMN ? tid / BM : tid / BK;          // first varying coord
    constexpr int a_vs = AL::kFastMN ? NT / BM : NT / BK;        // varying step
    // B mapping
    const int b_c  = BL::kFastMN ? tid % BN : tid % BK;
    const int b_v0 = BL::kFastMN ? tid / BN : tid / BK;
    constexpr int b_vs = BL::kFastMN ? NT / BN : NT / BK;

    typename AL::St ast; typename BL::St bst;
    if (AL::kFastMN) ast = al.prep(m0 + a_c);
    if (BL::kFastMN) bst = bl.prep(n0 + b_c);

    float ra[ACNT], rb[BCNT];
    auto load_tile = [&](int k0) {
        if (AL::kFastMN) {
            al.template load<ACNT>(ast, m0 + a_c < M, k0 + a_v0, a_vs, kend, ra);
        } else {
            ast = al.prep(k0 + a_c);
            al.template load<ACNT>(ast, k0 + a_c < kend, m0 + a_v0, a_vs, M, ra);
        }
        if (BL::kFastMN) {
            bl.template load<BCNT>(bst, n0 + b_c < N, k0 + b_v0, b_vs, kend, rb);
        } else {
            bst = bl.prep(k0 + b_c);
            bl.template load<BCNT>(bst, k0 + b_c < kend, n0 + b_v0, b_vs, N, rb);
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int j = 0; j < ACNT; ++j) {
            if (AL::kFastMN) As[a_v0 + j * a_vs][a_c] = ra[j];
            else             As[a_c][a_v0 + j * a_vs] = ra[j];
        }
#pragma unroll
        for (int j = 0; j < BCNT; ++j) {
            if (BL::kFastMN) Bs[b_v0 + j * b_vs][b_c] = rb[j];
            else             Bs[b_c][b_v0 + j * b_vs] = rb[j];
        }
    };

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (kbeg < kend) load_tile(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        store_tile();
        __syncthreads();
        if (k0 + BK < kend) load_tile(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n < N) ep.store(m, n, acc[i][j], blockIdx.z);
        }
    }
}

template <class AL, class BL, class EP>
int launch_gemm(const char* name, AL al, BL bl, EP ep, int M, int N, int K, int splits,
                int k_per_split, cudaStream_t st) {
    if (M <= 0 || N <= 0) return NM_OK;
    // tile shape by problem shape: narrow N (Dense heads, bias columns) and short M (wgrad)
    const bool narrowN = N <= 16, shortM = M <= 64;
    if (narrowN && shortM) {
        dim3 g(nm_cdiv(M, 64), nm_cdiv(N, 16), splits);
        gemm_f32_kernel<64, 16, 4, 1, AL, BL, EP><<<g, NT, 0, st>>>(al, bl, ep, M, N, K, k_per_split);
    } else if (narrowN) {
        dim3 g(nm_cdiv(M, 128), nm_cdiv(N, 16), splits);
        gemm_f32_kernel<128, 16, 8, 1, AL, BL, EP><<<g, NT, 0, st>>>(al, bl, ep, M, N, K, k_per_split);
    } else if (shortM) {
        dim3 g(nm_cdiv(M, 64), nm_cdiv(N, 64), splits);
        gemm_f32_kernel<64, 64, 4, 4, AL, BL, EP><<<g, NT, 0, st>>>(al, bl, ep, M, N, K, k_per_split);
    } else {
        dim3 g(nm_cdiv(M, 128), nm_cdiv(N, 64), splits);
        gemm_f32_kernel<128, 64, 8, 4, AL, BL, EP><<<g, NT, 0, st>>>(al, bl, ep, M, N, K, k_per_split);
    }
    return nm::launched(name);
}

// split-K plan shared by the workspace query and the launch
struct SplitPlan { int splits, k_per_split; };
SplitPlan plan_splitk(int M, int N, long long K) {
    int bm = M <= 64 ? 64 : 128, bn = N <= 16 ? 16 : 64;
    long long blocks = (long long)nm_cdiv(M, bm) * nm_cdiv(N, bn);
    long long want = (4LL * 148 + blocks - 1) / blocks;           // ~4 waves of 148 SMs
    long long max_splits = (K + 8 * BK - 1) / (8 * BK);            // >= 8 k-tiles per split
    long long s = want < 1 ? 1 : (want > max_splits ? max_splits : want);
    if (s < 1) s = 1;
    long long kps = ((K + s - 1) / s + BK - 1) / BK * BK;
    s = (K + kps - 1) / kps;
    return SplitPlan{(int)s, (int)kps};
}

// out[i] = scale * sum_z ws[z][i]  (+ regularisation), fixed z order => deterministic
__global__ void reduce_conv_wgrad_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                         float* __restrict__ dbias, int splits, int OC, int Kc, int ncols) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= OC * ncols) return;
    int oc = i / ncols, j = i - oc * ncols;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(size_t)z * OC * ncols + i];
    if (j < Kc) dw[(size_t)oc * Kc + j] = s;
    else if (dbias) dbias[oc] = s;
}

__global__ void reduce_dense_wgrad_kernel(const float* __restrict__ ws, const float* __restrict__ w,
                                          float* __restrict__ dw, float* __restrict__ db,
                                          int splits, int n_in, int n_out, float scale, float l1, float l2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int total = (n_in + 1) * n_out;
    if (i >= total) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[(size_t)z * total + i];
    s *= scale;
    if (i < n_in * n_out) {
        if (l1 > 0.f) { float wv = w[i]; s += l1 * (wv > 0.f ? 1.f : (wv < 0.f ? -1.f : 0.f)); }
        if (l2 > 0.f) s += 2.f * l2 * w[i];
        dw[i] = s;
    } else if (db) {
        db[i - n_in * n_out] = s;
    }
}

bool conv_dims_ok(int N, int C, int H, int W, int OC, int kH, int kW, int pt, int pb, int pl, int pr) {
    return N > 0 && C > 0 && H > 0 && W > 0 && OC > 0 && kH > 0 && kW > 0 && pt >= 0 && pb >= 0 &&
           pl >= 0 && pr >= 0 && H + pt + pb - kH + 1 > 0 && W + pl + pr - kW + 1 > 0;
}

}  // namespace

extern "C" {

int nm_conv2d_fprop_f32(const float* x, const float* w, const float* bias, float* y,
                        int N, int C, int H, int W, int OC, int kH, int kW,
                        int pad_t, int pad_b, int pad_l, int pad_r, int relu, void* stream) {
    NM_REQUIRE(x && w && y, "null pointer");
    NM_REQUIRE(conv_dims_ok(N, C, H, W, OC, kH, kW, pad_t, pad_b, pad_l, pad_r), "bad conv dimensions");
    const int OH = H + pad_t + pad_b - kH + 1, OW = W + pad_l + pad_r - kW + 1;
    NM_REQUIRE((long long)N * OH * OW < (1LL << 31), "N*OH*OW exceeds int32");
    ConvGatherA al{x, C, H, W, OH, OW, kH, kW, -pad_t, -pad_l, +1};
    ConvWeightB bl{w, C * kH * kW, kH * kW, 0};
    EpiNCHW ep{y, bias, OC, OH * OW, relu};
    int K = C * kH * kW;
    return launch_gemm("conv2d_fprop_f32", al, bl, ep, N * OH * OW, OC, K, 1, (K + BK - 1) / BK * BK, nm::S(stream));
}

int nm_conv2d_dgrad_f32(const float* dy, const float* w, float* dx,
                        int N, int C, int H, int W, int OC, int kH, int kW,
                        int pad_t, int pad_b, int pad_l, int pad_r, void* stream) {
    NM_REQUIRE(dy && w && dx, "null pointer");
    NM_REQUIRE(conv_dims_ok(N, C, H, W, OC, kH, kW, pad_t, pad_b, pad_l, pad_r), "bad conv dimensions");
    const int OH = H + pad_t + pad_b - kH + 1, OW = W + pad_l + pad_r - kW + 1;
    NM_REQUIRE((long long)N * H * W < (1LL << 31), "N*H*W exceeds int32");
    // dx[n,c,y,x] = sum_{oc,p,q} dy[n,oc,y-p+pt,x-q+pl] * w[oc,c,p,q]
    ConvGatherA al{dy, OC, OH, OW, H, W, kH, kW, pad_t, pad_l, -1};
    ConvWeightB bl{w, kH * kW, kH * kW, C * kH * kW};
    EpiNCHW ep{dx, nullptr, C, H * W, 0};
    int K = OC * kH * kW;
    return launch_gemm("conv2d_dgrad_f32", al, bl, ep, N * H * W, C, K, 1, (K + BK - 1) / BK * BK, nm::S(stream));
}

size_t nm_conv2d_wgrad_f32_workspace(int N, int C, int H, int W, int OC, int kH, int kW,
                                     int pad_t, int pad_b, int pad_l, int pad_r) {
    if (!conv_dims_ok(N, C, H, W, OC, kH, kW, pad_t, pad_b, pad_l, pad_r)) return 0;
    const int OH = H + pad_t + pad_b - kH + 1, OW = W + pad_l + pad_r - kW + 1;
    int ncols = C * kH * kW + 1;
    SplitPlan p = plan_splitk(OC, ncols, (long long)N * OH * OW);
    return (size_t)p.splits * OC * ncols * sizeof(float);
}

int nm_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* dbias,
                        void* workspace, size_t workspace_bytes,
                        int N, int C, int H, int W, int OC, int kH, int kW,
                        int pad_t, int pad_b, int pad_l, int pad_r, void* stream) {
    NM_REQUIRE(x && dy && dw && workspace, "null pointer");
    NM_REQUIRE(conv_dims_ok(N, C, H, W, OC, kH, kW, pad_t, pad_b, pad_l, pad_r), "bad conv dimensions");
    const int OH = H + pad_t + pad_b - kH + 1, OW = W + pad_l + pad_r - kW + 1;
    NM_REQUIRE((long long)N * OH * OW < (1LL << 31), "N*OH*OW exceeds int32");
    const int Kc = C * kH * kW, ncols = Kc + 1;
    SplitPlan p = plan_splitk(OC, ncols, (long long)N * OH * OW);
    NM_REQUIRE(workspace_bytes >= (size_t)p.splits * OC * ncols * sizeof(float), "workspace too small");
    WgradDyA al{dy, OC, OH * OW};
    WgradXB bl{x, C, H, W, OH, OW, kH, kW, pad_t, pad_l, Kc};
    EpiSplitK ep{(float*)workspace, OC, ncols};
    int rc = launch_gemm("conv2d_wgrad_f32", al, bl, ep, OC, ncols, N * OH * OW, p.splits, p.k_per_split, nm::S(stream));
    if (rc) return rc;
    int total = OC * ncols;
    reduce_conv_wgrad_kernel<<<nm_cdiv(total, 256), 256, 0, nm::S(stream)>>>((const float*)workspace, dw, dbias,
                                                                            p.splits, OC, Kc, ncols);
    return nm::launched("reduce_conv_wgrad");
}

int nm_dense_fprop_f32(const float* x, const float* w, const float* b, float* y,
                       int B, int n_in, int n_out, int relu, void* stream) {
    NM_REQUIRE(x && w && y, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0, "bad dimensions");
    Strided<false> al{x, (size_t)n_in, 1, -1};       // A[m,k] = x[m*in + k]  (k contiguous)
    Strided<true>  bl{w, 1, (size_t)n_out, -1};      // B[k,n] = w[k*out + n] (n contiguous)
    EpiRowMajor ep{y, b, n_out, relu};
    return launch_gemm("dense_fprop_f32", al, bl, ep, B, n_out, n_in, 1, (n_in + BK - 1) / BK * BK, nm::S(stream));
}

int nm_dense_dgrad_f32(const float* dy, const float* w, float* dx,
                       int B, int n_in, int n_out, void* stream) {
    NM_REQUIRE(dy && w && dx, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0, "bad dimensions");
    Strided<false> al{dy, (size_t)n_out, 1, -1};     // A[m,k] = dy[m*out + k]
    Strided<false> bl{w, (size_t)n_out, 1, -1};      // B[k,n] = w[n*out + k]  (k contiguous)
    EpiRowMajor ep{dx, nullptr, n_in, 0};
    return launch_gemm("dense_dgrad_f32", al, bl, ep, B, n_in, n_out, 1, (n_out + BK - 1) / BK * BK, nm::S(stream));
}

size_t nm_dense_wgrad_f32_workspace(int B, int n_in, int n_out) {
    if (B <= 0 || n_in <= 0 || n_out <= 0) return 0;
    SplitPlan p = plan_splitk(n_in + 1, n_out, B);
    return (size_t)p.splits * (n_in + 1) * n_out * sizeof(float);
}

int nm_dense_wgrad_f32(const float* x, const float* dy, const float* w,
                       float* dw, float* db, void* workspace, size_t workspace_bytes,
                       int B, int n_in, int n_out, float scale, float l1, float l2, void* stream) {
    NM_REQUIRE(x && dy && dw && workspace, "null pointer");
    NM_REQUIRE(B > 0 && n_in > 0 && n_out > 0, "bad dimensions");
    NM_REQUIRE(!(l1 > 0.f || l2 > 0.f) || w, "regularisation needs w");
    SplitPlan p = plan_splitk(n_in + 1, n_out, B);
    NM_REQUIRE(workspace_bytes >= (size_t)p.splits * (n_in + 1) * n_out * sizeof(float), "workspace too small");
    Strided<true> al{x, 1, (size_t)n_in, n_in};      // A[m,k] = x[k*in + m]; row m==in is all ones
    Strided<true> bl{dy, 1, (size_t)n_out, -1};      // B[k,n] = dy[k*out + n]
    EpiSplitK ep{(float*)workspace, n_in + 1, n_out};
    int rc = launch_gemm("dense_wgrad_f32", al, bl, ep, n_in + 1, n_out, B, p.splits, p.k_per_split, nm::S(stream));
    if (rc) return rc;
    int total = (n_in + 1) * n_out;
    reduce_dense_wgrad_kernel<<<nm_cdiv(total, 256), 256, 0, nm::S(stream)>>>((const float*)workspace, w, dw, db,
                                                                             p.splits, n_in, n_out, scale, l1, l2);
    return nm::launched("reduce_dense_wgrad");
}

}  // extern "C"
