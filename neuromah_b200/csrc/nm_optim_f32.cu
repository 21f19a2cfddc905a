// The remaining optimizers of the reference as variants of one coalesced, 128-bit vectorised slab kernel
// (SURVEY.md 8f row 3), and Layer_Dropout.  Each is HBM-bound: one pass over (param, grad, state...) per step,
// the update applied n_apply times in registers (Model registers every trainable layer twice: Model.py:56-59, :82-83).
//   Optimizer_SGD      src/optimizers/SGD.py:85-97      12 B/param (20 with momentum)
//   Optimizer_RMSprop  src/optimizers/RMSprop.py:103-109  20 B/param
//   Optimizer_Adagrad  src/optimizers/Adagrad.py (cache += g^2; p -= lr g / (sqrt(cache) + eps))  20 B/param
//   Optimizer_Nadam    src/optimizers/Nadam.py (eps added to the bias corrections as well)  28 B/param
//   Layer_Dropout      src/layers/Dropout.py:15-40: mask = Bernoulli(keep) / keep, drawn with Philox4x32-10 on the device
#include "nm_common.cuh"
#include "nm_adam.cuh"
#include "nm_philox.cuh"      // nm_sqrt_approx: MUFU square root / division keep these kernels HBM-bound (see there)

namespace {

enum { K_SGD = 0, K_SGD_MOM = 1, K_RMSPROP = 2, K_ADAGRAD = 3, K_NADAM = 4 };
struct OptHP { float lr, a, b, eps, c1, c2; };

template <int KIND>
__device__ __forceinline__ void opt_one(float& p, float g, float& s1, float& s2, const OptHP& h, int n_apply) {
    for (int it = 0; it < n_apply; ++it) {
        if (KIND == K_SGD) {
            p += -h.lr * g;                                           // SGD.py:97
        } else if (KIND == K_SGD_MOM) {
            s1 = s1 * h.a + h.lr * g;                                 // SGD.py:92-95 (v = mu v + lr g; p -= v)
            p -= s1;
        } else if (KIND == K_RMSPROP) {
            s1 = h.a * s1 + (1.f - h.a) * g * g;                      // RMSprop.py:103-106
            p -= h.lr * __fdividef(g, nm_sqrt_approx(s1) + h.eps);    // RMSprop.py:109
        } else if (KIND == K_ADAGRAD) {
            s1 += g * g;
            p -= h.lr * __fdividef(g, nm_sqrt_approx(s1) + h.eps);
        } else {
            s1 = h.a * s1 + (1.f - h.a) * g;                          // Nadam: momentums / caches
            s2 = h.b * s2 + (1.f - h.b) * g * g;
            const float i1 = __frcp_rn(h.c1 + h.eps), i2 = __frcp_rn(h.c2 + h.eps);   // c1 = 1 - beta1^(t+1), c2 = 1 - beta2^(t+1)
            const float mh = s1 * i1, vh = s2 * i2;
            const float mn = h.a * mh + (1.f - h.a) * g * i1;
            p -= h.lr * __fdividef(mn, nm_sqrt_approx(vh) + h.eps);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
opt_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s1, float* __restrict__ s2, size_t n,
           OptHP h, int n_apply, int32_t* __restrict__ nonfinite) {
    constexpr bool HAS1 = KIND != K_SGD, HAS2 = KIND == K_NADAM;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | (HAS1 ? reinterpret_cast<uintptr_t>(s1) : 0) |
                       (HAS2 ? reinterpret_cast<uintptr_t>(s2) : 0)) & 15) == 0;
    bool bad = false;
    for (; i < n; i += stride) {
        if (vec && i + 4 <= n) {
            float4 P = *reinterpret_cast<float4*>(p + i);
            const float4 G = *reinterpret_cast<const float4*>(g + i);
            float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
            if (HAS1) A = *reinterpret_cast<float4*>(s1 + i);
            if (HAS2) B = *reinterpret_cast<float4*>(s2 + i);
            bad |= !(isfinite(G.x) && isfinite(G.y) && isfinite(G.z) && isfinite(G.w));
            opt_one<KIND>(P.x, G.x, A.x, B.x, h, n_apply); opt_one<KIND>(P.y, G.y, A.y, B.y, h, n_apply);
            opt_one<KIND>(P.z, G.z, A.z, B.z, h, n_apply); opt_one<KIND>(P.w, G.w, A.w, B.w, h, n_apply);
            *reinterpret_cast<float4*>(p + i) = P;
            if (HAS1) *reinterpret_cast<float4*>(s1 + i) = A;
            if (HAS2) *reinterpret_cast<float4*>(s2 + i) = B;
        } else {
            for (size_t j = i; j < n && j < i + 4; ++j) {
                float a = HAS1 ? s1[j] : 0.f, b = HAS2 ? s2[j] : 0.f, pv = p[j];
                const float gv = g[j];
                bad |= !isfinite(gv);
                opt_one<KIND>(pv, gv, a, b, h, n_apply);
                p[j] = pv;
                if (HAS1) s1[j] = a;
                if (HAS2) s2[j] = b;
            }
        }
    }
    if (nonfinite && bad) *nonfinite = 1;
}

inline unsigned opt_blocks(size_t n) {
    const size_t want = (n + 1023) / 1024;
    return (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
}

template <int KIND>
int launch_opt(const char* name, float* p, const float* g, float* s1, float* s2, size_t n, OptHP h, int n_apply,
               int32_t* nonfinite, void* stream) {
    if (!n) return NM_OK;
    opt_kernel<KIND><<<opt_blocks(n), 256, 0, nm::S(stream)>>>(p, g, s1, s2, n, h, n_apply, nonfinite);
    return nm::launched(name);
}

// y = x * mask, mask = (u < keep) / keep with u uniform in [0,1) from 32 random bits (24 used: exact in fp32)
__global__ void __launch_bounds__(256)
dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ mask, size_t n, float keep,
                   unsigned long long seed, unsigned long long offset) {
    const float inv = 1.f / keep;
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;                 // group of 4 elements
    const size_t stride = (size_t)gridDim.x * blockDim.x, groups = (n + 3) / 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
    for (; q < groups; q += stride) {
        const uint32_t kb = dropout_keep4(q, seed, offset, keep);
        float m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) m[u] = (kb >> u) & 1u ? inv : 0.f;
        const size_t i = q * 4;
        if (vec && i + 4 <= n) {
            const float4 X = *reinterpret_cast<const float4*>(x + i);
            *reinterpret_cast<float4*>(mask + i) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(y + i) = make_float4(X.x * m[0], X.y * m[1], X.z * m[2], X.w * m[3]);
        } else {
            for (int u = 0; u < 4 && i + u < n; ++u) { mask[i + u] = m[u]; y[i + u] = x[i + u] * m[u]; }
        }
    }
}

__global__ void __launch_bounds__(256)
mul_kernel(float* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    for (; i < n; i += stride) {
        if (vec && i + 4 <= n) {
            const float4 A = *reinterpret_cast<const float4*>(a + i), B = *reinterpret_cast<const float4*>(b + i);
            *reinterpret_cast<float4*>(y + i) = make_float4(A.x * B.x, A.y * B.y, A.z * B.z, A.w * B.w);
        } else {
            for (size_t j = i; j < n && j < i + 4; ++j) y[j] = a[j] * b[j];
        }
    }
}

}  // namespace

extern "C" {

int nm_sgd_step_f32(float* param, const float* grad, float* mom, size_t n, float lr, float momentum, int n_apply,
                    int32_t* nonfinite, void* stream) {
    NM_REQUIRE(param && grad, "null pointer");
    NM_REQUIRE(n_apply >= 1, "n_apply must be >= 1");
    NM_REQUIRE(momentum == 0.f || mom, "momentum buffer is NULL");
    const OptHP h{lr, momentum, 0.f, 0.f, 0.f, 0.f};
    if (momentum != 0.f) return launch_opt<K_SGD_MOM>("sgd_step_f32", param, grad, mom, nullptr, n, h, n_apply, nonfinite, stream);
    return launch_opt<K_SGD>("sgd_step_f32", param, grad, nullptr, nullptr, n, h, n_apply, nonfinite, stream);
}

int nm_rmsprop_step_f32(float* param, const float* grad, float* cache, size_t n, float lr, float rho, float eps,
                        int n_apply, int32_t* nonfinite, void* stream) {
    NM_REQUIRE(param && grad && cache, "null pointer");
    NM_REQUIRE(n_apply >= 1, "n_apply must be >= 1");
    return launch_opt<K_RMSPROP>("rmsprop_step_f32", param, grad, cache, nullptr, n, OptHP{lr, rho, 0.f, eps, 0.f, 0.f}, n_apply, nonfinite, stream);
}

int nm_adagrad_step_f32(float* param, const float* grad, float* cache, size_t n, float lr, float eps,
                        int n_apply, int32_t* nonfinite, void* stream) {
    NM_REQUIRE(param && grad && cache, "null pointer");
    NM_REQUIRE(n_apply >= 1, "n_apply must be >= 1");
    return launch_opt<K_ADAGRAD>("adagrad_step_f32", param, grad, cache, nullptr, n, OptHP{lr, 0.f, 0.f, eps, 0.f, 0.f}, n_apply, nonfinite, stream);
}

int nm_nadam_step_f32(float* param, const float* grad, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                      float eps, float bc1, float bc2, int n_apply, void* stream) {
    NM_REQUIRE(param && grad && m && v, "null pointer");
    NM_REQUIRE(n_apply >= 1, "n_apply must be >= 1");
    return launch_opt<K_NADAM>("nadam_step_f32", param, grad, m, v, n, OptHP{lr, beta1, beta2, eps, bc1, bc2}, n_apply, nullptr, stream);
}

int nm_dropout_fwd_f32(const float* x, float* y, float* mask, size_t n, float keep, unsigned long long seed,
                       unsigned long long offset, void* stream) {
    NM_REQUIRE(x && y && mask, "null pointer");
    NM_REQUIRE(keep > 0.f && keep <= 1.f, "keep probability must be in (0, 1]");
    if (!n) return NM_OK;
    dropout_fwd_kernel<<<opt_blocks(n), 256, 0, nm::S(stream)>>>(x, y, mask, n, keep, seed, offset);
    return nm::launched("dropout_fwd_f32");
}

int nm_mul_f32(float* y, const float* a, const float* b, size_t n, void* stream) {
    NM_REQUIRE(y && a && b, "null pointer");
    if (!n) return NM_OK;
    mul_kernel<<<opt_blocks(n), 256, 0, nm::S(stream)>>>(y, a, b, n);
    return nm::launched("mul_f32");
}

}  // extern "C"
