// FP32 kernels of the other output activations / losses a Model.train step can end in:
// Activation_Sigmoid, Loss_MeanSquaredError, Loss_MeanAbsoluteError, Loss_BinaryCrossentropy.
// All are HBM-bound streams: coalesced grid-stride loops over a multiple of the SM count; the per-sample
// losses are reduced by one warp per sample with a fixed shuffle tree (deterministic).
#include "nm_common.cuh"

namespace {

constexpr int kSMs = 148;

inline int ew_blocks(size_t n, int threads) {
    size_t want = (n + (size_t)threads - 1) / (size_t)threads;
    size_t cap = (size_t)kSMs * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// ------------------------------------------------------------- sigmoid ----
__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = 1.f / (1.f + expf(-x[i]));           // Sigmoid.py:7
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ out,
                                   float* __restrict__ dx, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float o = out[i];
        dx[i] = dy[i] * (1.f - o) * o;                                     // Sigmoid.py:10
    }
}

// -------------------------------------------------------------- losses ----
// clip(p, 1e-7, 1 - 1e-7) and 1 - clip(p): the complement is clipped on its own side, so that it keeps its
// precision near p = 1 (1 - 1e-7 is not a float32 number; 1 - p is exact there)
__device__ __forceinline__ float clip_p(float p) { return fminf(fmaxf(p, 1e-7f), 1.f - 1e-7f); }
__device__ __forceinline__ float clip_q(float p) { return fminf(fmaxf(1.f - p, 1e-7f), 1.f - 1e-7f); }

__device__ __forceinline__ float loss_term(int kind, float p, float t) {
    if (kind == NM_LOSS_MSE) { const float d = t - p; return d * d; }                          // MSE.py:13
    if (kind == NM_LOSS_MAE) return fabsf(t - p);                                              // MAE.py:13
    return -(t * logf(clip_p(p)) + (1.f - t) * logf(clip_q(p)));                               // binary_crossentropy.py:9-13
}

__device__ __forceinline__ float loss_grad(int kind, float p, float t) {
    if (kind == NM_LOSS_MSE) return -2.f * (t - p);                                            // MSE.py:23
    if (kind == NM_LOSS_MAE) { const float d = t - p; return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }   // MAE.py:23 (np.sign)
    return -(t / clip_p(p) - (1.f - t) / clip_q(p));                                           // binary_crossentropy.py:25-28
}

// one warp per sample: mean of the per-element terms over the D outputs of the sample
__global__ void loss_fwd_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ target,
                                float* __restrict__ sample_loss, int B, int D) {
    const int lane = threadIdx.x & 31;
    const int warps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int r = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < B; r += warps) {      // warp-uniform
        const float* p = pred + (size_t)r * D;
        const float* t = target + (size_t)r * D;
        float s = 0.f;
        for (int j = lane; j < D; j += 32) s += loss_term(kind, p[j], t[j]);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sample_loss[r] = s / (float)D;
    }
}

__global__ void loss_bwd_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ target,
                                float* __restrict__ dx, size_t n, float norm) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dx[i] = loss_grad(kind, pred[i], target[i]) / norm;
}

}  // namespace

extern "C" {

int nm_sigmoid_fwd_f32(const float* x, float* y, size_t n, void* stream) {
    NM_REQUIRE(x && y, "null pointer");
    if (!n) return NM_OK;
    sigmoid_fwd_kernel<<<ew_blocks(n, 256), 256, 0, nm::S(stream)>>>(x, y, n);
    return nm::launched("sigmoid_fwd_f32");
}

int nm_sigmoid_bwd_f32(const float* dy, const float* out, float* dx, size_t n, void* stream) {
    NM_REQUIRE(dy && out && dx, "null pointer");
    if (!n) return NM_OK;
    sigmoid_bwd_kernel<<<ew_blocks(n, 256), 256, 0, nm::S(stream)>>>(dy, out, dx, n);
    return nm::launched("sigmoid_bwd_f32");
}

int nm_loss_fwd_f32(int kind, const float* pred, const float* target, float* sample_loss, int B, int D, void* stream) {
    NM_REQUIRE(kind == NM_LOSS_MSE || kind == NM_LOSS_MAE || kind == NM_LOSS_BCE, "unknown loss kind");
    NM_REQUIRE(pred && target && sample_loss, "null pointer");
    NM_REQUIRE(B > 0 && D > 0, "bad dimensions");
    loss_fwd_kernel<<<ew_blocks((size_t)B * 32, 256), 256, 0, nm::S(stream)>>>(kind, pred, target, sample_loss, B, D);
    return nm::launched("loss_fwd_f32");
}

int nm_loss_bwd_f32(int kind, const float* pred, const float* target, float* dx, int B, int D, void* stream) {
    NM_REQUIRE(kind == NM_LOSS_MSE || kind == NM_LOSS_MAE || kind == NM_LOSS_BCE, "unknown loss kind");
    NM_REQUIRE(pred && target && dx, "null pointer");
    NM_REQUIRE(B > 0 && D > 0, "bad dimensions");
    const size_t n = (size_t)B * D;
    loss_bwd_kernel<<<ew_blocks(n, 256), 256, 0, nm::S(stream)>>>(kind, pred, target, dx, n, (float)D);
    return nm::launched("loss_bwd_f32");
}

}  // extern "C"
