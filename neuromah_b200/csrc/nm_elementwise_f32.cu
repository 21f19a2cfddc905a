// FP32 memory-bound kernels of the hot path: ReLU, max-pool, softmax, cross-entropy,
// Adam / SGD over a flat slab.  All are coalesced, 128-bit vectorised where the data
// allows it, and sized as grid-stride loops over a multiple of the SM count.
#include "nm_common.cuh"
#include "nm_adam.cuh"
#include <cfloat>

namespace {

constexpr int kSMs = 148;

inline int ew_blocks(size_t n, int threads, int per_thread = 4) {
    size_t want = (n + (size_t)threads * per_thread - 1) / ((size_t)threads * per_thread);
    size_t cap = (size_t)kSMs * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// ---------------------------------------------------------------- ReLU ----
__global__ void relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4, stride = (size_t)gridDim.x * blockDim.x * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    for (; i < n; i += stride) {
        if (vec && i + 4 <= n) {
            float4 v = *reinterpret_cast<const float4*>(x + i);
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            *reinterpret_cast<float4*>(y + i) = v;
        } else {
            for (size_t j = i; j < n && j < i + 4; ++j) y[j] = fmaxf(x[j], 0.f);
        }
    }
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ ref,
                                float* __restrict__ dx, size_t n) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4, stride = (size_t)gridDim.x * blockDim.x * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(ref) |
                       reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    for (; i < n; i += stride) {
        if (vec && i + 4 <= n) {
            float4 d = *reinterpret_cast<const float4*>(dy + i);
            float4 r = *reinterpret_cast<const float4*>(ref + i);
            d.x = r.x <= 0.f ? 0.f : d.x; d.y = r.y <= 0.f ? 0.f : d.y;
            d.z = r.z <= 0.f ? 0.f : d.z; d.w = r.w <= 0.f ? 0.f : d.w;
            *reinterpret_cast<float4*>(dx + i) = d;
        } else {
            for (size_t j = i; j < n && j < i + 4; ++j) dx[j] = ref[j] <= 0.f ? 0.f : dy[j];
        }
    }
}

// ------------------------------------------------------------ max-pool ----
// One thread per output element; zero padding participates (maxpooling_backend.cpp:59),
// strict '>' keeps the first max of the row-major scan (:104).
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int32_t* __restrict__ idx,
                                   size_t total, int H, int W, int ph, int pw, int sh, int sw,
                                   int pt, int pl, int OH, int OW, int PH, int PW) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int ow = (int)(i % OW); size_t t = i / OW; int oh = (int)(t % OH); size_t nc = t / OH;
        const float* xp = x + nc * (size_t)H * W;
        int h0 = oh * sh, w0 = ow * sw;
        int h1 = min(h0 + ph, PH), w1 = min(w0 + pw, PW);
        float best = -INFINITY; int bi = h0, bj = w0;
        for (int a = h0; a < h1; ++a) {
            int yy = a - pt;
            for (int b = w0; b < w1; ++b) {
                int xx = b - pl;
                float v = ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) ? __ldg(xp + (size_t)yy * W + xx) : 0.f;
                if (v > best) { best = v; bi = a; bj = b; }
            }
        }
        y[i] = best;
        reinterpret_cast<int2*>(idx)[i] = make_int2(bi, bj);
    }
}

// Gather formulation of the reference's scatter-add: each input element sums the
// gradients of the windows whose saved argmax is this element (fixed order, no atomics).
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx,
                                   float* __restrict__ dx, size_t total, int H, int W,
                                   int ph, int pw, int sh, int sw, int pt, int pl, int OH, int OW) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int w = (int)(i % W); size_t t = i / W; int h = (int)(t % H); size_t nc = t / H;
        int hp = h + pt, wp = w + pl;                      // padded coordinates
        int oh_lo = hp - ph + 1; oh_lo = oh_lo <= 0 ? 0 : (oh_lo + sh - 1) / sh;
        int ow_lo = wp - pw + 1; ow_lo = ow_lo <= 0 ? 0 : (ow_lo + sw - 1) / sw;
        int oh_hi = min(hp / sh, OH - 1), ow_hi = min(wp / sw, OW - 1);
        float s = 0.f;
        for (int oh = oh_lo; oh <= oh_hi; ++oh)
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                size_t o = (nc * OH + oh) * (size_t)OW + ow;
                int2 ij = __ldg(reinterpret_cast<const int2*>(idx) + o);
                if (ij.x == hp && ij.y == wp) s += __ldg(dy + o);
            }
        dx[i] = s;
    }
}

// ------------------------------------------------------------- softmax ----
// One warp per row; lanes stride over the classes.
__global__ void softmax_fwd_kernel(const float* __restrict__ z, float* __restrict__ p, int B, int n) {
    int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* zr = z + (size_t)row * n;
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, zr[j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += expf(zr[j] - mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int j = lane; j < n; j += 32) p[(size_t)row * n + j] = expf(zr[j] - mx) / s;
}

__global__ void softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dy,
                                   float* __restrict__ dx, int B, int n) {
    int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* pr = p + (size_t)row * n; const float* dr = dy + (size_t)row * n;
    float dot = 0.f;
    for (int j = lane; j < n; j += 32) dot += pr[j] * dr[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int j = lane; j < n; j += 32) dx[(size_t)row * n + j] = pr[j] * (dr[j] - dot);
}

// per-sample loss / correct flag / fused gradient; one warp per row
__global__ void softmax_ce_kernel(const float* __restrict__ p, const int32_t* __restrict__ labels,
                                  float* __restrict__ dlogits, float* __restrict__ sample_loss,
                                  int32_t* __restrict__ correct, int B, int n, float inv_batch) {
    int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* pr = p + (size_t)row * n;
    int y = labels[row];
    float best = -INFINITY; int bj = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
        float v = pr[j];
        if (v > best) { best = v; bj = j; }              // first max within the lane's stride
        if (dlogits) dlogits[(size_t)row * n + j] = (v - (j == y ? 1.f : 0.f)) * inv_batch;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {                         // np.argmax: lowest index among equal maxima
        float ob = __shfl_xor_sync(0xffffffffu, best, o); int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (lane == 0) {
        float py = (y >= 0 && y < n) ? pr[y] : 1e-7f;
        py = fminf(fmaxf(py, 1e-7f), 1.f - 1e-7f);         // np.clip(p, 1e-7, 1-1e-7)
        if (sample_loss) sample_loss[row] = -logf(py);
        if (correct) correct[row] = (bj == y) ? 1 : 0;
    }
}

// single block, fixed order -> deterministic accumulation into double[2]
__global__ void loss_acc_reduce_kernel(const float* __restrict__ sample_loss, const int32_t* __restrict__ correct,
                                       double* __restrict__ accum, int B) {
    __shared__ double sl[256]; __shared__ double sc[256];
    double l = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { l += (double)sample_loss[i]; c += (double)correct[i]; }
    sl[threadIdx.x] = l; sc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) { sl[threadIdx.x] += sl[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { accum[0] += sl[0]; accum[1] += sc[0]; }
}

__global__ void cce_bwd_kernel(const float* __restrict__ p, const int32_t* __restrict__ labels,
                               float* __restrict__ dx, int B, int n, float inv_batch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * n) return;
    int row = (int)(i / n), j = (int)(i - (size_t)row * n);
    dx[i] = (j == labels[row]) ? (-1.f / p[i]) * inv_batch : 0.f;   // -y/p / samples, y one-hot
}

// ----------------------------------------------------------- optimizers ---
template <bool DEV>
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, AdamHP hp, nm_adam_state* ds, int n_apply, int advance) {
    pdl_trigger();
    pdl_wait();
    if (DEV) hp = adam_hp_from(ds);
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4, stride = (size_t)gridDim.x * blockDim.x * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    for (; i < n; i += stride) {
        if (vec && i + 4 <= n) {
            float4 P = *reinterpret_cast<float4*>(p + i), M = *reinterpret_cast<float4*>(m + i),
                   V = *reinterpret_cast<float4*>(v + i);
            const float4 G = *reinterpret_cast<const float4*>(g + i);
            adam_one(P.x, G.x, M.x, V.x, hp, n_apply); adam_one(P.y, G.y, M.y, V.y, hp, n_apply);
            adam_one(P.z, G.z, M.z, V.z, hp, n_apply); adam_one(P.w, G.w, M.w, V.w, hp, n_apply);
            *reinterpret_cast<float4*>(p + i) = P; *reinterpret_cast<float4*>(m + i) = M;
            *reinterpret_cast<float4*>(v + i) = V;
        } else {
            for (size_t j = i; j < n && j < i + 4; ++j) adam_one(p[j], g[j], m[j], v[j], hp, n_apply);
        }
    }
    if (DEV && advance) {
        // every block read the hyper-parameters before it got here, so the last one to arrive may advance them
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&ds->ticket, 1u) == gridDim.x - 1) {
            ds->ticket = 0u;
            adam_state_advance(ds);
        }
    }
}

__global__ void adam_state_init_kernel(nm_adam_state* s, double lr, double decay, double b1, double b2, float eps,
                                       long long it) {
    s->ticket = 0u; s->reserved = 0u;
    s->base_lr_d = lr; s->decay_d = decay; s->beta1_d = b1; s->beta2_d = b2;
    s->base_lr = (float)lr; s->decay = (float)decay; s->beta1 = (float)b1; s->beta2 = (float)b2; s->eps = eps; s->iterations = it;
    s->beta1_pow = pow(b1, (double)(it + 1)); s->beta2_pow = pow(b2, (double)(it + 1));
    s->bc1 = (float)(1.0 - s->beta1_pow); s->bc2 = (float)(1.0 - s->beta2_pow);
    s->lr = decay != 0.0 ? (float)(lr * (1.0 / (1.0 + decay * (double)it))) : (float)lr;
}

// post_update_params (Adam.py:103-107) + next step's pre_update_params (:41-47), on the device
__global__ void adam_advance_kernel(nm_adam_state* s) {
    pdl_trigger();
    pdl_wait();
    adam_state_advance(s);
}

__global__ void scale_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n, float a, int accumulate) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = accumulate ? y[i] + a * x[i] : a * x[i];
}

// raw uint8 pixels -> float32 with the reference's host-side normalisation `X.astype(np.float32) / 255.0`
// (examples/test_Conv2D_MNIST.py:27): IEEE division, so the result is bit-identical to NumPy's.  4 pixels per thread.
__global__ void u8_to_f32_div_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, size_t n, float denom) {
    const size_t n4 = n / 4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        const uchar4 u = reinterpret_cast<const uchar4*>(x)[i];
        reinterpret_cast<float4*>(y)[i] = make_float4(__fdiv_rn((float)u.x, denom), __fdiv_rn((float)u.y, denom),
                                                      __fdiv_rn((float)u.z, denom), __fdiv_rn((float)u.w, denom));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) y[n4 * 4 + threadIdx.x] = __fdiv_rn((float)x[n4 * 4 + threadIdx.x], denom);
}

// ---- TensorMonitor.log_histogram (src/utils/TensorMonitor.py:93-104: np.histogram(values.flatten(), bins)) on the device ----
// pass 1: min / max of the tensor (order-independent, so the block/atomic combination is deterministic)
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ out) {
    __shared__ unsigned smin[8], smax[8];
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned o = f2ord(x[i]);
        lo = min(lo, o); hi = max(hi, o);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s)); }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = min(lo, smin[w]); hi = max(hi, smax[w]); }
        atomicMin(out, lo); atomicMax(out + 1, hi);
    }
}
__global__ void minmax_finish_kernel(unsigned* out) {                 // ordered keys -> floats, in place
    const float lo = ord2f(out[0]), hi = ord2f(out[1]);
    reinterpret_cast<float*>(out)[0] = lo; reinterpret_cast<float*>(out)[1] = hi;
}
// pass 2: bin counts with NumPy's own index arithmetic for uniform bins (float64: scaled index, then the two corrections
// against the linspace edges), integer atomics -- the counts are exact and independent of the order of summation
__global__ void __launch_bounds__(256)
histogram_kernel(const float* __restrict__ x, size_t n, double first, double last, int bins, unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned hist[];
    for (int b = threadIdx.x; b < bins; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    const double step = (last - first) / bins, norm = bins / (last - first);
    auto edge = [&](int k) { return k == bins ? last : k * step + first; };      // np.linspace(first, last, bins + 1)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        if (!(v >= first && v <= last)) continue;                                  // outside the range (or NaN): not counted
        int k = (int)((v - first) * norm);
        if (k == bins) k = bins - 1;
        if (v < edge(k)) --k;
        else if (k != bins - 1 && v >= edge(k + 1)) ++k;
        atomicAdd(&hist[k], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += blockDim.x)
        if (hist[b]) atomicAdd(&counts[b], (unsigned long long)hist[b]);
}

}  // namespace

extern "C" {

int nm_minmax_f32(const float* x, size_t n, float* out2, void* stream) {
    NM_REQUIRE(x && out2 && n > 0, "null pointer or empty tensor");
    unsigned* o = reinterpret_cast<unsigned*>(out2);
    const unsigned init[2] = {0xFFFFFFFFu, 0u};
    NM_CUDA(cudaMemcpyAsync(o, init, sizeof(init), cudaMemcpyHostToDevice, nm::S(stream)));
    minmax_kernel<<<ew_blocks(n, 256, 8), 256, 0, nm::S(stream)>>>(x, n, o);
    minmax_finish_kernel<<<1, 1, 0, nm::S(stream)>>>(o);
    return nm::launched("minmax_f32", 2);
}

int nm_histogram_f32(const float* x, size_t n, double first, double last, int bins, unsigned long long* counts, void* stream) {
    NM_REQUIRE(x && counts, "null pointer");
    NM_REQUIRE(bins > 0 && bins <= 4096 && last > first, "bins must be in [1, 4096] and the range non-empty");
    NM_CUDA(cudaMemsetAsync(counts, 0, (size_t)bins * sizeof(unsigned long long), nm::S(stream)));
    if (!n) return NM_OK;
    histogram_kernel<<<ew_blocks(n, 256, 8), 256, (size_t)bins * sizeof(unsigned), nm::S(stream)>>>(x, n, first, last, bins, counts);
    return nm::launched("histogram_f32");
}

int nm_u8_to_f32_div(const uint8_t* x, float* y, size_t n, float denom, void* stream) {
    NM_REQUIRE(x && y, "null pointer");
    NM_REQUIRE(denom != 0.f, "denom must be non-zero");
    NM_REQUIRE(((uintptr_t)x & 3) == 0 && ((uintptr_t)y & 15) == 0, "x must be 4-byte and y 16-byte aligned");
    if (!n) return NM_OK;
    u8_to_f32_div_kernel<<<ew_blocks(n, 256, 4), 256, 0, nm::S(stream)>>>(x, y, n, denom);
    return nm::launched("u8_to_f32_div");
}

int nm_relu_fwd_f32(const float* x, float* y, size_t n, void* stream) {
    NM_REQUIRE(x && y, "null pointer");
    if (!n) return NM_OK;
    relu_fwd_kernel<<<ew_blocks(n, 256, 4), 256, 0, nm::S(stream)>>>(x, y, n);
    return nm::launched("relu_fwd_f32");
}

int nm_relu_bwd_f32(const float* dy, const float* ref, float* dx, size_t n, void* stream) {
    NM_REQUIRE(dy && ref && dx, "null pointer");
    if (!n) return NM_OK;
    relu_bwd_kernel<<<ew_blocks(n, 256, 4), 256, 0, nm::S(stream)>>>(dy, ref, dx, n);
    return nm::launched("relu_bwd_f32");
}

int nm_maxpool2d_fwd_f32(const float* x, float* y, int32_t* idx,
                         int N, int C, int H, int W, int ph, int pw, int sh, int sw,
                         int pad_t, int pad_l, int OH, int OW, void* stream) {
    NM_REQUIRE(x && y && idx, "null pointer");
    NM_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && sh > 0 && sw > 0 && OH > 0 && OW > 0 &&
               pad_t >= 0 && pad_l >= 0, "bad pool dimensions");
    // padded extent as the reference computes it (maxpooling_backend.cpp:37-61):
    // 'same' pads up to (OH-1)*s + pool, 'valid' windows always fit inside H x W.
    const int PH = max(H, (OH - 1) * sh + ph), PW = max(W, (OW - 1) * sw + pw);
    size_t total = (size_t)N * C * OH * OW;
    maxpool_fwd_kernel<<<ew_blocks(total, 256, 1), 256, 0, nm::S(stream)>>>(x, y, idx, total, H, W, ph, pw, sh, sw,
                                                                           pad_t, pad_l, OH, OW, PH, PW);
    return nm::launched("maxpool2d_fwd_f32");
}

int nm_maxpool2d_bwd_f32(const float* dy, const int32_t* idx, float* dx,
                         int N, int C, int H, int W, int ph, int pw, int sh, int sw,
                         int pad_t, int pad_l, int OH, int OW, void* stream) {
    NM_REQUIRE(dy && idx && dx, "null pointer");
    NM_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && sh > 0 && sw > 0 && OH > 0 && OW > 0,
               "bad pool dimensions");
    size_t total = (size_t)N * C * H * W;
    maxpool_bwd_kernel<<<ew_blocks(total, 256, 1), 256, 0, nm::S(stream)>>>(dy, idx, dx, total, H, W, ph, pw, sh, sw,
                                                                           pad_t, pad_l, OH, OW);
    return nm::launched("maxpool2d_bwd_f32");
}

int nm_softmax_fwd_f32(const float* logits, float* probs, int B, int n, void* stream) {
    NM_REQUIRE(logits && probs, "null pointer");
    NM_REQUIRE(B > 0 && n > 0, "bad dimensions");
    softmax_fwd_kernel<<<nm_cdiv((size_t)B * 32, 256), 256, 0, nm::S(stream)>>>(logits, probs, B, n);
    return nm::launched("softmax_fwd_f32");
}

int nm_softmax_bwd_f32(const float* probs, const float* dy, float* dx, int B, int n, void* stream) {
    NM_REQUIRE(probs && dy && dx, "null pointer");
    NM_REQUIRE(B > 0 && n > 0, "bad dimensions");
    softmax_bwd_kernel<<<nm_cdiv((size_t)B * 32, 256), 256, 0, nm::S(stream)>>>(probs, dy, dx, B, n);
    return nm::launched("softmax_bwd_f32");
}

int nm_softmax_ce_f32(const float* probs, const int32_t* labels, float* dlogits,
                      float* sample_loss, int32_t* correct, double* accum,
                      int B, int n, float inv_batch, void* stream) {
    NM_REQUIRE(probs && labels, "null pointer");
    NM_REQUIRE(B > 0 && n > 0, "bad dimensions");
    NM_REQUIRE(!accum || (sample_loss && correct), "accum needs sample_loss and correct buffers");
    softmax_ce_kernel<<<nm_cdiv((size_t)B * 32, 256), 256, 0, nm::S(stream)>>>(probs, labels, dlogits, sample_loss,
                                                                              correct, B, n, inv_batch);
    int rc = nm::launched("softmax_ce_f32");
    if (rc || !accum) return rc;
    loss_acc_reduce_kernel<<<1, 256, 0, nm::S(stream)>>>(sample_loss, correct, accum, B);
    return nm::launched("loss_acc_reduce");
}

int nm_cce_bwd_f32(const float* probs, const int32_t* labels, float* dx, int B, int n, float inv_batch, void* stream) {
    NM_REQUIRE(probs && labels && dx, "null pointer");
    NM_REQUIRE(B > 0 && n > 0, "bad dimensions");
    cce_bwd_kernel<<<nm_cdiv((size_t)B * n, 256), 256, 0, nm::S(stream)>>>(probs, labels, dx, B, n, inv_batch);
    return nm::launched("cce_bwd_f32");
}

int nm_adam_step_f32(float* param, const float* grad, float* m, float* v, size_t n,
                     float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                     int n_apply, void* stream) {
    NM_REQUIRE(param && grad && m && v, "null pointer");
    NM_REQUIRE(n_apply >= 1, "n_apply must be >= 1");
    if (!n) return NM_OK;
    AdamHP hp{lr, beta1, beta2, eps, bc1, bc2};
    return nm::launch_pdl("adam_step", adam_kernel<false>, dim3(ew_blocks(n, 256, 4)), dim3(256), 0, nm::S(stream),
                          param, grad, m, v, n, hp, (nm_adam_state*)nullptr, n_apply, 0);
}

int nm_adam_state_init(nm_adam_state* dev_state, double lr, double decay, double beta1,
                       double beta2, float eps, long long iterations, void* stream) {
    NM_REQUIRE(dev_state, "null pointer");
    adam_state_init_kernel<<<1, 1, 0, nm::S(stream)>>>(dev_state, lr, decay, beta1, beta2, eps, iterations);
    return nm::launched("adam_state_init");
}

int nm_adam_step_dev_f32(float* param, const float* grad, float* m, float* v, size_t n,
                         const nm_adam_state* dev_state, int n_apply, void* stream) {
    NM_REQUIRE(param && grad && m && v && dev_state, "null pointer");
    NM_REQUIRE(n_apply >= 1, "n_apply must be >= 1");
    if (!n) return NM_OK;
    return nm::launch_pdl("adam_step_dev", adam_kernel<true>, dim3(ew_blocks(n, 256, 4)), dim3(256), 0, nm::S(stream),
                          param, grad, m, v, n, AdamHP{}, const_cast<nm_adam_state*>(dev_state), n_apply, 0);
}

int nm_adam_step_advance_dev_f32(float* param, const float* grad, float* m, float* v, size_t n,
                                 nm_adam_state* dev_state, int n_apply, void* stream) {
    NM_REQUIRE(param && grad && m && v && dev_state, "null pointer");
    NM_REQUIRE(n_apply >= 1 && n > 0, "n_apply must be >= 1 and n > 0");
    return nm::launch_pdl("adam_step_advance_dev", adam_kernel<true>, dim3(ew_blocks(n, 256, 4)), dim3(256), 0, nm::S(stream),
                          param, grad, m, v, n, AdamHP{}, dev_state, n_apply, 1);
}

int nm_adam_advance_f32(nm_adam_state* dev_state, void* stream) {
    NM_REQUIRE(dev_state, "null pointer");
    return nm::launch_pdl("adam_advance", adam_advance_kernel, dim3(1), dim3(1), 0, nm::S(stream), dev_state);
}

int nm_scale_f32(float* y, const float* x, size_t n, float a, int accumulate, void* stream) {
    NM_REQUIRE(x && y, "null pointer");
    if (!n) return NM_OK;
    scale_kernel<<<ew_blocks(n, 256, 1), 256, 0, nm::S(stream)>>>(y, x, n, a, accumulate);
    return nm::launched("scale_f32");
}

}  // extern "C"
