// Optimizer_Adam.update_params (src/optimizers/Adam.py:86-100) for one element, shared by the single-GPU Adam kernel
// (nm_elementwise_f32.cu) and the fused peer all-reduce + Adam kernel (nm_dp.cu) so the two cannot drift apart.
#pragma once
#include "nm_common.cuh"

struct AdamHP { float lr, b1, b2, eps, bc1, bc2; };

// The update is two divisions and a square root per application; with IEEE-exact versions (~10 instructions each, plus
// slow paths for zero / denormal operands) two applications per parameter made the kernel ALU-bound (2.2 TB/s on an
// all-zero gradient slab, 5.1 TB/s otherwise).  Reciprocals of the bias corrections + MUFU-based division / square root
// (<= 2 ulp each, far inside the 1e-4 parity bar) make it HBM-bound for any data.
__device__ __forceinline__ float nm_sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ AdamHP adam_hp_from(const nm_adam_state* ds) {
    return AdamHP{ds->lr, ds->beta1, ds->beta2, ds->eps, ds->bc1, ds->bc2};
}

// n_apply > 1: the reference registers every trainable layer twice (Model.py:56-59, :82-83), so each step applies the
// update twice with the same gradient and the same t; done here in registers.
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHP& h, int n_apply) {
    const float inv1 = __frcp_rn(h.bc1), inv2 = __frcp_rn(h.bc2);
    for (int a = 0; a < n_apply; ++a) {
        m = h.b1 * m + (1.f - h.b1) * g;
        v = h.b2 * v + (1.f - h.b2) * g * g;
        const float mh = m * inv1, vh = v * inv2;
        p -= h.lr * __fdividef(mh, nm_sqrt_approx(vh) + h.eps);
    }
}

// post_update_params (Adam.py:103-107) + the next step's pre_update_params (:41-47), on the device.  Same formula as the
// host (pow of the iteration count, Adam.py:94-95), not a running product: no drift between the paths.
__device__ __forceinline__ void adam_state_advance(nm_adam_state* s) {
    s->iterations += 1;
    s->beta1_pow = pow(s->beta1_d, (double)(s->iterations + 1));
    s->beta2_pow = pow(s->beta2_d, (double)(s->iterations + 1));
    s->bc1 = (float)(1.0 - s->beta1_pow); s->bc2 = (float)(1.0 - s->beta2_pow);
    if (s->decay_d != 0.0) s->lr = (float)(s->base_lr_d * (1.0 / (1.0 + s->decay_d * (double)s->iterations)));
}
