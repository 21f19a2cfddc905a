// Optimizer_Adam.update_params (src/optimizers/Adam.py:86-100) for one element, shared by the single-GPU Adam kernel
// (nm_elementwise_f32.cu) and the fused peer all-reduce + Adam kernel (nm_dp.cu) so the two cannot drift apart.
#pragma once
#include "nm_common.cuh"

struct AdamHP { float lr, b1, b2, eps, bc1, bc2; };

__device__ __forceinline__ AdamHP adam_hp_from(const nm_adam_state* ds) {
    return AdamHP{ds->lr, ds->beta1, ds->beta2, ds->eps, ds->bc1, ds->bc2};
}

// n_apply > 1: the reference registers every trainable layer twice (Model.py:56-59, :82-83), so each step applies the
// update twice with the same gradient and the same t; done here in registers.
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHP& h, int n_apply) {
    for (int a = 0; a < n_apply; ++a) {
        m = h.b1 * m + (1.f - h.b1) * g;
        v = h.b2 * v + (1.f - h.b2) * g * g;
        float mh = m / h.bc1, vh = v / h.bc2;
        p -= h.lr * mh / (sqrtf(vh) + h.eps);
    }
}
