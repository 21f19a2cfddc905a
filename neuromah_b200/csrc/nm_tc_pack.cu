// One launch that refreshes every packed operand copy of the CNN plan (conv1 Wext, conv2 fprop/dgrad weights, Dense
// packed + hi/lo weights) from the FP32 master weights after an optimizer step: three tiny kernels on the critical path
// of every training step become one.
#include "nm_common.cuh"
#include "nm_tc_pack.cuh"

namespace {

struct PackCnnArgs {
    const float* w1; __nv_bfloat16* wext;                                  // conv1 [32][1][3][3]
    const float* w2; __nv_bfloat16* wf; __nv_bfloat16* wd; int OC2, C2;    // conv2 OIHW
    const float* w3; float* wp; __nv_bfloat16* whl; int K, n_out, HW, C3;  // Dense (c,h,w) rows
    int blocks1, blocks2;                                                  // block ranges of the first two parts
};

__global__ void __launch_bounds__(256) pack_cnn_weights_kernel(PackCnnArgs a) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    if (b < a.blocks1) {
        const int i = b * 256 + threadIdx.x;
        if (i < 128 * 16) tcpack::conv1_elem(i, a.w1, a.wext);
    } else if (b < a.blocks1 + a.blocks2) {
        const int i = (b - a.blocks1) * 256 + threadIdx.x;
        if (i < a.OC2 * a.C2 * 9) tcpack::conv3x3_elem(i, a.w2, a.wf, a.wd, a.OC2, a.C2);
    } else {
        const int i = (b - a.blocks1 - a.blocks2) * 256 + threadIdx.x;
        if (i < a.K * 16) tcpack::dense_elem(i, a.w3, a.wp, a.whl, a.K, a.n_out, a.HW, a.C3);
    }
}

}  // namespace

extern "C" int nm_tc_pack_cnn_weights_bf16(const float* w1, void* w1_ext, const float* w2, void* w2_fprop, void* w2_dgrad, int OC2, int C2,
                                           const float* w3, float* w3_packed, void* w3_hilo, int K, int n_out, int HW, int C3,
                                           void* stream) {
    NM_REQUIRE(w1 && w1_ext && w2 && (w2_fprop || w2_dgrad) && w3 && (w3_packed || w3_hilo), "null pointer");
    NM_REQUIRE(OC2 > 0 && C2 > 0 && K > 0 && n_out > 0 && n_out <= 16 && HW * C3 == K, "bad dimensions");
    PackCnnArgs a{w1, static_cast<__nv_bfloat16*>(w1_ext), w2, static_cast<__nv_bfloat16*>(w2_fprop), static_cast<__nv_bfloat16*>(w2_dgrad),
                  OC2, C2, w3, w3_packed, static_cast<__nv_bfloat16*>(w3_hilo), K, n_out, HW, C3,
                  (int)nm_cdiv(128 * 16, 256), (int)nm_cdiv((size_t)OC2 * C2 * 9, 256)};
    const unsigned blocks = a.blocks1 + a.blocks2 + nm_cdiv((size_t)K * 16, 256);
    return nm::launch_pdl("pack_cnn_weights", pack_cnn_weights_kernel, dim3(blocks), dim3(256), 0, nm::S(stream), a);
}
