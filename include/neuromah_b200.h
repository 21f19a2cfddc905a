/*
 * neuromah_b200.h -- C ABI of libneuromah_b200.so (sm_100a).
 *
 * This is the drop-in boundary for Neuromah's training hot path.  Every entry
 * point is `extern "C"`, takes plain device/host pointers and sizes, launches
 * asynchronously on the caller's CUDA stream (a `cudaStream_t` passed as
 * `void*`, NULL = legacy default stream) and returns an int status
 * (0 = NM_OK, negative = error; text via nm_last_error()).  No torch types,
 * no exceptions, never exit().
 *
 * Reference interfaces replaced (paths relative to the Neuromah repository):
 *   - pybind11 module _Conv2DBackend_cpp   (src/layers/CPU_kernels/Conv2D.cpp:242-245,
 *     call site src/layers/Conv_wrapepr.py:91)
 *   - pybind11 module _MaxPooling2DBackend_cpp (CPU_kernels/maxpooling_backend.cpp:240-247,
 *     call sites src/layers/MaxPooling2D.py:53-58, :73-80)
 *   - the NumPy bodies of Layer_Dense.forward/backward (src/layers/Dense.py:64-108),
 *     Activation_ReLU (src/activations/Relu.py:7-14), Activation_Softmax
 *     (src/activations/Softmax.py:7-22), Loss_CategoricalCrossentropy.forward
 *     (src/losses/categorical_crossentropy.py:8-25), the fused softmax-CE gradient
 *     (src/losses/Activation_Softmax_Loss_CategoricalCrossentropy.py:4-19),
 *     Optimizer_Adam.update_params (src/optimizers/Adam.py:49-100),
 *     Optimizer_SGD.update_params (src/optimizers/SGD.py:55-97) and
 *     federated_average (Federated/utils.py:20-23).
 *
 * Layouts: FP32 tensors are exactly the reference's (NCHW activations, OIHW conv
 * weights, (in,out) Dense weights, row-major).  The bf16 tensor-core path
 * (nm_tc_*) uses its own NHWC-padded activation layout, documented per call.
 */
#ifndef NEUROMAH_B200_H
#define NEUROMAH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------- */
#define NM_OK               0
#define NM_ERR_BAD_ARG     -1   /* Python wrapper re-raises ValueError            */
#define NM_ERR_CUDA        -2   /* RuntimeError (reference: std::runtime_error)   */
#define NM_ERR_NCCL        -3
#define NM_ERR_OOM         -4   /* MemoryError                                    */
#define NM_ERR_UNSUPPORTED -5

const char* nm_last_error(void);          /* thread-local, never NULL */
int nm_version(void);                     /* ABI version of this header: 1 */

/* ---- device, memory, streams, events, graphs ---------------------------- */
int nm_device_count(int* count);
int nm_init(int device);                  /* cudaSetDevice + context warm-up */
int nm_device_props(int device, int* sm_count, int* cc_major, int* cc_minor,
                    size_t* total_mem);
int nm_malloc(void** dptr, size_t bytes);
int nm_free(void* dptr);
int nm_host_alloc(void** hptr, size_t bytes);       /* pinned host memory */
int nm_host_free(void* hptr);
int nm_memset(void* dptr, int byte, size_t bytes, void* stream);
int nm_h2d(void* dst, const void* src, size_t bytes, void* stream);
int nm_d2h(void* dst, const void* src, size_t bytes, void* stream);
int nm_d2d(void* dst, const void* src, size_t bytes, void* stream);
int nm_stream_create(void** stream);
int nm_stream_destroy(void* stream);
int nm_stream_sync(void* stream);
int nm_device_sync(void);
int nm_event_create(void** event);
int nm_event_destroy(void* event);
int nm_event_record(void* event, void* stream);
int nm_event_sync(void* event);
int nm_event_elapsed_ms(void* start, void* stop, float* ms);
int nm_stream_wait_event(void* stream, void* event);
int nm_graph_begin(void* stream);                        /* stream capture (relaxed mode) */
int nm_graph_end(void* stream, void** graph_exec);
int nm_graph_abort(void* stream);                        /* close a capture whose body failed; discards the partial graph */
int nm_graph_launch(void* graph_exec, void* stream);
int nm_graph_destroy(void* graph_exec);
/* cudaProfilerStart/Stop: bracket the region `ncu --profile-from-start off` captures (profiles/prof_step.py) */
int nm_profiler_start(void);
int nm_profiler_stop(void);
/* number of kernels this library launched since process start (bench "gpu_launches") */
int nm_launch_count(unsigned long long* count);

/* ---- FP32-exact path (CUDA-core FMA, reference layouts) ------------------ */

/* conv2d_cpu replacement (Conv2D.cpp:83-146) + the np.pad of Conv_wrapepr.py:78-84
 * folded in as pad_* (zeros), + optional bias (NULL = reference behaviour, Q4)
 * + optional ReLU epilogue (Conv_wrapepr.py:96-98).
 * x[N,C,H,W], w[OC,C,kH,kW], y[N,OC,OH,OW], OH = H+pad_t+pad_b-kH+1, stride 1. */
int nm_conv2d_fprop_f32(const float* x, const float* w, const float* bias, float* y,
                        int N, int C, int H, int W, int OC, int kH, int kW,
                        int pad_t, int pad_b, int pad_l, int pad_r, int relu, void* stream);
/* dgrad of the above (conv2d_backward_cpu, Conv2D.cpp:218-222 + col2im :48-74),
 * returned already cropped to the un-padded input: dx[N,C,H,W]. */
int nm_conv2d_dgrad_f32(const float* dy, const float* w, float* dx,
                        int N, int C, int H, int W, int OC, int kH, int kW,
                        int pad_t, int pad_b, int pad_l, int pad_r, void* stream);
/* wgrad (Conv2D.cpp:207-216: sum over batch, no normalisation) + bias gradient
 * (Conv_wrapepr.py:107-109: sum over n,h,w -> [OC]).  dbias may be NULL.
 * workspace: device scratch of nm_conv2d_wgrad_f32_workspace() bytes; the
 * split-K partials are reduced in a fixed order (deterministic). */
size_t nm_conv2d_wgrad_f32_workspace(int N, int C, int H, int W, int OC, int kH, int kW,
                                     int pad_t, int pad_b, int pad_l, int pad_r);
int nm_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* dbias,
                        void* workspace, size_t workspace_bytes,
                        int N, int C, int H, int W, int OC, int kH, int kW,
                        int pad_t, int pad_b, int pad_l, int pad_r, void* stream);

/* Activation_ReLU forward / backward (Relu.py:7-14).  `ref` is the tensor the mask
 * is taken from: ref <= 0 -> 0 (pre-activation and post-activation are equivalent). */
int nm_relu_fwd_f32(const float* x, float* y, size_t n, void* stream);
int nm_relu_bwd_f32(const float* dy, const float* ref, float* dx, size_t n, void* stream);

/* maxpool2d_forward_cpu / maxpool2d_backward_cpu (maxpooling_backend.cpp:24-133, :145-238).
 * idx[N,C,OH,OW,2] int32 = (row, col) of the max in PADDED coordinates; zero padding
 * participates in the max; ties keep the first element of the row-major window scan. */
int nm_maxpool2d_fwd_f32(const float* x, float* y, int32_t* idx,
                         int N, int C, int H, int W, int ph, int pw, int sh, int sw,
                         int pad_t, int pad_l, int OH, int OW, void* stream);
int nm_maxpool2d_bwd_f32(const float* dy, const int32_t* idx, float* dx,
                         int N, int C, int H, int W, int ph, int pw, int sh, int sw,
                         int pad_t, int pad_l, int OH, int OW, void* stream);

/* Layer_Dense.forward: y = x @ w + b (+ReLU)  (Dense.py:79-85). x[B,in] w[in,out] b[out] */
int nm_dense_fprop_f32(const float* x, const float* w, const float* b, float* y,
                       int B, int n_in, int n_out, int relu, void* stream);
/* dinputs = dy @ w^T (Dense.py:108) */
int nm_dense_dgrad_f32(const float* dy, const float* w, float* dx,
                       int B, int n_in, int n_out, void* stream);
/* dweights = x^T dy * scale + l1*sign(w) + 2*l2*w ; dbiases = sum dy * scale
 * (Dense.py:96-105; scale = 1/batch -- the GLOBAL batch under data parallelism). */
size_t nm_dense_wgrad_f32_workspace(int B, int n_in, int n_out);
int nm_dense_wgrad_f32(const float* x, const float* dy, const float* w,
                       float* dw, float* db, void* workspace, size_t workspace_bytes,
                       int B, int n_in, int n_out, float scale, float l1, float l2,
                       void* stream);

/* Activation_Softmax.forward (Softmax.py:7-11): row softmax with max subtraction. */
int nm_softmax_fwd_f32(const float* logits, float* probs, int B, int n, void* stream);
/* Jacobian backward (Softmax.py:13-18), closed form s*(d - <s,d>). */
int nm_softmax_bwd_f32(const float* probs, const float* dy, float* dx, int B, int n, void* stream);
/* Loss_CategoricalCrossentropy.forward per-sample losses with clip [1e-7, 1-1e-7]
 * (categorical_crossentropy.py:8-25), argmax==label flags (Accuracy_Categorical),
 * and -- if dlogits != NULL -- the fused gradient (p - onehot)*inv_batch
 * (Activation_Softmax_Loss_CategoricalCrossentropy.py:4-19).  labels = class ids.
 * accum (may be NULL): double[2] device accumulators {sum of losses, #correct},
 * updated by a single thread block in a fixed order (deterministic). */
int nm_softmax_ce_f32(const float* probs, const int32_t* labels, float* dlogits,
                      float* sample_loss, int32_t* correct, double* accum,
                      int B, int n, float inv_batch, void* stream);
/* Loss_CategoricalCrossentropy.backward: -onehot/p * inv_batch (:27-38) */
int nm_cce_bwd_f32(const float* probs, const int32_t* labels, float* dx,
                   int B, int n, float inv_batch, void* stream);

/* Activation_Sigmoid.forward / backward (src/activations/Sigmoid.py:5-10): y = 1 / (1 + exp(-x));
 * dx = dy * (1 - out) * out with `out` the forward output. */
int nm_sigmoid_fwd_f32(const float* x, float* y, size_t n, void* stream);
int nm_sigmoid_bwd_f32(const float* dy, const float* out, float* dx, size_t n, void* stream);

/* The element-wise losses of src/losses: Loss_MeanSquaredError (MSE.py:10-28), Loss_MeanAbsoluteError
 * (MAE.py:11-28), Loss_BinaryCrossentropy (binary_crossentropy.py:9-33).  pred / target are [B, D]
 * (every non-batch dimension flattened into D).
 * forward: sample_loss[b] = mean over D of (t - p)^2 | |t - p| | -(t log p' + (1 - t) log(1 - p')),
 *          p' = clip(p, 1e-7, 1 - 1e-7); one warp per sample, fixed reduction order.
 * backward: dx = -2 (t - p) / D | sign(t - p) / D (the reference's sign, MAE.py:23) |
 *          -(t / p' - (1 - t) / (1 - p')) / D.  As in the reference the gradient is divided by the
 *          outputs per sample only (prod(shape) / samples), NOT by the batch: Layer_Dense.backward
 *          applies the 1 / batch (Dense.py:96-98). */
#define NM_LOSS_MSE 0
#define NM_LOSS_MAE 1
#define NM_LOSS_BCE 2
int nm_loss_fwd_f32(int kind, const float* pred, const float* target, float* sample_loss,
                    int B, int D, void* stream);
int nm_loss_bwd_f32(int kind, const float* pred, const float* target, float* dx,
                    int B, int D, void* stream);

/* ---- optimizers over a flat FP32 slab ------------------------------------ */
/* Optimizer_Adam.update_params (Adam.py:86-100) applied n_apply times in registers
 * with the same t (Model registers every trainable layer twice: Model.py:56-59,:82-83).
 * bc1 = 1-beta1^(t+1), bc2 = 1-beta2^(t+1) computed by the host in double.
 * One coalesced 128-bit pass: 28 B/param regardless of n_apply. */
int nm_adam_step_f32(float* param, const float* grad, float* m, float* v, size_t n,
                     float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                     int n_apply, void* stream);
/* Graph-capturable variant: hyper-parameters live in a device struct that
 * nm_adam_advance_f32 refreshes on the device each step (no host round trip). */
typedef struct nm_adam_state {
    float lr, beta1, beta2, eps;
    float bc1, bc2, base_lr, decay;
    long long iterations;
    double beta1_pow, beta2_pow;      /* beta^(iterations+1) */
    unsigned ticket, reserved;        /* last-block ticket of the fused step + advance launch (always left at 0) */
    double base_lr_d, decay_d, beta1_d, beta2_d;   /* the host's float64 hyper-parameters: the bias corrections and the
                                                      decayed rate are formed in double exactly as Adam.py:41-47, :94-95 */
} nm_adam_state;
int nm_adam_state_init(nm_adam_state* dev_state, double lr, double decay, double beta1,
                       double beta2, float eps, long long iterations, void* stream);
int nm_adam_step_dev_f32(float* param, const float* grad, float* m, float* v, size_t n,
                         const nm_adam_state* dev_state, int n_apply, void* stream);
int nm_adam_advance_f32(nm_adam_state* dev_state, void* stream);   /* post_update_params */
/* nm_adam_step_dev_f32 + nm_adam_advance_f32 in ONE launch: the last block to finish advances the state (every block has
 * read the hyper-parameters by then), which takes a 1-thread kernel off the end of every training step. */
int nm_adam_step_advance_dev_f32(float* param, const float* grad, float* m, float* v, size_t n,
                                 nm_adam_state* dev_state, int n_apply, void* stream);
/* Optimizer_SGD.update_params (SGD.py:85-97); mom may be NULL when momentum == 0.
 * nonfinite (may be NULL): device int set to 1 if any gradient is NaN/Inf (SGD.py:80-83). */
int nm_sgd_step_f32(float* param, const float* grad, float* mom, size_t n,
                    float lr, float momentum, int n_apply, int32_t* nonfinite, void* stream);
/* The remaining optimizers, same one-pass 128-bit slab kernel, update applied n_apply times:
 * Optimizer_RMSprop.update_params (RMSprop.py:103-109): cache = rho cache + (1-rho) g^2; p -= lr g / (sqrt(cache) + eps)
 * Optimizer_Adagrad.update_params (Adagrad.py): cache += g^2; p -= lr g / (sqrt(cache) + eps)
 * Optimizer_Nadam.update_params (Nadam.py): bc1 = 1-beta1^(t+1), bc2 = 1-beta2^(t+1) from the host in double;
 *   m_hat = m/(bc1+eps), v_hat = v/(bc2+eps), p -= lr (beta1 m_hat + (1-beta1) g/(bc1+eps)) / (sqrt(v_hat) + eps) */
int nm_rmsprop_step_f32(float* param, const float* grad, float* cache, size_t n, float lr, float rho, float eps,
                        int n_apply, int32_t* nonfinite, void* stream);
int nm_adagrad_step_f32(float* param, const float* grad, float* cache, size_t n, float lr, float eps,
                        int n_apply, int32_t* nonfinite, void* stream);
int nm_nadam_step_f32(float* param, const float* grad, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                      float eps, float bc1, float bc2, int n_apply, void* stream);
/* Layer_Dropout.forward (Dropout.py:15-30): mask = Bernoulli(keep)/keep drawn on the device (Philox4x32-10, counter =
 * (element/4, offset), key = seed), y = x * mask; the mask is kept for backward (Dropout.py:32-40: dx = dy * mask). */
int nm_dropout_fwd_f32(const float* x, float* y, float* mask, size_t n, float keep, unsigned long long seed,
                       unsigned long long offset, void* stream);
int nm_mul_f32(float* y, const float* a, const float* b, size_t n, void* stream);      /* y = a * b elementwise */
/* y = a*x (+ y if accumulate): FedAvg pre-scale n_i/N (Federated/utils.py:21-22). */
int nm_scale_f32(float* y, const float* x, size_t n, float a, int accumulate, void* stream);
/* Raw uint8 pixels -> float32: y = float(x) / denom with IEEE division, i.e. the host-side preprocessing
 * `X.astype(np.float32) / 255.0` of examples/test_Conv2D_MNIST.py:27 moved behind the host->device copy (4x fewer bytes
 * over PCIe); bit-identical to NumPy.  The bf16 plans take uint8 inputs directly (nm_tc_*_u8_bf16 below). */
int nm_u8_to_f32_div(const uint8_t* x, float* y, size_t n, float denom, void* stream);
/* TensorMonitor.log_histogram (src/utils/TensorMonitor.py:93-104: np.histogram(values.flatten(), bins)) without pulling the
 * tensor to the host: nm_minmax_f32 leaves {min, max} in out2 (device floats); nm_histogram_f32 counts into `bins` uniform
 * bins over [first, last] with NumPy's index arithmetic (float64 scaled index + the two edge corrections), so the counts
 * equal np.histogram's exactly; values outside the range and NaNs are not counted. */
int nm_minmax_f32(const float* x, size_t n, float* out2, void* stream);
int nm_histogram_f32(const float* x, size_t n, double first, double last, int bins, unsigned long long* counts, void* stream);

/* ---- bf16 tensor-core path (tcgen05 + TMA + TMEM), see DESIGN.md --------- */
/* Activations: bf16 NHWC with a one-pixel zero halo, [N, H+2, W+2, C] ("padded grid"); the halo is
 * written once (zeros) and never touched again.  Parameters and parameter gradients stay FP32 in the
 * reference layouts (arena); the MMAs read packed bf16 copies refreshed after every optimizer step.
 * Pool arg-max: one nibble per pooled element = window position (dy*2+dx, first max in the reference's
 * row-major scan, maxpooling_backend.cpp:94-108) | (value > 0) << 2  (the ReLU mask, Relu.py:12-14). */

/* OIHW f32 -> bf16 [OC][tap][C] (fprop B operand) and bf16 [C][8-tap][OC] (dgrad B operand); either may be NULL */
int nm_tc_pack_conv_weights_bf16(const float* w_oihw, void* w_fprop, void* w_dgrad, int OC, int C, void* stream);
/* Layer_Conv2D.forward (3x3, 'same', ReLU, no bias: Conv_wrapepr.py:73-98) fused with the following
 * Layer_MaxPooling2D(2, 2) forward.  x_pad bf16 [N,H+2,W+2,C]; y bf16 [N,H/2,W/2,OC] (out_padded = 0)
 * or [N,H/2+2,W/2+2,OC] written at (+1,+1) (out_padded = 1); idx nibbles [N,H/2,W/2,OC/2] bytes.
 * tcgen05 implicit GEMM; instantiated for C=32, OC=64, H=W=14 (NM_ERR_UNSUPPORTED otherwise). */
int nm_tc_conv3x3_fprop_pool_bf16(const void* x_pad, const void* w_packed, void* y, void* idx,
                                  int N, int H, int W, int C, int OC, int out_padded, void* stream);
/* dgrad (Conv2D.cpp:218-222 + col2im): dy_pad bf16 [N,H+2,W+2,OC] (zero halo) -> dx bf16 on the un-shifted
 * grid [N,H+2,W+2,C] (pixel (y,x) at row y*(W+2)+x; rows/cols >= H/W hold don't-care values). */
int nm_tc_conv3x3_dgrad_bf16(const void* dy_pad, const void* w_packed_dgrad, void* dx_grid,
                             int N, int H, int W, int C, int OC, void* stream);
/* wgrad (Conv2D.cpp:207-216, summed over the batch, no normalisation) -> dw f32 OIHW; split over the
 * pixel range across one CTA per SM, partials reduced in a fixed order (deterministic). */
size_t nm_tc_conv3x3_wgrad_workspace(int C, int OC);
int nm_tc_conv3x3_wgrad_bf16(const void* x_pad, const void* dy_pad, float* dw, void* workspace, size_t workspace_bytes,
                             int N, int H, int W, int C, int OC, void* stream);
/* dgrad AND wgrad of the 3x3 conv in one pass over dY (the same shared-memory dY tile is the K-major operand of the
 * dgrad MMAs and the MN-major operand of the wgrad MMAs).  Same inputs/outputs as the two calls above. */
size_t nm_tc_conv3x3_bwd_workspace(int C, int OC);
int nm_tc_conv3x3_bwd_bf16(const void* x_pad, const void* dy_pad, const void* w_packed_dgrad, float* dw, void* dx_grid,
                           void* workspace, size_t workspace_bytes, int N, int H, int W, int C, int OC, void* stream);
/* dw == NULL in nm_tc_conv3x3_bwd_*: the per-CTA wgrad partials stay in the workspace and this call reduces them (fixed
 * order) later -- e.g. on a sibling graph branch while the previous layer's backward already consumes dx_grid. */
int nm_tc_conv3x3_bwd_reduce_f32(const void* workspace, float* dw, int N, int H, int W, int C, int OC, void* stream);
/* The same, with dY never materialised: it is rebuilt per tile in shared memory from the ReLU-masked gradient of the
 * POOLED conv output (dpool bf16 [N][H/2*W/2][OC], nm_tc_dense_bwd_pool_bf16) and the pool arg-max nibbles
 * (idx [N][H/2*W/2][OC/2]) -- MaxPooling2D.backward's routing is one byte permute per two channels. */
int nm_tc_conv3x3_bwd_pool_bf16(const void* x_pad, const void* dpool, const void* idx, const void* w_packed_dgrad, float* dw,
                                void* dx_grid, void* workspace, size_t workspace_bytes, int N, int H, int W, int C, int OC,
                                void* stream);

/* ---- general-shape 3x3 'same' convolution (streamed filter bank), channel counts padded to multiples of 64 ----
 * The VGG-style stack of BASELINE.json config 4.  All activation tensors are haloed bf16 NHWC [N,H+2,W+2,Cp]; the
 * kernels write the halo of their output as zero, so outputs feed the next layer (or its backward) directly. */
/* OIHW f32 [OC][C][3][3] -> bf16 [OCp][tap][Cp] (fprop operand) and bf16 [Cp][8-tap][OCp] (dgrad operand), padded
 * channels zero; either destination may be NULL. */
int nm_tc_pack_conv_weights_gen_bf16(const float* w_oihw, void* w_fprop, void* w_dgrad, int OC, int C, int OCp, int Cp, void* stream);
/* f32 NCHW [N,C,H,W] -> haloed bf16 NHWC [N,H+2,W+2,Cp] (replaces np.pad, Conv_wrapepr.py:78-84); every element written. */
int nm_tc_pack_input_nhwc_bf16(const float* x_nchw, void* x_pad, int N, int C, int H, int W, int Cp, void* stream);
/* ... from raw uint8 pixels, value = float(x) * scale (scale = 1/255: the example scripts' normalisation). */
int nm_tc_pack_input_nhwc_u8_bf16(const uint8_t* x_nchw, float scale, void* x_pad, int N, int C, int H, int W, int Cp, void* stream);
/* y_pad[N,H+2,W+2,Cout] = conv3x3(x_pad[N,H+2,W+2,Cin], w_packed[Cout][9][Cin]) (+ bias[Cout] if not NULL)
 * (ReLU if relu != 0) (zeroed where relu_mask[N,H+2,W+2,Cout] <= 0, if not NULL).  tcgen05 implicit GEMM, K loop over
 * (64-channel chunk, tap).  fprop (Conv2D.cpp:83-146 + Conv_wrapepr.py:96-98): w_packed = the fprop operand, relu = 1.
 * dgrad (Conv2D.cpp:218-222 + col2im, then the ReLU backward of the layer below, Relu.py:12-14): x_pad = dY,
 * w_packed = the dgrad operand, Cin = OCp, Cout = Cp, relu_mask = that layer's (post-ReLU) output. */
int nm_tc_conv3x3_gen_bf16(const void* x_pad, const void* w_packed, const float* bias, const void* relu_mask, void* y_pad,
                           int N, int H, int W, int Cin, int Cout, int relu, void* stream);
/* wgrad (Conv2D.cpp:207-216: summed over the batch, no normalisation) -> dw f32 OIHW [OC][C][3][3], and, if dbias is
 * not NULL, the bias gradient sum_{n,h,w} dY -> dbias f32 [OC] (Conv_wrapepr.py:107-109).  Deterministic. */
size_t nm_tc_conv3x3_wgrad_gen_workspace(int N, int H, int W, int C, int Cp, int OCp);
int nm_tc_conv3x3_wgrad_gen_bf16(const void* x_pad, const void* dy_pad, float* dw, float* dbias, void* workspace, size_t workspace_bytes,
                                 int N, int H, int W, int C, int OC, int Cp, int OCp, void* stream);

/* Layer_MaxPooling2D(2, 2).forward (MaxPooling2D.py:37-61, maxpooling_backend.cpp:24-133) on a haloed bf16 NHWC tensor:
 * x_pad [N,H+2,W+2,C] -> y bf16 [N,H/2,W/2,C] (out_padded = 0) or haloed [N,H/2+2,W/2+2,C] with the halo written as
 * zero (out_padded = 1); idx nibbles [N,H/2,W/2,C/2] bytes in the format above (first-max position | (max > 0) << 2). */
int nm_tc_maxpool2x2_fwd_bf16(const void* x_pad, void* y, void* idx, int N, int H, int W, int C, int out_padded, void* stream);
/* Layer_MaxPooling2D.backward (maxpooling_backend.cpp:145-238) + the ReLU backward of the conv below (Relu.py:12-14):
 * dpool bf16 [N,H/2,W/2,C] (in_padded = 0) or haloed (in_padded = 1) + idx -> dy_pad bf16 [N,H+2,W+2,C], every element
 * written (the halo and the non-arg-max positions as zero). */
int nm_tc_maxpool2x2_bwd_bf16(const void* dpool, const void* idx, void* dy_pad, int N, int H, int W, int C, int in_padded, void* stream);
/* Layer_Dense dinputs (Dense.py:108) for the tensor-core mode: dx bf16 [B][K] = dlogits f32 [B][n_out] . w_packed
 * (f32 [n_out][K] from nm_tc_pack_dense_weights, K in (h,w,c) order).  n_out <= 16, K % 8 == 0. */
int nm_tc_dense_dgrad_bf16(const float* dlogits, const float* w_packed, void* dx, int B, int K, int n_out, void* stream);
/* ... with the ReLU backward of the layer that produced the Dense input (Relu.py:12-14): dx = 0 where relu_mask[B][K] <= 0,
 * dx * mask_scale elsewhere (mask_scale = 1 / keep when a Layer_Dropout sits in between: its surviving elements are > 0). */
int nm_tc_dense_dgrad_mask_bf16(const float* dlogits, const float* w_packed, const void* relu_mask, float mask_scale, void* dx, int B, int K,
                                int n_out, void* stream);

/* ---- general tcgen05 Dense layers (hidden layers of an MLP), nm_tc_gemm.cu ----------------------------------------
 * Layer_Dense.forward / backward (Dense.py:79-85, :96-108) + the embedded ReLU (Relu.py:7-14) for any n_in % 8 == 0 and
 * n_out % 64 == 0, TMA-fed tcgen05 tiles with FP32 accumulation; activations and their gradients are bf16 [B][features].
 *   nm_tc_to_bf16 / nm_tc_u8_to_bf16   the first layer's A operand from float32 inputs / raw uint8 pixels * scale
 *   nm_tc_pack_dense_gen_bf16          w f32 [n_in][n_out] -> w_kn bf16 (same layout, dgrad operand), w_nk bf16 [n_out][n_in]
 *                                      (fprop operand); either may be NULL; refreshed after every optimizer step
 *   nm_tc_dense_fprop_gen_bf16         y = x w (+ bias) (ReLU), then a fused Layer_Dropout (Dropout.py:15-30) when drop_keep < 1:
 *                                      y *= Bernoulli(drop_keep) / drop_keep from the Philox stream of nm_dropout_fwd_f32
 *                                      (key drop_seed, offset drop_offset + *drop_offset_dev if not NULL)
 *   nm_tc_dense_dgrad_gen_bf16         dx = dy w^T; with relu_mask: 0 where relu_mask <= 0, x mask_scale elsewhere
 *   nm_tc_dense_wgrad_gen_bf16         dw = scale x^T dy + l1 sign(w) + 2 l2 w, db = scale sum_b dy; both operands are read
 *                                      MN-major straight from x / dy; split over the batch, partials summed in a fixed order */
int nm_tc_to_bf16(const float* x, void* y, size_t n, void* stream);
int nm_tc_u8_to_bf16(const uint8_t* x, float scale, void* y, size_t n, void* stream);
int nm_tc_pack_dense_gen_bf16(const float* w, void* w_kn, void* w_nk, int n_in, int n_out, void* stream);
int nm_tc_dense_fprop_gen_bf16(const void* x, const void* w_nk, const float* bias, void* y, int B, int n_in, int n_out, int relu,
                               float drop_keep, unsigned long long drop_seed, unsigned long long drop_offset,
                               const long long* drop_offset_dev, void* stream);
int nm_tc_dense_dgrad_gen_bf16(const void* dy, const void* w_kn, const void* relu_mask, float mask_scale, void* dx, int B, int n_in,
                               int n_out, void* stream);
size_t nm_tc_dense_wgrad_gen_workspace(int B, int n_in, int n_out);
int nm_tc_dense_wgrad_gen_bf16(const void* x, const void* dy, const float* w, float* dw, float* db, void* workspace, size_t workspace_bytes,
                               int B, int n_in, int n_out, float scale, float l1, float l2, void* stream);

/* ---- FP32-exact Dense layers on the tensor cores (mode "bf16x3"), nm_tc_gemm.cu ------------------------------------
 * The same three contractions (Dense.py:79-85, :96-108) with every operand carried as THREE bf16 planes h + m + l
 * (h = bf16(v), m = bf16(v - h), l = bf16(v - h - m): 24 mantissa bits) and the six plane pairs above 2^-24 accumulated
 * in one FP32 TMEM tile, smallest first; the epilogues split their FP32 results back into planes.  Meets the FP32-exact
 * tolerance of the CUDA-core kernels (rtol 1e-4 / atol 1e-5 against the NumPy path).  A "x3" matrix is three bf16
 * [rows][cols] planes `*_plane` elements apart (multiple of 8, >= rows * cols); weight planes are rows * cols apart.
 *   nm_tc_split3_f32 / nm_tc_split3_u8   float32 (optionally through a ReLU mask: 0 where relu_mask <= 0, x mask_scale
 *                                        elsewhere) / raw uint8 pixels divided by denom -> planes
 *   nm_tc_pack_dense_gen_x3              w f32 [n_in][n_out] -> w_kn3 [3][n_in][n_out], w_nk3 [3][n_out][n_in]
 *   nm_tc_dense_fprop_gen_x3             as nm_tc_dense_fprop_gen_bf16; y_f32 (optional) also receives the FP32 result
 *   nm_tc_dense_dgrad_gen_x3             relu_mask = the h plane of the activation below
 *   nm_tc_dense_wgrad_gen_x3             workspace: nm_tc_dense_wgrad_gen_workspace */
int nm_tc_split3_f32(const float* x, const void* relu_mask, float mask_scale, void* planes, size_t n, size_t plane_stride, void* stream);
int nm_tc_split3_u8(const uint8_t* x, float denom, void* planes, size_t n, size_t plane_stride, void* stream);
int nm_tc_pack_dense_gen_x3(const float* w, void* w_kn3, void* w_nk3, int n_in, int n_out, void* stream);
int nm_tc_dense_fprop_gen_x3(const void* x3, size_t x_plane, const void* w_nk3, const float* bias, void* y3, size_t y_plane, float* y_f32,
                             int B, int n_in, int n_out, int relu, float drop_keep, unsigned long long drop_seed,
                             unsigned long long drop_offset, const long long* drop_offset_dev, void* stream);
int nm_tc_dense_dgrad_gen_x3(const void* dy3, size_t dy_plane, const void* w_kn3, const void* relu_mask, float mask_scale, void* dx3,
                             size_t dx_plane, int B, int n_in, int n_out, void* stream);
int nm_tc_dense_wgrad_gen_x3(const void* x3, size_t x_plane, const void* dy3, size_t dy_plane, const float* w, float* dw, float* db,
                             void* workspace, size_t workspace_bytes, int B, int n_in, int n_out, float scale, float l1, float l2,
                             void* stream);

/* ---- FP32-exact conv stacks on the tensor cores (mode "bf16x3"), nm_tc_conv_gen.cu / nm_tc_pool.cu -------------------
 * The general conv kernels above on three bf16 planes per tensor (see the Dense "x3" block): same layouts per plane
 * (zero-haloed NHWC, packed weights), planes `*_plane` elements apart (weights: OCp * 9 * Cp), six plane-pair passes per
 * contraction into one TMEM accumulator, FP32 results split back into planes in the epilogue.
 *   nm_tc_pack_conv_weights_gen_x3 / nm_tc_pack_input_nhwc_x3 / _u8_x3   planes of the packed weights / the haloed input
 *                                      (uint8 pixels are DIVIDED by denom, the FP32 path's division)
 *   nm_tc_conv3x3_gen_x3               fprop (+ bias + ReLU) or dgrad (+ ReLU mask = h plane of the activation below)
 *   nm_tc_conv3x3_wgrad_gen_x3         dW, db (workspace: nm_tc_conv3x3_wgrad_gen_workspace)
 *   nm_tc_maxpool2x2_fwd_x3            window maximum of the exact FP32 values (l + m) + h, first maximum wins; idx as in
 *                                      nm_tc_maxpool2x2_fwd_bf16 (the backward is nm_tc_maxpool2x2_bwd_bf16 once per plane)
 *   nm_tc_merge3_nhwc_to_nchw_f32      planes [3][N][HW][Cp] -> float32 [N][C * HW] in Flatten's order (input of the FP32 head;
 *                                      C <= Cp: the zero-padded channels of the last conv are dropped)
 *   nm_tc_split3_nchw_to_nhwc          float32 [N][C * HW] -> planes [3][N][HW][Cp] (the head's input gradient; zeros in the padding) */
int nm_tc_pack_conv_weights_gen_x3(const float* w_oihw, void* w_fprop3, void* w_dgrad3, int OC, int C, int OCp, int Cp, void* stream);
int nm_tc_pack_input_nhwc_x3(const float* x_nchw, void* x_pad3, size_t plane, int N, int C, int H, int W, int Cp, void* stream);
int nm_tc_pack_input_nhwc_u8_x3(const uint8_t* x_nchw, float denom, void* x_pad3, size_t plane, int N, int C, int H, int W, int Cp, void* stream);
int nm_tc_conv3x3_gen_x3(const void* x_pad3, size_t x_plane, const void* w_packed3, const float* bias, const void* relu_mask, void* y_pad3,
                         size_t y_plane, int N, int H, int W, int Cin, int Cout, int relu, void* stream);
int nm_tc_conv3x3_wgrad_gen_x3(const void* x_pad3, size_t x_plane, const void* dy_pad3, size_t dy_plane, float* dw, float* dbias,
                               void* workspace, size_t workspace_bytes, int N, int H, int W, int C, int OC, int Cp, int OCp, void* stream);
int nm_tc_maxpool2x2_fwd_x3(const void* x_pad3, size_t x_plane, void* y3, size_t y_plane, void* idx, int N, int H, int W, int C, int out_padded,
                            void* stream);
int nm_tc_merge3_nhwc_to_nchw_f32(const void* planes, size_t plane_stride, float* out, int N, int HW, int C, int Cp, void* stream);
int nm_tc_split3_nchw_to_nhwc(const float* x, void* planes, size_t plane_stride, int N, int HW, int C, int Cp, void* stream);

/* First conv of the net (C_in = 1, 3x3 'same', ReLU, no bias) + MaxPooling2D(2,2) on tcgen05: the 4x4 input patch of a
 * pool window is the K = 16 dimension and the weights are expanded to Wext[(window position, oc)][4x4] (zero outside the
 * 3x3 support), so one MMA yields all four conv values of a window per TMEM lane and the pool needs no cross-lane traffic.
 * w f32 [OC,1,3,3] -> w_ext bf16 [4*OC][16]. */
int nm_tc_pack_conv1_weights_bf16(const float* w, void* w_ext, int OC, void* stream);
/* x f32 [N,1,H,W] (rounded to bf16 as the MMA operand) -> y_pad bf16 [N,H/2+2,W/2+2,OC] (interior; the zero halo is
 * never written), idx nibbles [N,H/2,W/2,OC/2].  OC = 32. */
int nm_tc_conv1_fwd_pool_bf16(const float* x, const void* w_ext, void* y_pad, void* idx, int N, int H, int W, int OC, void* stream);
/* ... the same from raw uint8 pixels: the operand is bf16(float(x) * scale); with scale = 1/255 that equals
 * bf16(float32(x) / 255) -- the reference's normalisation -- for all 256 pixel values. */
int nm_tc_conv1_fwd_pool_u8_bf16(const uint8_t* x, float scale, const void* w_ext, void* y_pad, void* idx, int N, int H, int W, int OC,
                                 void* stream);
/* The same forward with the step's weight re-pack folded in (one launch less on the critical path of every training step):
 * the B operand is built from the FP32 conv1 weights w1 directly, and two spare warps per CTA refresh the packed copies of
 * conv2's and the Dense head's weights exactly as nm_tc_pack_cnn_weights_bf16 does.  x: float32, or raw uint8 pixels
 * (x_is_u8 != 0) normalised by u8_scale. */
int nm_tc_conv1_fwd_pool_pack_bf16(const void* x, int x_is_u8, float u8_scale, const float* w1, void* y_pad, void* idx, int N, int H, int W,
                                   int OC, const float* w2, void* w2_fprop, void* w2_dgrad, int OC2, int C2,
                                   const float* w3, float* w3_packed, void* w3_hilo, int K, int n_out, int HW, int C3, void* stream);
/* ... and backward: pool backward + ReLU backward (the nibble) + wgrad + bias gradient in one tcgen05 pass
 * (E[(pos,oc)][4x4 patch] accumulated in TMEM, one FP32 partial per CTA, fixed-order reduction).
 * dp_grid = gradient of the pooled output, bf16 on the un-shifted grid [N,H/2+2,W/2+2,OC] (nm_tc_conv3x3_dgrad_bf16). */
size_t nm_tc_conv1_bwd_workspace(int OC);
int nm_tc_conv1_bwd_bf16(const float* x, const void* dp_grid, const void* idx, float* dw, float* db,
                         void* workspace, size_t workspace_bytes, int N, int H, int W, int OC, void* stream);
int nm_tc_conv1_bwd_u8_bf16(const uint8_t* x, float scale, const void* dp_grid, const void* idx, float* dw, float* db,
                            void* workspace, size_t workspace_bytes, int N, int H, int W, int OC, void* stream);
/* Dense weights in the reference layout [(c,h,w) rows][n_out] f32 ->
 *   w_packed f32 [n_out][K] with K in the kernels' (h,w,c) order (CUDA-core dgrad; may be NULL)
 *   w_hilo  bf16 [32][K]: rows [0,16) = bf16(w), rows [16,32) = bf16(w - bf16(w)), classes >= n_out zero
 *           (tcgen05 B operand; the hi/lo pair keeps ~16 mantissa bits of the weights; may be NULL) */
int nm_tc_pack_dense_weights(const float* w, float* w_packed, void* w_hilo, int K, int n_out, int HW, int C, void* stream);
/* The three packs of the CNN plan (conv1 Wext, conv2 fprop/dgrad operands, Dense packed + hi/lo) in ONE launch. */
int nm_tc_pack_cnn_weights_bf16(const float* w1, void* w1_ext, const float* w2, void* w2_fprop, void* w2_dgrad, int OC2, int C2,
                                const float* w3, float* w3_packed, void* w3_hilo, int K, int n_out, int HW, int C3, void* stream);
/* f32 dlogits [B][n_out] -> bf16 hi/lo [B][32] (only needed when dlogits did not come from nm_tc_dense_softmax_ce_bf16) */
int nm_tc_pack_dlogits_bf16(const float* dlogits, void* dl_hilo, int B, int n_out, void* stream);
/* Layer_Dense.forward + Activation_Softmax + CCE losses + accuracy flags + fused gradient (p-onehot)*inv_batch
 * (Dense.py:79-80, Softmax.py:7-11, categorical_crossentropy.py:8-25, Activation_Softmax_Loss_...py:4-19).
 * TMA-fed tcgen05 tiles (M = 128 rows, N = 32 = hi+lo classes), split-K over CTAs, one finalize kernel.
 * x bf16 [B,K] (K % 8 == 0); dl_hilo (may be NULL) receives the gradient as bf16 hi/lo [B][32] for
 * nm_tc_dense_wgrad_bf16; accum (may be NULL) = double[2] device accumulators {loss sum, #correct}, updated
 * in a fixed order.  workspace: nm_tc_dense_softmax_ce_workspace() bytes, ZERO-FILLED once before first use
 * (it holds a ticket counter that every call leaves at zero).  n_out <= 16. */
size_t nm_tc_dense_softmax_ce_workspace(int B, int K);
int nm_tc_dense_softmax_ce_bf16(const void* x, const void* w_hilo, const float* bias, const int32_t* labels,
                                float* probs, float* dlogits, void* dl_hilo, float* sample_loss, int32_t* correct,
                                double* accum, void* workspace, size_t workspace_bytes,
                                int B, int K, int n_out, float inv_batch, void* stream);
/* Dense dinputs (Dense.py:108) + MaxPooling2D backward (maxpooling_backend.cpp:145-238) + ReLU backward
 * (Relu.py:12-14) -> the last conv's dY, bf16 [B,2PH+2,2PW+2,C] (interior); dbias[C] = sum of dY
 * (Conv_wrapepr.py:107-109).  One thread per (pooled pixel, 8 channels): PH*PW*C/8 <= 416. */
size_t nm_tc_dense_bwd_unpool_workspace(int C);
/* ... or only up to the pooled tensor: dpool bf16 [B][PH*PW][C] = ReLU-masked Dense dinputs (no arg-max routing). */
int nm_tc_dense_bwd_pool_bf16(const float* dlogits, const float* w_packed, const void* idx, void* dpool,
                              float* dbias, void* workspace, size_t workspace_bytes,
                              int B, int PH, int PW, int C, int n_out, void* stream);
int nm_tc_dense_bwd_unpool_bf16(const float* dlogits, const float* w_packed, const void* idx, void* dy_pad,
                                float* dbias, void* workspace, size_t workspace_bytes,
                                int B, int PH, int PW, int C, int n_out, void* stream);
/* Dense dweights / dbiases (Dense.py:96-105): tcgen05 tiles over X^T (M = 64 features) x dlogits hi/lo (N = 32),
 * split over the batch rows, fixed-order reduction, written in the reference row order (+ L1/L2 terms). */
size_t nm_tc_dense_wgrad_workspace(int B, int K);
int nm_tc_dense_wgrad_bf16(const void* x, const void* dl_hilo, const float* dlogits, const float* w, float* dw, float* db,
                           void* workspace, size_t workspace_bytes, int B, int K, int n_out, int HW, int C,
                           float scale, float l1, float l2, void* stream);

/* ---- data-parallel communication (NCCL over NVLink) ---------------------- */
int nm_comm_load(const char* libnccl_path);                 /* dlopen libnccl.so.2 */
int nm_comm_unique_id(void* id128);                          /* rank 0: 128-byte ncclUniqueId */
int nm_comm_init(void** comm, const void* id128, int rank, int world);
int nm_comm_destroy(void* comm);
int nm_comm_allreduce_sum_f32(void* comm, float* buf, size_t n, void* stream);  /* in place */
int nm_comm_broadcast_f32(void* comm, float* buf, size_t n, int root, void* stream);

/* ---- data-parallel exchange fused with the optimizer, over NVLink peer memory ----
 * No reference equivalent (the reference is single-process; its one cross-replica operation is the FedAvg mean,
 * Federated/utils.py:20-23).  Each rank owns a mailbox in its HBM that the peers map through cudaIpc; ONE kernel
 * pushes the local gradient chunks to every mailbox, waits for the peers' chunks, sums them in rank order (replicas
 * stay bit-identical) and applies Optimizer_Adam.update_params (Adam.py:86-100) -- the all-reduce and the optimizer
 * launch of the NCCL path in one launch, CUDA-graph replayable (the step number lives on the device).
 *   nm_dp_create   allocate the mailbox for slabs of up to max_floats floats (world <= 16)
 *   nm_dp_export   64-byte cudaIpcMemHandle_t of this rank's mailbox, to be all-gathered by the launcher
 *   nm_dp_connect  handles = world x 64 bytes in rank order (own entry ignored); maps the peers' mailboxes
 *   nm_dp_status   timed_out != 0 if a wait for a peer exceeded the timeout (default 20 s); blocking */
int nm_dp_create(void** dp, int rank, int world, size_t max_floats);
int nm_dp_export(void* dp, void* handle64);
int nm_dp_connect(void* dp, const void* handles);
int nm_dp_set_timeout_ms(void* dp, unsigned ms);
int nm_dp_allreduce_sum_f32(void* dp, float* buf, size_t n, void* stream);      /* in place, every rank gets the sum */
int nm_dp_allreduce_adam_dev_f32(void* dp, float* param, float* grad, float* m, float* v, size_t n,
                                 nm_adam_state* dev_state, int n_apply, int advance, void* stream);   /* advance != 0: + nm_adam_advance_f32 */
 /* Exchange overlapped with backward: nm_dp_push_f32 delivers chunks [chunk_lo, chunk_hi) of the gradient slab (chunk =
 * nm_dp_chunk_floats() floats) to every peer as soon as their producers have been enqueued -- stores only, no waits, meant
 * for a side stream behind each wgrad -- and nm_dp_collect_adam_dev_f32 is the fused kernel above told that chunks >=
 * pushed_from_chunk are already on their way: it pushes the rest (the first layer's gradients), collects, sums in rank
 * order and applies Adam.  All pushes of a step must precede its collect in stream / graph order. */
int nm_dp_chunk_floats(void);
int nm_dp_push_f32(void* dp, const float* grad, size_t n, int chunk_lo, int chunk_hi, void* stream);
int nm_dp_collect_adam_dev_f32(void* dp, float* param, float* grad, float* m, float* v, size_t n,
                               nm_adam_state* dev_state, int n_apply, int advance, int pushed_from_chunk, void* stream);
int nm_dp_status(void* dp, int* timed_out, unsigned* steps_done);
int nm_dp_destroy(void* dp);

#ifdef __cplusplus
}
#endif
#endif /* NEUROMAH_B200_H */
