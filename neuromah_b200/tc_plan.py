"""bf16 tensor-core execution plan for the Conv-Pool-Conv-Pool-Flatten-Dense-Softmax net.

``Model.finalize(mode="bf16")`` recognises the layer pattern of examples/test_Conv2D_MNIST.py:47-53 and
replaces the per-layer FP32 launches with nine fused launches per training step:

  forward   conv1+ReLU+pool (CUDA cores, C_in=1)  ->  conv2+ReLU+pool (tcgen05 implicit GEMM)
            ->  Dense+softmax+CE+grad (+ loss/accuracy sums)
  backward  Dense wgrad  |  Dense dgrad+unpool+ReLU' -> dY2 (+conv2 bias grad)  |  conv2 wgrad (tcgen05)
            |  conv2 dgrad (tcgen05)  |  conv1 unpool+ReLU'+wgrad+bias grad
  update    one Adam/SGD launch over the arena  +  re-pack of the bf16 operand copies

Activations are bf16 NHWC with a zero halo; parameters, gradients and optimizer state stay FP32 in the
arena in the reference's layouts, so optimizers, all-reduce, FedAvg, save/load are shared with the FP32 path.
Numerics: bf16 operands, FP32 accumulation (stated tolerance in tests/test_gpu_tc_model.py).
"""
from __future__ import annotations

import numpy as np

from . import ops
from ._lib import lib
from .device import DeviceArray, Scratch, stream


def _timed(name, fn, *a):
    tok = ops._tic(name)
    fn(*a)
    ops._toc(tok)


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


class Bf16Activation:
    """Host-convertible view of an internal bf16 NHWC (optionally haloed) activation: np.asarray(view)
    yields the reference's float32 NCHW tensor."""

    def __init__(self, buf: DeviceArray, n, halo, channels=None):
        self.buf, self.n, self.halo = buf, n, halo
        _, hp, wp, c = buf.shape
        self.channels = c if channels is None else channels       # the layer's own channels (the buffer may be zero-padded)
        self.shape = (n, self.channels, hp - 2 * halo, wp - 2 * halo)
        self.dtype = np.dtype(np.float32)
        self.ndim = 4

    def __len__(self):
        return self.n

    def numpy(self):
        a = bf16_bits_to_f32(self.buf.numpy()[:self.n])
        if self.halo:
            a = a[:, self.halo:-self.halo, self.halo:-self.halo, :]
        return np.ascontiguousarray(a[..., :self.channels].transpose(0, 3, 1, 2))

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


class TcCnnPlan:
    fused_pack = True       # weight re-pack inside the conv1 forward kernel (False: the separate nm_tc_pack_cnn_weights_bf16 launch)
    PATTERN = "Conv2D(1->32,3x3,same,ReLU) Pool(2,2) Conv2D(32->64,3x3,same,ReLU) Pool(2,2) Flatten Dense(->n<=16) Softmax"

    @staticmethod
    def match(model):
        from .src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten
        from .src.activations import Activation_ReLU, Activation_Softmax
        L = model.layers
        if len(L) != 7:
            return False
        c1, p1, c2, p2, fl, de, sm = L
        ok = (isinstance(c1, Layer_Conv2D) and isinstance(c2, Layer_Conv2D) and isinstance(p1, Layer_MaxPooling2D)
              and isinstance(p2, Layer_MaxPooling2D) and isinstance(fl, Layer_Flatten) and isinstance(de, Layer_Dense)
              and isinstance(sm, Activation_Softmax))
        if not ok:
            return False
        for c, shape in ((c1, (32, 1, 3, 3)), (c2, (64, 32, 3, 3))):
            if c.weights.shape != shape or c.padding != "same" or type(c.activation) is not Activation_ReLU:
                return False
        for p in (p1, p2):
            if p.pool_size != (2, 2) or p.strides != (2, 2) or p.padding != "valid":
                return False
        return de.activation is None and de.weights.shape[1] <= 16 and de.weights.shape[0] % 64 == 0

    def __init__(self, model):
        if not self.match(model):
            raise ValueError(f"mode='bf16' supports the layer pattern: {self.PATTERN}")
        if model.softmax_classifier_output is None:
            raise ValueError("mode='bf16' needs Loss_CategoricalCrossentropy with a final Activation_Softmax")
        from . import config
        if config.conv_backward != "wired" or config.conv_bias_in_forward:
            raise ValueError("mode='bf16' implements the default quirk settings only")
        self.model = model
        self.c1, self.p1, self.c2, self.p2, self.fl, self.de, self.sm = model.layers
        self.n_out = self.de.weights.shape[1]
        self.B = 0
        self.scratch = Scratch()
        self.w2f = DeviceArray((64, 9, 32), np.uint16)
        self.w2d = DeviceArray((32, 9, 64), np.uint16)
        self.w1e = DeviceArray((128, 16), np.uint16)          # conv1 weights expanded over the 4 pool-window positions
        self.w3p = None
        self.launches_per_step = 0
        # side branch of the backward pass: created here, never inside a CUDA-graph capture
        from .device import new_event, new_stream
        self._side, self._ev_fork, self._ev_join, self._ev_fork2 = new_stream(), new_event(), new_event(), new_event()

    # ------------------------------------------------------------------ buffers ----
    def _ensure(self, B, H, W):
        if B <= self.B and (H, W) == getattr(self, "hw", None):
            return
        if (H, W) != (28, 28):
            raise ValueError("mode='bf16' is instantiated for 28x28 inputs")
        # a RE-allocation (larger batch / other input size) makes every captured CUDA graph stale: Model keys its graphs by this
        self.generation = getattr(self, "generation", 0) + (1 if self.B else 0)
        self.hw, self.B = (H, W), B
        K = 7 * 7 * 64
        if self.de.weights.shape[0] != K:
            raise ValueError(f"Dense expects {self.de.weights.shape[0]} inputs, the conv stack produces {K}")
        self.K = K
        z = DeviceArray.zeros
        self.act1_pad = z((B, 16, 16, 32), np.uint16)       # conv2 input, zero halo kept forever
        self.idx1 = z((B, 14, 14, 16), np.uint8)
        self.act2 = z((B, 7, 7, 64), np.uint16)
        self.idx2 = z((B, 7, 7, 32), np.uint8)
        self.probs = z((B, self.n_out))
        self.dlogits = z((B, self.n_out))
        self.losses = z((B,))
        self.correct = z((B,), np.int32)
        self.dy2_pad = z((B, 16, 16, 64), np.uint16)      # zero halo kept forever
        self.dp1 = z((B, 16, 16, 32), np.uint16)
        self.dummy_labels = z((B,), np.int32)
        self.dl_hilo = z((B, 32), np.uint16)                 # dlogits as bf16 hi | lo (tcgen05 Dense wgrad operand)
        self.w3p = DeviceArray((self.n_out, K), np.float32)
        self.w3hl = z((32, K), np.uint16)                    # Dense weights as bf16 hi | lo rows (tcgen05 Dense fprop operand)
        ws = max(lib.nm_tc_conv3x3_wgrad_workspace(32, 64), lib.nm_tc_conv3x3_bwd_workspace(32, 64), lib.nm_tc_conv1_bwd_workspace(32),
                 lib.nm_tc_dense_wgrad_workspace(B, K), lib.nm_tc_dense_bwd_unpool_workspace(64))
        self.ws = DeviceArray((ws,), np.uint8)
        self.ws_c1 = DeviceArray((lib.nm_tc_conv1_bwd_workspace(32),), np.uint8)     # conv1's partials: conv2's may still be being reduced
        self.ws_dw = DeviceArray((max(lib.nm_tc_dense_wgrad_workspace(B, K), 256),), np.uint8)   # own workspace: the Dense
        #                                                     wgrad may run on a side branch next to the conv backward kernels
        self.ws_head = z((lib.nm_tc_dense_softmax_ce_workspace(B, K),), np.uint8)   # zero-filled once (ticket counter)

    def pack_weights(self):
        """Refresh the bf16 / permuted operand copies from the FP32 master weights in the arena."""
        s = stream()
        if self.w3p is not None:
            lib.nm_tc_pack_cnn_weights_bf16(self.c1.weights.p, self.w1e.p, self.c2.weights.p, self.w2f.p, self.w2d.p, 64, 32,
                                            self.de.weights.p, self.w3p.p, self.w3hl.p, self.K, self.n_out, 49, 64, s)
        else:
            lib.nm_tc_pack_conv1_weights_bf16(self.c1.weights.p, self.w1e.p, 32, s)
            lib.nm_tc_pack_conv_weights_bf16(self.c2.weights.p, self.w2f.p, self.w2d.p, 64, 32, s)

    # ------------------------------------------------------------------ passes ----
    def forward(self, x: DeviceArray, labels=None, accum=None, inv_batch=None):
        B = x.shape[0]
        if x.ndim != 4 or x.shape[1] != 1:
            raise RuntimeError("Input and kernel must be 4D tensors")
        self._ensure(B, x.shape[2], x.shape[3])
        s = stream()
        self.x = x
        # The master weights may have changed (optimizer, set_parameters, FedAvg): every packed operand copy is refreshed
        # per forward pass -- inside the conv1 kernel (its B tile comes from the FP32 weights, two spare warps re-pack conv2's
        # and the Dense head's), which takes the separate pack launch off the critical path of the step.
        u8 = x.dtype == np.uint8       # raw pixels: normalised by 1 / uint8_denominator inside the patch builder
        if self.fused_pack and self.w3p is not None:
            _timed('tc_conv1_fwd_pool', lib.nm_tc_conv1_fwd_pool_pack_bf16, x.p, int(u8), self._u8_scale() if u8 else 1.0, self.c1.weights.p,
                   self.act1_pad.p, self.idx1.p, B, 28, 28, 32, self.c2.weights.p, self.w2f.p, self.w2d.p, 64, 32,
                   self.de.weights.p, self.w3p.p, self.w3hl.p, self.K, self.n_out, 49, 64, s)
        else:
            self.pack_weights()
            if u8:
                _timed('tc_conv1_fwd_pool', lib.nm_tc_conv1_fwd_pool_u8_bf16, x.p, self._u8_scale(), self.w1e.p, self.act1_pad.p, self.idx1.p, B, 28, 28, 32, s)
            else:
                _timed('tc_conv1_fwd_pool', lib.nm_tc_conv1_fwd_pool_bf16, x.p, self.w1e.p, self.act1_pad.p, self.idx1.p, B, 28, 28, 32, s)
        _timed('tc_conv2_fprop_pool', lib.nm_tc_conv3x3_fprop_pool_bf16, self.act1_pad.p, self.w2f.p, self.act2.p, self.idx2.p, B, 14, 14, 32, 64, 0, s)
        lab = labels if labels is not None else self.dummy_labels
        _timed('tc_dense_softmax_ce', lib.nm_tc_dense_softmax_ce_bf16, self.act2.p, self.w3hl.p, self.de.biases.p, lab.p,
               self.probs.p, self.dlogits.p, self.dl_hilo.p, self.losses.p, self.correct.p,
               accum.p if accum is not None else None, self.ws_head.p, self.ws_head.nbytes,
               B, self.K, self.n_out, float(1.0 / B if inv_batch is None else inv_batch), s)
        self.dl_hilo_valid = labels is not None
        out = self.probs[0:B]
        # reference-style attribute views (host conversion on demand)
        self.p1.output = Bf16Activation(self.act1_pad, B, 1)
        self.p2.output = Bf16Activation(self.act2, B, 0)
        self.fl.output = self.p2.output
        self.sm.output = out
        self.model.softmax_classifier_output.dinputs = self.dlogits[0:B]
        return out

    def backward(self, grad_batch=None, comm=None, fork=False, dp=None):
        """``comm``: data-parallel communicator.  The gradient slab is all-reduced in two buckets on the communication
        stream: everything but the first conv's gradients as soon as conv2's wgrad has been enqueued (it then overlaps
        conv2 dgrad + the whole conv1 backward), the first conv's 320 floats after the last kernel.
        ``fork``: put the Dense wgrad on a side stream (a sibling branch when the step is captured into a CUDA graph): it
        only needs the head's outputs, so it fills the SMs that the unpool / conv backward kernels leave idle in their tails
        instead of sitting on the critical path; the streams join before the optimizer.
        ``dp``: communicator with the peer-memory exchange (nm_dp_*): every gradient but the first conv's is PUSHED to the
        peers as soon as conv2's wgrad has been reduced -- on the side branch when forked -- so the transfer overlaps the
        whole conv1 backward and the fused collect + Adam kernel finds it delivered; ``self.pushed_from`` tells it which
        slab chunks are already on their way."""
        self.pushed_from = None
        B = self.x.shape[0]
        s = stream()
        scale = 1.0 / (B if grad_batch is None else grad_batch)
        if not getattr(self, "dl_hilo_valid", False):      # dlogits were written by the public-API backward
            lib.nm_tc_pack_dlogits_bf16(self.dlogits.p, self.dl_hilo.p, B, self.n_out, s)
        sd = s
        if fork:
            sd = self._side
            lib.nm_event_record(self._ev_fork, s)
            lib.nm_stream_wait_event(sd, self._ev_fork)
        _timed('tc_dense_wgrad', lib.nm_tc_dense_wgrad_bf16, self.act2.p, self.dl_hilo.p, self.dlogits.p, self.de.weights.p, self.de.dweights.p, self.de.dbiases.p,
                                   self.ws_dw.p, self.ws_dw.nbytes, B, self.K, self.n_out, 49, 64, float(scale),
                                   float(self.de.weight_regularizer_l1), float(self.de.weight_regularizer_l2), sd)
        if fork:
            lib.nm_event_record(self._ev_join, sd)
        # Dense dgrad + pool backward + ReLU' -> conv2's dY (+ conv2 bias gradient), then conv2 dgrad + wgrad in one pass
        # over dY.  (nm_tc_dense_bwd_pool_bf16 + nm_tc_conv3x3_bwd_pool_bf16 keep dY out of HBM altogether by rebuilding
        # the tile in shared memory, but measured slower: 23 + 100 us against 38 + 73 us -- the builder warps' generic
        # shared-memory stores starve behind the tensor core's operand reads, which saturate the 128 B/cycle port.)
        _timed('tc_dense_bwd_unpool', lib.nm_tc_dense_bwd_unpool_bf16, self.dlogits.p, self.w3p.p, self.idx2.p, self.dy2_pad.p,
                                        self.c2.bias_gradients.p, self.ws.p, self.ws.nbytes, B, 7, 7, 64, self.n_out, s)
        if fork and comm is None:
            # conv2's wgrad partials are reduced on the side branch while conv1's backward already consumes dp1
            _timed('tc_conv2_bwd', lib.nm_tc_conv3x3_bwd_bf16, self.act1_pad.p, self.dy2_pad.p, self.w2d.p, None,
                   self.dp1.p, self.ws.p, self.ws.nbytes, B, 14, 14, 32, 64, s)
            lib.nm_event_record(self._ev_fork2, s)
            lib.nm_stream_wait_event(sd, self._ev_fork2)
            lib.nm_tc_conv3x3_bwd_reduce_f32(self.ws.p, self.c2.weight_gradients.p, B, 14, 14, 32, 64, sd)
            self._early_push(dp, sd)
            lib.nm_event_record(self._ev_join, sd)
        else:
            _timed('tc_conv2_bwd', lib.nm_tc_conv3x3_bwd_bf16, self.act1_pad.p, self.dy2_pad.p, self.w2d.p, self.c2.weight_gradients.p,
                   self.dp1.p, self.ws.p, self.ws.nbytes, B, 14, 14, 32, 64, s)
            if dp is not None and fork:         # the Dense gradients were written on the side branch
                lib.nm_stream_wait_event(s, self._ev_join)
            self._early_push(dp, s)
        if comm is not None:
            if fork:        # the Dense gradients of this bucket were written on the side branch: join it first
                lib.nm_stream_wait_event(s, self._ev_join)
            lo = self._first_conv_floats()
            comm.allreduce_async(self.model.arena.grads, lo, self.model.arena.total - lo)
        if self.x.dtype == np.uint8:
            _timed('tc_conv1_bwd', lib.nm_tc_conv1_bwd_u8_bf16, self.x.p, self._u8_scale(), self.dp1.p, self.idx1.p, self.c1.weight_gradients.p,
                   self.c1.bias_gradients.p, self.ws_c1.p, self.ws_c1.nbytes, B, 28, 28, 32, s)
        else:
            _timed('tc_conv1_bwd', lib.nm_tc_conv1_bwd_bf16, self.x.p, self.dp1.p, self.idx1.p, self.c1.weight_gradients.p, self.c1.bias_gradients.p,
                   self.ws_c1.p, self.ws_c1.nbytes, B, 28, 28, 32, s)
        if comm is not None:
            comm.allreduce_async(self.model.arena.grads, 0, self._first_conv_floats())
        if fork:
            lib.nm_stream_wait_event(s, self._ev_join)

    def _early_push(self, dp, st):
        """Deliver every whole slab chunk above the first conv's gradients to the peers now (nm_dp_push_f32)."""
        if dp is None or getattr(dp, "peer", None) is None:
            return
        arena = self.model.arena
        chunk = lib.nm_dp_chunk_floats()
        lo = (self._first_conv_floats() + chunk - 1) // chunk
        hi = (arena.total + chunk - 1) // chunk
        if lo < hi:
            lib.nm_dp_push_f32(dp.peer, arena.grads.p, arena.total, lo, hi, st)
            self.pushed_from = lo

    def _u8_scale(self):
        return float(np.float32(1.0) / np.float32(self.model.uint8_denominator))

    def _first_conv_floats(self):
        """Length (floats, padded) of the first conv's slice at the start of the arena slabs."""
        for layer, _, off, _, _ in self.model.arena.table:
            if layer is not self.c1:
                return off
        return self.model.arena.total
