"""bf16 tensor-core execution plan for multi-layer perceptrons (BASELINE.json config 1, examples/test_Dense_MNIST.py):

    Dense(n_in -> h1, ReLU) [-> Dropout] [-> Dense(h1 -> h2, ReLU) [-> Dropout] ...] -> Dense(h_k -> n <= 16) -> Softmax

with n_in % 8 == 0 and every hidden width a multiple of 64 (the stack of examples/test_Dense_MNIST.py:39-44 without its
doubled softmax).  ``Model.finalize(mode="bf16")`` picks this plan for such
stacks.  Every Dense contraction runs on tcgen05 tensor cores: hidden layers through the general TMA-fed GEMMs of
nm_tc_gemm.cu (fprop with bias + ReLU + the Layer_Dropout mask in the epilogue -- the mask is drawn from the same Philox
stream as the FP32 layer's, keyed by the layer's seed and offset by the optimizer's step count so that a CUDA-graph replay
draws a fresh one --, dgrad with the ReLU backward of the layer below (and the 1 / keep of a dropout in between) in the epilogue,
wgrad with both operands read MN-major from the activation / gradient tensors), the last layer through the Dense head of
nm_tc_dense.cu (logits + softmax + CE + gradient in one launch; hi/lo split weights).  Activations and their gradients
are bf16 [B][features]; parameters, gradients and optimizer state stay FP32 in the arena in the reference's layouts, so the
optimizers, the data-parallel exchange, FedAvg and save/load are shared with the FP32 path.
Numerics: bf16 operands and stored activations, FP32 accumulation (stated tolerance in tests/test_gpu_tc_mlp.py).
Reference arithmetic: Dense.py:64-108 (incl. the second division by the batch), Relu.py:7-14.

``Model.finalize(mode="bf16x3")`` runs the same plan FP32-EXACT on the tensor cores: every hidden-layer operand is three
bf16 planes h + m + l (24 mantissa bits) and each contraction the six plane-pair products above 2^-24 in one FP32 TMEM
accumulator (nm_tc_dense_*_gen_x3); the narrow head (K x n_out <= 16) runs on the FP32 CUDA-core kernels.  Tolerance: the
FP32-exact one, rtol 1e-4 / atol 1e-5 against the NumPy path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib
from .device import DeviceArray, new_event, new_stream, stream
from .tc_plan import _timed, bf16_bits_to_f32


class Bf16Matrix:
    """Host-convertible view of an internal bf16 [B][features] activation: np.asarray(view) is float32 [n][features]."""

    def __init__(self, buf: DeviceArray, n):
        self.buf, self.n = buf, n
        self.shape = (n, buf.shape[1])
        self.dtype = np.dtype(np.float32)
        self.ndim = 2

    def __len__(self):
        return self.n

    def numpy(self):
        return bf16_bits_to_f32(self.buf.numpy()[:self.n])

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


class Split3Matrix:
    """Host-convertible view of an activation kept as three bf16 planes [3][B][features]: h + m + l in float32."""

    def __init__(self, buf: DeviceArray, n):
        self.buf, self.n = buf, n
        self.shape = (n, buf.shape[2])
        self.dtype = np.dtype(np.float32)
        self.ndim = 2

    def __len__(self):
        return self.n

    def numpy(self):
        planes = bf16_bits_to_f32(self.buf.numpy()[:, :self.n])
        return (planes[2] + planes[1]) + planes[0]

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


class TcMlpPlan:
    PATTERN = ("(Dense(n_in % 8 == 0 -> h % 64 == 0, ReLU) [-> Dropout]) x k (k >= 1) -> Dense(-> n <= 16, no activation) "
               "-> Softmax")

    @staticmethod
    def parse(model):
        from .src.layers import Layer_Dense, Layer_Dropout
        from .src.activations import Activation_ReLU, Activation_Softmax
        L = list(model.layers)
        if len(L) < 3 or not isinstance(L[-1], Activation_Softmax):
            return None
        dense, drops = [], {}
        for layer in L[:-1]:
            if isinstance(layer, Layer_Dense):
                dense.append(layer)
            elif isinstance(layer, Layer_Dropout) and dense and id(dense[-1]) not in drops and layer.fixed_mask is None:
                drops[id(dense[-1])] = layer            # a dropout right behind a hidden Dense is fused into its epilogue
            else:
                return None
        if len(dense) < 2 or id(dense[-1]) in drops:
            return None
        hidden, head = dense[:-1], dense[-1]
        if head.activation is not None or head.weights.shape[1] > 16 or head.weights.shape[0] % 8:
            return None
        n_prev = None
        for d in hidden:
            n_in, n_out = d.weights.shape
            if type(d.activation) is not Activation_ReLU or n_out % 64 or n_in % 8 or (n_prev is not None and n_in != n_prev):
                return None
            n_prev = n_out
        if head.weights.shape[0] != n_prev:
            return None
        return hidden, head, L[-1], drops

    @staticmethod
    def match(model):
        return TcMlpPlan.parse(model) is not None

    def __init__(self, model, split=False):
        parsed = self.parse(model)
        mode = "bf16x3" if split else "bf16"
        if parsed is None:
            raise ValueError(f"mode='{mode}' supports the layer pattern: {self.PATTERN}")
        if model.softmax_classifier_output is None:
            raise ValueError(f"mode='{mode}' needs Loss_CategoricalCrossentropy with a final Activation_Softmax")
        self.model = model
        self.split = bool(split)          # FP32-exact: three bf16 planes per operand
        self.hidden, self.de, self.sm, self.drops = parsed
        self.n_in = self.hidden[0].weights.shape[0]
        self.K = self.de.weights.shape[0]
        self.n_out = self.de.weights.shape[1]
        self.B = 0
        self.packed = {}
        pl = (3,) if self.split else ()
        for d in self.hidden:
            n_in, n_out = d.weights.shape
            self.packed[id(d)] = (DeviceArray(pl + (n_in, n_out), np.uint16), DeviceArray(pl + (n_out, n_in), np.uint16))
        if not self.split:
            self.w3p = DeviceArray((self.n_out, self.K), np.float32)
            self.w3hl = DeviceArray.zeros((32, self.K), np.uint16)
        self._side, self._ev_fork, self._ev_join = new_stream(), new_event(), new_event()
        self.pushed_from = None

    # ------------------------------------------------------------------ buffers ----
    def _ensure(self, B, n_in):
        if n_in != self.n_in:
            raise ValueError(f"Input has {n_in} features, the first Dense expects {self.n_in}")
        if B <= self.B:
            return
        # a RE-allocation (larger batch) makes every captured CUDA graph stale: Model keys its graphs by this
        self.generation = getattr(self, "generation", 0) + (1 if self.B else 0)
        self.B = B
        z = DeviceArray.zeros
        pl = (3,) if self.split else ()
        self.x0 = z(pl + (B, self.n_in), np.uint16)
        self.acts = [z(pl + (B, d.weights.shape[1]), np.uint16) for d in self.hidden]
        self.dacts = [z(pl + (B, d.weights.shape[1]), np.uint16) for d in self.hidden]
        self.probs = z((B, self.n_out))
        self.dlogits = z((B, self.n_out))
        self.losses = z((B,))
        self.correct = z((B,), np.int32)
        self.dummy_labels = z((B,), np.int32)
        self.dl_hilo = z((B, 32), np.uint16)
        ws = max(lib.nm_tc_dense_wgrad_gen_workspace(B, *d.weights.shape) for d in self.hidden)
        self.ws = DeviceArray((max(ws, 256),), np.uint8)
        if self.split:        # the head on the FP32 CUDA-core kernels: FP32 copy of its input, logits, dX
            self.last_f32 = z((B, self.K))
            self.logits = z((B, self.n_out))
            self.dx_head = z((B, self.K))
            self.ws_dw = DeviceArray((max(lib.nm_dense_wgrad_f32_workspace(B, self.K, self.n_out), 256),), np.uint8)
            return
        self.ws_dw = DeviceArray((max(lib.nm_tc_dense_wgrad_workspace(B, self.K), 256),), np.uint8)
        self.ws_head = z((lib.nm_tc_dense_softmax_ce_workspace(B, self.K),), np.uint8)     # zero-filled once (ticket counter)

    def pack_weights(self):
        """Refresh the bf16 operand copies from the FP32 master weights in the arena."""
        s = stream()
        pack = lib.nm_tc_pack_dense_gen_x3 if self.split else lib.nm_tc_pack_dense_gen_bf16
        for d in self.hidden:
            w_kn, w_nk = self.packed[id(d)]
            pack(d.weights.p, w_kn.p, w_nk.p, d.weights.shape[0], d.weights.shape[1], s)
        if not self.split:
            lib.nm_tc_pack_dense_weights(self.de.weights.p, self.w3p.p, self.w3hl.p, self.K, self.n_out, 1, self.K, s)

    # ------------------------------------------------------------------ passes ----
    def _dropout_args(self, d, training):
        """(keep, seed, host offset, device offset pointer) of the Layer_Dropout fused behind hidden layer ``d``.  The stream
        offset is the optimizer's step count -- read on the DEVICE when the optimizer keeps it there (Adam: the captured
        step then draws a fresh mask at every replay), from the host attribute otherwise."""
        drop = self.drops.get(id(d))
        if drop is None or not training:
            return 1.0, 0, 0, None
        opt = self.model.optimizer
        if hasattr(opt, "sync_device_state"):
            st = opt.sync_device_state()
            return float(drop.rate), drop.seed, 0, C.c_void_p(st.ptr + 32)          # nm_adam_state.iterations
        return float(drop.rate), drop.seed, int(opt.iterations), None

    def forward(self, x: DeviceArray, labels=None, accum=None, inv_batch=None, training=None):
        training = (labels is not None) if training is None else training
        if x.ndim != 2:
            raise ValueError("The input must be a 2D array (batch, features)")            # Dense.py:73-78
        B = x.shape[0]
        self._ensure(B, x.shape[1])
        s = stream()
        self.x = x
        self.pack_weights()          # master weights may have changed (optimizer, set_parameters, FedAvg)
        if self.split:
            return self._forward_x3(x, labels, accum, inv_batch, training)
        if x.dtype == np.uint8:
            lib.nm_tc_u8_to_bf16(x.p, float(np.float32(1.0) / np.float32(self.model.uint8_denominator)), self.x0.p, x.size, s)
        else:
            lib.nm_tc_to_bf16(x.p, self.x0.p, x.size, s)
        cur = self.x0
        for d, act in zip(self.hidden, self.acts):
            n_in, n_out = d.weights.shape
            keep, seed, off, off_dev = self._dropout_args(d, training)
            _timed(f"tc_dense_gen_fprop[{n_in}->{n_out}]", lib.nm_tc_dense_fprop_gen_bf16, cur.p, self.packed[id(d)][1].p, d.biases.p, act.p,
                   B, n_in, n_out, 1, keep, seed, off, off_dev, s)
            d.output = Bf16Matrix(act, B)         # with a fused dropout: the DROPPED activation (what the next layer reads)
            if id(d) in self.drops:
                self.drops[id(d)].output = d.output
            cur = act
        self._trained_forward = training
        lab = labels if labels is not None else self.dummy_labels
        _timed('tc_dense_softmax_ce', lib.nm_tc_dense_softmax_ce_bf16, cur.p, self.w3hl.p, self.de.biases.p, lab.p,
               self.probs.p, self.dlogits.p, self.dl_hilo.p, self.losses.p, self.correct.p,
               accum.p if accum is not None else None, self.ws_head.p, self.ws_head.nbytes,
               B, self.K, self.n_out, float(1.0 / B if inv_batch is None else inv_batch), s)
        self.dl_hilo_valid = labels is not None
        out = self.probs[0:B]
        self.sm.output = out
        self.model.softmax_classifier_output.dinputs = self.dlogits[0:B]
        return out

    def _forward_x3(self, x, labels, accum, inv_batch, training):
        """FP32-exact forward: hidden layers on planes (nm_tc_dense_fprop_gen_x3), the head on the FP32 CUDA-core kernels."""
        B = x.shape[0]
        s = stream()
        cap = self.B                                   # rows of a plane (the buffers' capacity)
        if x.dtype == np.uint8:                        # the division Layer_Input's FP32 path performs
            lib.nm_tc_split3_u8(x.p, float(self.model.uint8_denominator), self.x0.p, x.size, cap * self.n_in, s)
        else:
            lib.nm_tc_split3_f32(x.p, None, 1.0, self.x0.p, x.size, cap * self.n_in, s)
        cur, cur_n = self.x0, self.n_in
        for i, (d, act) in enumerate(zip(self.hidden, self.acts)):
            n_in, n_out = d.weights.shape
            keep, seed, off, off_dev = self._dropout_args(d, training)
            last = i == len(self.hidden) - 1
            _timed(f"tc_dense_gen_fprop_x3[{n_in}->{n_out}]", lib.nm_tc_dense_fprop_gen_x3, cur.p, cap * cur_n, self.packed[id(d)][1].p,
                   d.biases.p, act.p, cap * n_out, self.last_f32.p if last else None, B, n_in, n_out, 1, keep, seed, off, off_dev, s)
            d.output = Split3Matrix(act, B)
            if id(d) in self.drops:
                self.drops[id(d)].output = d.output
            cur, cur_n = act, n_out
        self._trained_forward = training
        lib.nm_dense_fprop_f32(self.last_f32.p, self.de.weights.p, self.de.biases.p, self.logits.p, B, self.K, self.n_out, 0, s)
        lib.nm_softmax_fwd_f32(self.logits.p, self.probs.p, B, self.n_out, s)
        if labels is not None:
            lib.nm_softmax_ce_f32(self.probs.p, labels.p, self.dlogits.p, self.losses.p, self.correct.p,
                                  accum.p if accum is not None else None, B, self.n_out,
                                  float(1.0 / B if inv_batch is None else inv_batch), s)
        self.dl_hilo_valid = True                      # no hi/lo copy in this mode: the backward reads dlogits itself
        self.de.output = self.logits[0:B]
        out = self.probs[0:B]
        self.sm.output = out
        self.model.softmax_classifier_output.dinputs = self.dlogits[0:B]
        return out

    def _backward_x3(self, grad_batch, comm):
        B = self.x.shape[0]
        s = stream()
        cap = self.B
        scale = 1.0 / (B if grad_batch is None else grad_batch)

        def mscale(d):
            drop = self.drops.get(id(d))
            return 1.0 / float(drop.rate) if (drop is not None and getattr(self, "_trained_forward", False)) else 1.0
        de = self.de
        lib.nm_dense_wgrad_f32(self.last_f32.p, self.dlogits.p, de.weights.p, de.dweights.p, de.dbiases.p, self.ws_dw.p, self.ws_dw.nbytes,
                               B, self.K, self.n_out, float(scale), float(de.weight_regularizer_l1), float(de.weight_regularizer_l2), s)
        lib.nm_dense_dgrad_f32(self.dlogits.p, de.weights.p, self.dx_head.p, B, self.K, self.n_out, s)
        # ReLU backward of the last hidden layer (and 1 / keep of a dropout behind it) while splitting into planes
        lib.nm_tc_split3_f32(self.dx_head.p, self.acts[-1].p, mscale(self.hidden[-1]), self.dacts[-1].p, B * self.K, cap * self.K, s)
        for i in range(len(self.hidden) - 1, -1, -1):
            d = self.hidden[i]
            n_in, n_out = d.weights.shape
            x_in = self.acts[i - 1] if i > 0 else self.x0
            _timed(f"tc_dense_gen_wgrad_x3[{n_in}->{n_out}]", lib.nm_tc_dense_wgrad_gen_x3, x_in.p, cap * n_in, self.dacts[i].p, cap * n_out,
                   d.weights.p, d.dweights.p, d.dbiases.p, self.ws.p, self.ws.nbytes, B, n_in, n_out, float(scale),
                   float(d.weight_regularizer_l1), float(d.weight_regularizer_l2), s)
            if i > 0:
                _timed(f"tc_dense_gen_dgrad_x3[{n_in}->{n_out}]", lib.nm_tc_dense_dgrad_gen_x3, self.dacts[i].p, cap * n_out,
                       self.packed[id(d)][0].p, self.acts[i - 1].p, mscale(self.hidden[i - 1]), self.dacts[i - 1].p, cap * n_in,
                       B, n_in, n_out, s)
        if comm is not None:
            comm.allreduce_async(self.model.arena.grads, 0, self.model.arena.total)

    def backward(self, grad_batch=None, comm=None, fork=False, dp=None):
        """``comm``: data-parallel communicator (NCCL path) -- the gradient slab is all-reduced after the last kernel.
        ``dp`` / ``fork`` are accepted for interface parity with the conv plans (the fused peer exchange pushes the whole
        slab itself: an MLP's backward is a handful of short kernels)."""
        B = self.x.shape[0]
        s = stream()
        self.pushed_from = None
        if self.split:
            return self._backward_x3(grad_batch, comm)
        scale = 1.0 / (B if grad_batch is None else grad_batch)
        if not getattr(self, "dl_hilo_valid", False):      # dlogits were written by the public-API backward
            lib.nm_tc_pack_dlogits_bf16(self.dlogits.p, self.dl_hilo.p, B, self.n_out, s)
        last = self.acts[-1]
        _timed('tc_dense_wgrad', lib.nm_tc_dense_wgrad_bf16, last.p, self.dl_hilo.p, self.dlogits.p, self.de.weights.p,
               self.de.dweights.p, self.de.dbiases.p, self.ws_dw.p, self.ws_dw.nbytes, B, self.K, self.n_out, 1, self.K,
               float(scale), float(self.de.weight_regularizer_l1), float(self.de.weight_regularizer_l2), s)
        # dX of the head, with the ReLU backward of the last hidden layer -> its dY
        def mscale(d):          # surviving elements of a fused dropout carry 1 / keep (Dropout.py:40: dinputs = dvalues * mask)
            drop = self.drops.get(id(d))
            return 1.0 / float(drop.rate) if (drop is not None and getattr(self, "_trained_forward", False)) else 1.0
        _timed('tc_dense_dgrad', lib.nm_tc_dense_dgrad_mask_bf16, self.dlogits.p, self.w3p.p, last.p, mscale(self.hidden[-1]),
               self.dacts[-1].p, B, self.K, self.n_out, s)
        for i in range(len(self.hidden) - 1, -1, -1):
            d = self.hidden[i]
            n_in, n_out = d.weights.shape
            x_in = self.acts[i - 1] if i > 0 else self.x0
            _timed(f"tc_dense_gen_wgrad[{n_in}->{n_out}]", lib.nm_tc_dense_wgrad_gen_bf16, x_in.p, self.dacts[i].p, d.weights.p, d.dweights.p,
                   d.dbiases.p, self.ws.p, self.ws.nbytes, B, n_in, n_out, float(scale), float(d.weight_regularizer_l1),
                   float(d.weight_regularizer_l2), s)
            if i > 0:      # dX + the ReLU backward of the layer below -> its dY
                _timed(f"tc_dense_gen_dgrad[{n_in}->{n_out}]", lib.nm_tc_dense_dgrad_gen_bf16, self.dacts[i].p, self.packed[id(d)][0].p,
                       self.acts[i - 1].p, mscale(self.hidden[i - 1]), self.dacts[i - 1].p, B, n_in, n_out, s)
        if comm is not None:
            comm.allreduce_async(self.model.arena.grads, 0, self.model.arena.total)
