"""bf16 tensor-core execution plan for a general VGG-style stack (BASELINE.json config 4):

    ( Conv2D(3x3, same, ReLU) x k  ->  MaxPooling2D(2, 2) ) x m  ->  Flatten  ->  Dense(-> n <= 16)  ->  Softmax

The last conv's out_channels must be a multiple of 64 (the Dense head reads its pooled output unpadded); every other conv may
have ANY width: its filters are zero-padded to the next multiple of 64 (the padded channels stay zero through ReLU, pool,
dgrad and are never written to the parameter gradients), the first conv's in_channels to 32 or a multiple of 64.
``Model.finalize(mode="bf16")`` picks this plan when the net is not the MNIST pattern of tc_plan.py.

Every contraction runs on tcgen05 tensor cores: the convs through the streamed-filter implicit GEMM of
nm_tc_conv_gen.cu (fprop + ReLU; dgrad + the ReLU backward of the layer below; wgrad + bias gradient), the Dense head
through nm_tc_dense.cu (logits + softmax + CE + gradient in one launch; wgrad).  Pool forward / backward and the Dense
input gradient are HBM-bound CUDA-core kernels (nm_tc_pool.cu).  Activations are bf16 NHWC with a zero halo;
parameters, gradients and optimizer state stay FP32 in the arena in the reference's layouts, so the optimizers, the
data-parallel exchange, FedAvg and save/load are shared with the FP32 path.
Numerics: bf16 operands and stored activations, FP32 accumulation (stated tolerance in tests/test_gpu_vgg_tc_plan.py).

``Model.finalize(mode="bf16x3")`` runs the same plan FP32-EXACT on the tensor cores: every activation, gradient and packed
weight tensor is three bf16 planes h + m + l (24 mantissa bits), each conv contraction the six plane-pair products above
2^-24 in one FP32 TMEM accumulator (nm_tc_conv3x3_gen_x3 / nm_tc_conv3x3_wgrad_gen_x3), the pool compares the exact FP32
values, and the narrow Dense head runs on the FP32 CUDA-core kernels.  Tolerance: the FP32-exact one (rtol 1e-4 / atol 1e-5
against the NumPy path, tests/test_gpu_vgg_tc_x3_plan.py).
"""
from __future__ import annotations

import numpy as np

import ctypes as C

from ._lib import lib
from .device import DeviceArray, stream
from .tc_plan import Bf16Activation, _timed, bf16_bits_to_f32


class Split3Activation:
    """Host-convertible view of an activation kept as three bf16 NHWC planes [3][B][Hp][Wp][C]: np.asarray(view) yields the
    reference's float32 NCHW tensor (h + m + l)."""

    def __init__(self, buf: DeviceArray, n, halo, channels=None):
        self.buf, self.n, self.halo = buf, n, halo
        _, _, hp, wp, c = buf.shape
        self.channels = c if channels is None else channels
        self.shape = (n, self.channels, hp - 2 * halo, wp - 2 * halo)
        self.dtype = np.dtype(np.float32)
        self.ndim = 4

    def __len__(self):
        return self.n

    def numpy(self):
        a = bf16_bits_to_f32(self.buf.numpy()[:, :self.n])
        a = (a[2] + a[1]) + a[0]
        if self.halo:
            a = a[:, self.halo:-self.halo, self.halo:-self.halo, :]
        return np.ascontiguousarray(a[..., :self.channels].transpose(0, 3, 1, 2))

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)


def _rup(a, b=64):
    return (a + b - 1) // b * b


def _rup_in(c):
    """Padded INPUT channel count of a conv: multiples of 64; a narrow first layer (RGB) pads to 32 only -- the general
    conv has a 32-channel K-chunk variant (64-byte rows), which halves the bytes of the input pack and of the first conv."""
    return 32 if c <= 32 else _rup(c)


class TcStackPlan:
    PATTERN = ("(Conv2D(3x3,same,ReLU) x k -> MaxPooling2D(2,2)) x m -> Flatten -> Dense(-> n <= 16) -> Softmax "
               "(mode 'bf16': out_channels of the last conv % 64 == 0; any width in mode 'bf16x3')")

    # ------------------------------------------------------------------ pattern ----
    @staticmethod
    def parse(model, split=False):
        """-> (blocks, flatten, dense, softmax) with blocks = [([conv, ...], pool), ...], or None.  ``split`` (mode "bf16x3"):
        the head runs on FP32 kernels behind a transpose that drops padded channels, so the last conv may be narrow too."""
        from .src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten
        from .src.activations import Activation_ReLU, Activation_Softmax
        L = list(model.layers)
        if len(L) < 5 or not (isinstance(L[-1], Activation_Softmax) and isinstance(L[-2], Layer_Dense)
                              and isinstance(L[-3], Layer_Flatten)):
            return None
        blocks, convs = [], []
        for layer in L[:-3]:
            if isinstance(layer, Layer_Conv2D):
                oc, _, kh, kw = layer.weights.shape
                if (kh, kw) != (3, 3) or layer.padding != "same" or type(layer.activation) is not Activation_ReLU:
                    return None
                if tuple(np.atleast_1d(layer.stride)) not in ((1,), (1, 1)):
                    return None
                convs.append(layer)
            elif isinstance(layer, Layer_MaxPooling2D):
                if not convs or layer.pool_size != (2, 2) or layer.strides != (2, 2) or layer.padding != "valid":
                    return None
                blocks.append((convs, layer))
                convs = []
            else:
                return None
        if convs or not blocks or (blocks[-1][0][-1].weights.shape[0] % 64 and not split):
            return None
        de = L[-2]
        if de.activation is not None or de.weights.shape[1] > 16:
            return None
        ic = None
        for cs, _ in blocks:
            for c in cs:
                if ic is not None and c.weights.shape[1] != ic:
                    return None
                ic = c.weights.shape[0]
        return blocks, L[-3], de, L[-1]

    @staticmethod
    def match(model, split=False):
        return TcStackPlan.parse(model, split) is not None

    def __init__(self, model, split=False):
        parsed = self.parse(model, split)
        mode = "bf16x3" if split else "bf16"
        if parsed is None:
            raise ValueError(f"mode='{mode}' supports the layer patterns: {self.PATTERN}")
        if model.softmax_classifier_output is None:
            raise ValueError(f"mode='{mode}' needs Loss_CategoricalCrossentropy with a final Activation_Softmax")
        from . import config
        if config.conv_backward != "wired" or config.conv_bias_in_forward:
            raise ValueError(f"mode='{mode}' implements the default quirk settings only")
        self.model = model
        self.split = bool(split)          # FP32-exact: three bf16 planes per tensor
        self._pl = (3,) if self.split else ()
        self.blocks, self.fl, self.de, self.sm = parsed
        self.convs = [c for cs, _ in self.blocks for c in cs]
        self.n_out = self.de.weights.shape[1]
        self.B = 0
        self.hw = None
        self.packed = {}
        prev_ocp = None
        for c in self.convs:
            oc, ic = c.weights.shape[:2]
            # a conv reads the PADDED channels of the tensor below it (zero weights for the padding)
            ocp, icp = _rup(oc), (_rup_in(ic) if prev_ocp is None else prev_ocp)
            prev_ocp = ocp
            self.packed[id(c)] = (DeviceArray(self._pl + (ocp, 9, icp), np.uint16), DeviceArray(self._pl + (icp, 9, ocp), np.uint16), ocp, icp)
        # side stream for the data-parallel early pushes (created here, never inside a CUDA-graph capture)
        from .device import new_event, new_stream
        self._side, self._ev_fork, self._ev_join = new_stream(), new_event(), new_event()
        self.pushed_from = None

    # ------------------------------------------------------------------ buffers ----
    def _ensure(self, B, C, H, W):
        if B <= self.B and (C, H, W) == self.hw:
            return
        if C != self.convs[0].weights.shape[1]:
            raise RuntimeError(f"Input has {C} channels, the first Conv2D expects {self.convs[0].weights.shape[1]}")
        # a RE-allocation (larger batch / other input size) makes every captured CUDA graph stale: Model keys its graphs by this
        self.generation = getattr(self, "generation", 0) + (1 if self.B else 0)
        self.hw, self.B = (C, H, W), B
        pl = self._pl

        def z(shape, dtype=np.float32):
            return DeviceArray.zeros((pl + tuple(shape)) if dtype == np.uint16 else shape, dtype)      # bf16 tensors: one or three planes
        self.x_pad = z((B, H + 2, W + 2, _rup_in(C)), np.uint16)
        self.steps = []                     # per block: dict(convs=[(layer, in_buf, act, dy, h, w)], pool=..., ...)
        cur, h, w = self.x_pad, H, W
        ws = 256
        for bi, (cs, pool) in enumerate(self.blocks):
            if h % 2 or w % 2:
                raise ValueError("mode='bf16' needs even feature-map sizes at every MaxPooling2D(2, 2)")
            entries = []
            for c in cs:
                _, _, ocp, icp = self.packed[id(c)]
                act = z((B, h + 2, w + 2, ocp), np.uint16)
                dy = z((B, h + 2, w + 2, ocp), np.uint16)
                entries.append(dict(layer=c, inp=cur, act=act, dy=dy, h=h, w=w, ocp=ocp, icp=icp,
                                    oc=c.weights.shape[0], ic=c.weights.shape[1]))
                ws = max(ws, lib.nm_tc_conv3x3_wgrad_gen_workspace(B, h, w, c.weights.shape[1], icp, ocp))
                cur = act
            last = bi == len(self.blocks) - 1
            ch = entries[-1]["ocp"]
            shape = (B, h // 2, w // 2, ch) if last else (B, h // 2 + 2, w // 2 + 2, ch)
            pooled, dpool = z(shape, np.uint16), z(shape, np.uint16)
            idx = z((B, h // 2, w // 2, ch // 2), np.uint8)
            self.steps.append(dict(convs=entries, pool=pool, pooled=pooled, dpool=dpool, idx=idx, h=h, w=w, ch=ch, padded=0 if last else 1))
            cur, h, w = pooled, h // 2, w // 2
        self.HW, self.Cl = h * w, self.steps[-1]["ch"]
        self.Cl_real = self.convs[-1].weights.shape[0]            # < Cl only in mode "bf16x3" (zero-padded last conv)
        K = self.HW * self.Cl_real
        if self.de.weights.shape[0] != K:
            raise ValueError(f"Dense expects {self.de.weights.shape[0]} inputs, the conv stack produces {K}")
        self.K = K
        self.probs = z((B, self.n_out))
        self.dlogits = z((B, self.n_out))
        self.losses = z((B,))
        self.correct = z((B,), np.int32)
        self.dummy_labels = z((B,), np.int32)
        self.ws = DeviceArray((ws,), np.uint8)
        if self.split:        # the head on the FP32 CUDA-core kernels: FP32 copy of its input (Flatten's order), logits, dX
            self.last_f32 = z((B, K))
            self.logits = z((B, self.n_out))
            self.dx_head = z((B, K))
            self.ws_dw = DeviceArray((max(lib.nm_dense_wgrad_f32_workspace(B, K, self.n_out), 256),), np.uint8)
            return
        self.dl_hilo = DeviceArray.zeros((B, 32), np.uint16)
        self.w3p = DeviceArray((self.n_out, K), np.float32)
        self.w3hl = DeviceArray.zeros((32, K), np.uint16)
        self.ws_dw = DeviceArray((max(lib.nm_tc_dense_wgrad_workspace(B, K), 256),), np.uint8)
        self.ws_head = z((lib.nm_tc_dense_softmax_ce_workspace(B, K),), np.uint8)     # zero-filled once (ticket counter)

    def pack_weights(self):
        """Refresh the bf16 operand copies from the FP32 master weights in the arena."""
        s = stream()
        pack = lib.nm_tc_pack_conv_weights_gen_x3 if self.split else lib.nm_tc_pack_conv_weights_gen_bf16
        for c in self.convs:
            wf, wd, ocp, icp = self.packed[id(c)]
            oc, ic = c.weights.shape[:2]
            pack(c.weights.p, wf.p, wd.p, oc, ic, ocp, icp, s)
        if not self.split:
            lib.nm_tc_pack_dense_weights(self.de.weights.p, self.w3p.p, self.w3hl.p, self.K, self.n_out, self.HW, self.Cl, s)

    # ------------------------------------------------------------------ passes ----
    def forward(self, x: DeviceArray, labels=None, accum=None, inv_batch=None):
        if x.ndim != 4:
            raise RuntimeError("Input and kernel must be 4D tensors")
        B, C, H, W = x.shape
        self._ensure(B, C, H, W)
        s = stream()
        self.x = x
        self.pack_weights()          # master weights may have changed (optimizer, set_parameters, FedAvg)
        if self.split:
            return self._forward_x3(x, labels, accum, inv_batch)
        if x.dtype == np.uint8:       # raw pixels: normalised by 1 / uint8_denominator while packing
            _timed('tc_pack_input', lib.nm_tc_pack_input_nhwc_u8_bf16, x.p, float(np.float32(1.0) / np.float32(self.model.uint8_denominator)),
                   self.x_pad.p, B, C, H, W, self.x_pad.shape[3], s)
        else:
            _timed('tc_pack_input', lib.nm_tc_pack_input_nhwc_bf16, x.p, self.x_pad.p, B, C, H, W, self.x_pad.shape[3], s)
        for st in self.steps:
            for e in st["convs"]:
                wf = self.packed[id(e["layer"])][0]
                _timed(f"tc_conv_gen_fprop[{e['ic']}->{e['oc']}@{e['h']}]", lib.nm_tc_conv3x3_gen_bf16, e["inp"].p, wf.p, None, None, e["act"].p,
                       B, e["h"], e["w"], e["icp"], e["ocp"], 1, s)
                e["layer"].output = Bf16Activation(e["act"], B, 1, e["oc"])
            _timed('tc_maxpool_fwd', lib.nm_tc_maxpool2x2_fwd_bf16, st["convs"][-1]["act"].p, st["pooled"].p, st["idx"].p,
                   B, st["h"], st["w"], st["ch"], st["padded"], s)
            st["pool"].output = Bf16Activation(st["pooled"], B, st["padded"], st["convs"][-1]["oc"])
        last = self.steps[-1]
        lab = labels if labels is not None else self.dummy_labels
        _timed('tc_dense_softmax_ce', lib.nm_tc_dense_softmax_ce_bf16, last["pooled"].p, self.w3hl.p, self.de.biases.p, lab.p,
               self.probs.p, self.dlogits.p, self.dl_hilo.p, self.losses.p, self.correct.p,
               accum.p if accum is not None else None, self.ws_head.p, self.ws_head.nbytes,
               B, self.K, self.n_out, float(1.0 / B if inv_batch is None else inv_batch), s)
        self.dl_hilo_valid = labels is not None
        out = self.probs[0:B]
        self.fl.output = last["pool"].output
        self.sm.output = out
        self.model.softmax_classifier_output.dinputs = self.dlogits[0:B]
        return out

    # ------------------------------------------------------------------ FP32-exact passes (three planes) ----
    @staticmethod
    def _ps(buf):
        """Elements between the planes of a [3][...] tensor."""
        return buf.size // 3

    @staticmethod
    def _plane(buf, pl):
        return C.c_void_p(buf.ptr + 2 * pl * (buf.size // 3))

    def _forward_x3(self, x, labels, accum, inv_batch):
        B, Cin, H, W = x.shape
        s = stream()
        ps = self._ps
        if x.dtype == np.uint8:       # raw pixels: DIVIDED by uint8_denominator like Layer_Input's FP32 path
            _timed('tc_pack_input_x3', lib.nm_tc_pack_input_nhwc_u8_x3, x.p, float(self.model.uint8_denominator), self.x_pad.p, ps(self.x_pad),
                   B, Cin, H, W, self.x_pad.shape[-1], s)
        else:
            _timed('tc_pack_input_x3', lib.nm_tc_pack_input_nhwc_x3, x.p, self.x_pad.p, ps(self.x_pad), B, Cin, H, W, self.x_pad.shape[-1], s)
        for st in self.steps:
            for e in st["convs"]:
                wf = self.packed[id(e["layer"])][0]
                _timed(f"tc_conv_gen_fprop_x3[{e['ic']}->{e['oc']}@{e['h']}]", lib.nm_tc_conv3x3_gen_x3, e["inp"].p, ps(e["inp"]), wf.p, None, None,
                       e["act"].p, ps(e["act"]), B, e["h"], e["w"], e["icp"], e["ocp"], 1, s)
                e["layer"].output = Split3Activation(e["act"], B, 1, e["oc"])
            last_act = st["convs"][-1]["act"]
            _timed('tc_maxpool_fwd_x3', lib.nm_tc_maxpool2x2_fwd_x3, last_act.p, ps(last_act), st["pooled"].p, ps(st["pooled"]), st["idx"].p,
                   B, st["h"], st["w"], st["ch"], st["padded"], s)
            st["pool"].output = Split3Activation(st["pooled"], B, st["padded"], st["convs"][-1]["oc"])
        last = self.steps[-1]
        # Flatten (Flatten.py:21-53: NCHW order) + the Dense head on the FP32 kernels
        lib.nm_tc_merge3_nhwc_to_nchw_f32(last["pooled"].p, ps(last["pooled"]), self.last_f32.p, B, self.HW, self.Cl_real, self.Cl, s)
        lib.nm_dense_fprop_f32(self.last_f32.p, self.de.weights.p, self.de.biases.p, self.logits.p, B, self.K, self.n_out, 0, s)
        lib.nm_softmax_fwd_f32(self.logits.p, self.probs.p, B, self.n_out, s)
        if labels is not None:
            lib.nm_softmax_ce_f32(self.probs.p, labels.p, self.dlogits.p, self.losses.p, self.correct.p,
                                  accum.p if accum is not None else None, B, self.n_out,
                                  float(1.0 / B if inv_batch is None else inv_batch), s)
        self.dl_hilo_valid = True                      # no hi/lo copy in this mode: the backward reads dlogits itself
        out = self.probs[0:B]
        self.fl.output = self.last_f32[0:B]
        self.de.output = self.logits[0:B]
        self.sm.output = out
        self.model.softmax_classifier_output.dinputs = self.dlogits[0:B]
        return out

    def _backward_x3(self, grad_batch, comm, dp):
        B = self.x.shape[0]
        peer = dp is not None and getattr(dp, "peer", None) is not None
        if peer:
            chunk = lib.nm_dp_chunk_floats()
            total_chunks = hi_chunk = (self.model.arena.total + chunk - 1) // chunk
        s = stream()
        ps, plane = self._ps, self._plane
        scale = 1.0 / (B if grad_batch is None else grad_batch)
        de = self.de
        lib.nm_dense_wgrad_f32(self.last_f32.p, self.dlogits.p, de.weights.p, de.dweights.p, de.dbiases.p, self.ws_dw.p, self.ws_dw.nbytes,
                               B, self.K, self.n_out, float(scale), float(de.weight_regularizer_l1), float(de.weight_regularizer_l2), s)
        if peer:
            hi_chunk = self._push_ready(dp, self.de, hi_chunk)
        lib.nm_dense_dgrad_f32(self.dlogits.p, de.weights.p, self.dx_head.p, B, self.K, self.n_out, s)
        last = self.steps[-1]
        lib.nm_tc_split3_nchw_to_nhwc(self.dx_head.p, last["dpool"].p, ps(last["dpool"]), B, self.HW, self.Cl_real, self.Cl, s)
        first = self.convs[0]
        for bi in range(len(self.steps) - 1, -1, -1):
            st = self.steps[bi]
            cs = st["convs"]
            for pl in range(3):       # the routing is a masked copy: the bf16 kernel once per plane
                _timed('tc_maxpool_bwd', lib.nm_tc_maxpool2x2_bwd_bf16, plane(st["dpool"], pl), st["idx"].p, plane(cs[-1]["dy"], pl),
                       B, st["h"], st["w"], st["ch"], st["padded"], s)
            for j in range(len(cs) - 1, -1, -1):
                e = cs[j]
                c = e["layer"]
                oc, ic = c.weights.shape[:2]
                _timed(f"tc_conv_gen_wgrad_x3[{ic}->{oc}@{e['h']}]", lib.nm_tc_conv3x3_wgrad_gen_x3, e["inp"].p, ps(e["inp"]), e["dy"].p, ps(e["dy"]),
                       c.weight_gradients.p, c.bias_gradients.p, self.ws.p, self.ws.nbytes, B, e["h"], e["w"], ic, oc, e["icp"], e["ocp"], s)
                if c is first:
                    continue
                if peer:
                    hi_chunk = self._push_ready(dp, c, hi_chunk)
                wd = self.packed[id(c)][1]
                if j > 0:
                    dst, mask = cs[j - 1]["dy"], cs[j - 1]["act"]
                else:
                    dst, mask = self.steps[bi - 1]["dpool"], None
                _timed(f"tc_conv_gen_dgrad_x3[{ic}->{oc}@{e['h']}]", lib.nm_tc_conv3x3_gen_x3, e["dy"].p, ps(e["dy"]), wd.p, None,
                       mask.p if mask is not None else None, dst.p, ps(dst), B, e["h"], e["w"], e["ocp"], e["icp"], 0, s)
        if peer and hi_chunk < total_chunks:
            lib.nm_event_record(self._ev_join, self._side)
            lib.nm_stream_wait_event(s, self._ev_join)
            self.pushed_from = hi_chunk
        if comm is not None:
            comm.allreduce_async(self.model.arena.grads, 0, self.model.arena.total)

    def _push_ready(self, dp, layer, hi_chunk):
        """All gradients at slab offsets >= ``layer``'s are final: deliver the whole chunks among them that have not been
        pushed yet to the peers, on the side stream (nm_dp_push_f32: stores over NVLink, no waits), so the exchange
        overlaps the rest of the backward pass.  Returns the new lower bound of the pushed range."""
        arena = self.model.arena
        chunk = lib.nm_dp_chunk_floats()
        off = min(o for (l, _, o, _, _) in arena.table if l is layer)
        lo = (off + chunk - 1) // chunk
        if lo >= hi_chunk:
            return hi_chunk
        s = stream()
        lib.nm_event_record(self._ev_fork, s)
        lib.nm_stream_wait_event(self._side, self._ev_fork)
        lib.nm_dp_push_f32(dp.peer, arena.grads.p, arena.total, lo, hi_chunk, self._side)
        return lo

    def backward(self, grad_batch=None, comm=None, fork=False, dp=None):
        """``comm``: data-parallel communicator (NCCL path) -- the gradient slab is all-reduced after the last kernel.
        ``dp``: communicator with the peer-memory exchange: each layer's gradients are pushed to the peers right behind
        its wgrad (``_push_ready``); ``self.pushed_from`` tells the fused collect + Adam kernel what is left to push.
        ``fork`` is accepted for interface parity with TcCnnPlan (the compute chain is not forked here)."""
        B = self.x.shape[0]
        peer = dp is not None and getattr(dp, "peer", None) is not None
        self.pushed_from = None
        if self.split:
            return self._backward_x3(grad_batch, comm, dp)
        if peer:
            chunk = lib.nm_dp_chunk_floats()
            total_chunks = hi_chunk = (self.model.arena.total + chunk - 1) // chunk
        s = stream()
        scale = 1.0 / (B if grad_batch is None else grad_batch)
        if not getattr(self, "dl_hilo_valid", False):      # dlogits were written by the public-API backward
            lib.nm_tc_pack_dlogits_bf16(self.dlogits.p, self.dl_hilo.p, B, self.n_out, s)
        last = self.steps[-1]
        _timed('tc_dense_wgrad', lib.nm_tc_dense_wgrad_bf16, last["pooled"].p, self.dl_hilo.p, self.dlogits.p, self.de.weights.p,
               self.de.dweights.p, self.de.dbiases.p, self.ws_dw.p, self.ws_dw.nbytes, B, self.K, self.n_out, self.HW, self.Cl,
               float(scale), float(self.de.weight_regularizer_l1), float(self.de.weight_regularizer_l2), s)
        if peer:
            hi_chunk = self._push_ready(dp, self.de, hi_chunk)
        _timed('tc_dense_dgrad', lib.nm_tc_dense_dgrad_bf16, self.dlogits.p, self.w3p.p, last["dpool"].p, B, self.K, self.n_out, s)
        first = self.convs[0]
        for bi in range(len(self.steps) - 1, -1, -1):
            st = self.steps[bi]
            cs = st["convs"]
            # pool backward + ReLU backward of the block's last conv -> its dY
            _timed('tc_maxpool_bwd', lib.nm_tc_maxpool2x2_bwd_bf16, st["dpool"].p, st["idx"].p, cs[-1]["dy"].p,
                   B, st["h"], st["w"], st["ch"], st["padded"], s)
            for j in range(len(cs) - 1, -1, -1):
                e = cs[j]
                c = e["layer"]
                oc, ic = c.weights.shape[:2]
                _timed(f"tc_conv_gen_wgrad[{ic}->{oc}@{e['h']}]", lib.nm_tc_conv3x3_wgrad_gen_bf16, e["inp"].p, e["dy"].p, c.weight_gradients.p,
                       c.bias_gradients.p, self.ws.p, self.ws.nbytes, B, e["h"], e["w"], ic, oc, e["icp"], e["ocp"], s)
                if c is first:
                    continue                                # no consumer for the first layer's input gradient
                if peer:
                    hi_chunk = self._push_ready(dp, c, hi_chunk)
                wd = self.packed[id(c)][1]
                if j > 0:       # input = the previous conv's post-ReLU output: dgrad + its ReLU backward -> its dY
                    dst, mask = cs[j - 1]["dy"], cs[j - 1]["act"]
                else:           # input = the previous block's pooled tensor: plain dgrad -> its dpool (haloed)
                    dst, mask = self.steps[bi - 1]["dpool"], None
                _timed(f"tc_conv_gen_dgrad[{ic}->{oc}@{e['h']}]", lib.nm_tc_conv3x3_gen_bf16, e["dy"].p, wd.p, None, mask.p if mask is not None else None,
                       dst.p, B, e["h"], e["w"], e["ocp"], e["icp"], 0, s)
        if peer and hi_chunk < total_chunks:
            lib.nm_event_record(self._ev_join, self._side)
            lib.nm_stream_wait_event(s, self._ev_join)
            self.pushed_from = hi_chunk
        if comm is not None:
            comm.allreduce_async(self.model.arena.grads, 0, self.model.arena.total)
