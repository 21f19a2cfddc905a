"""Python-side operator wrappers: DeviceArray in, DeviceArray out, one C-ABI call each.

These are the calls the layer / loss / optimizer classes make; they only check
shapes and (re)use output buffers -- all arithmetic happens in
libneuromah_b200.so.  FP32-exact path (reference layouts).
"""
from __future__ import annotations

import math

import numpy as np

from ._lib import lib
from .device import DeviceArray, Scratch, stream

_scratch = Scratch()


class KernelTimers:
    """Optional CUDA-event timing of individual operator launches on the library stream
    (bench.py's roofline leg).  Disabled (None) by default: zero overhead on the hot path."""

    def __init__(self):
        import ctypes as C
        self._C = C
        self.pending = []          # (name, start_event, stop_event)
        self.pool = []

    def _event(self):
        if self.pool:
            return self.pool.pop()
        e = self._C.c_void_p()
        lib.nm_event_create(self._C.byref(e))
        return e

    def tic(self, name):
        e0, e1 = self._event(), self._event()
        lib.nm_event_record(e0, stream())
        return (name, e0, e1)

    def toc(self, tok):
        lib.nm_event_record(tok[2], stream())
        self.pending.append(tok)

    def collect(self):
        """{name: [ms, ...]} for everything recorded since the last collect (synchronises)."""
        out = {}
        ms = self._C.c_float()
        for name, e0, e1 in self.pending:
            lib.nm_event_sync(e1)
            lib.nm_event_elapsed_ms(e0, e1, self._C.byref(ms))
            out.setdefault(name, []).append(ms.value)
            self.pool += [e0, e1]
        self.pending = []
        return out


timers: KernelTimers | None = None


def _tic(name):
    return timers.tic(name) if timers is not None else None


def _toc(tok):
    if tok is not None:
        timers.toc(tok)


def _buf(holder, name, shape, dtype=np.float32) -> DeviceArray:
    """Reuse holder.<name> if it already has this shape/dtype, else allocate."""
    cur = holder.__dict__.get(name) if holder is not None else None
    if isinstance(cur, DeviceArray) and cur.shape == tuple(shape) and cur.dtype == np.dtype(dtype):
        return cur
    new = DeviceArray(shape, dtype)
    if holder is not None:
        holder.__dict__[name] = new
    return new


def same_padding(k: int):
    """Conv_wrapepr.py:130-139: odd k -> ((k-1)//2,)*2; even k -> (k//2-1, k//2)."""
    return ((k - 1) // 2, (k - 1) // 2) if k % 2 else (k // 2 - 1, k // 2)


def u8_to_f32(x: DeviceArray, denom: float, holder=None, name="_xf"):
    """float32(x) / denom on the device (the reference scripts' host-side `X.astype(np.float32) / 255.0`)."""
    y = _buf(holder, name, x.shape)
    lib.nm_u8_to_f32_div(x.p, y.p, x.size, float(denom), stream())
    return y


# ------------------------------------------------------------------ conv ----
def conv2d_fprop(x: DeviceArray, w: DeviceArray, bias, pads, relu: bool, out=None, holder=None, name="_y"):
    n, c, h, wd = x.shape
    oc, ic, kh, kw = w.shape
    if ic != c:
        raise RuntimeError("Mismatch between input channels and kernel channels")   # Conv2D.cpp:101-102
    pt, pb, pl, pr = pads
    oh, ow = h + pt + pb - kh + 1, wd + pl + pr - kw + 1
    if oh <= 0 or ow <= 0:
        raise ValueError("kernel larger than the (padded) input")
    y = out if out is not None else _buf(holder, name, (n, oc, oh, ow))
    tok = _tic(f"conv2d_fprop_f32[{c}->{oc}]")
    lib.nm_conv2d_fprop_f32(x.p, w.p, bias.p if bias is not None else None, y.p,
                            n, c, h, wd, oc, kh, kw, pt, pb, pl, pr, int(relu), stream())
    _toc(tok)
    return y


def conv2d_dgrad(dy: DeviceArray, w: DeviceArray, in_shape, pads, holder=None, name="_dx"):
    n, c, h, wd = in_shape
    oc, _, kh, kw = w.shape
    pt, pb, pl, pr = pads
    dx = _buf(holder, name, (n, c, h, wd))
    tok = _tic(f"conv2d_dgrad_f32[{c}->{oc}]")
    lib.nm_conv2d_dgrad_f32(dy.p, w.p, dx.p, n, c, h, wd, oc, kh, kw, pt, pb, pl, pr, stream())
    _toc(tok)
    return dx


def conv2d_wgrad(x: DeviceArray, dy: DeviceArray, dw: DeviceArray, db, pads):
    n, c, h, wd = x.shape
    oc, _, kh, kw = dw.shape
    pt, pb, pl, pr = pads
    need = lib.nm_conv2d_wgrad_f32_workspace(n, c, h, wd, oc, kh, kw, pt, pb, pl, pr)
    ws = _scratch.get(need)
    tok = _tic(f"conv2d_wgrad_f32[{c}->{oc}]")
    lib.nm_conv2d_wgrad_f32(x.p, dy.p, dw.p, db.p if db is not None else None, ws.p, ws.nbytes,
                            n, c, h, wd, oc, kh, kw, pt, pb, pl, pr, stream())
    _toc(tok)
    return dw, db


# ------------------------------------------------------------------ relu ----
def relu_fwd(x: DeviceArray, holder=None, name="_y"):
    y = _buf(holder, name, x.shape)
    lib.nm_relu_fwd_f32(x.p, y.p, x.size, stream())
    return y


def relu_bwd(dy: DeviceArray, ref: DeviceArray, holder=None, name="_drelu"):
    dx = _buf(holder, name, dy.shape)
    lib.nm_relu_bwd_f32(dy.p, ref.p, dx.p, dy.size, stream())
    return dx


# ------------------------------------------------------------------ pool ----
def pool_geometry(h, w, ph, pw, sh, sw, padding):
    """maxpooling_backend.cpp:37-52."""
    if padding == "same":
        oh, ow = int(math.ceil(h / sh)), int(math.ceil(w / sw))
        pad_h = max((oh - 1) * sh + ph - h, 0)
        pad_w = max((ow - 1) * sw + pw - w, 0)
        return oh, ow, pad_h // 2, pad_w // 2
    return (h - ph) // sh + 1, (w - pw) // sw + 1, 0, 0


def maxpool_fwd(x: DeviceArray, ph, pw, sh, sw, padding, holder=None):
    n, c, h, w = x.shape
    oh, ow, pt, pl = pool_geometry(h, w, ph, pw, sh, sw, padding)
    if oh <= 0 or ow <= 0:
        raise ValueError("pool window larger than the input")
    y = _buf(holder, "_y", (n, c, oh, ow))
    idx = _buf(holder, "_idx", (n, c, oh, ow, 2), np.int32)
    lib.nm_maxpool2d_fwd_f32(x.p, y.p, idx.p, n, c, h, w, ph, pw, sh, sw, pt, pl, oh, ow, stream())
    return y, idx


def maxpool_bwd(dy: DeviceArray, idx: DeviceArray, in_shape, ph, pw, sh, sw, padding, holder=None):
    n, c, h, w = in_shape
    oh, ow, pt, pl = pool_geometry(h, w, ph, pw, sh, sw, padding)
    dx = _buf(holder, "_dx", (n, c, h, w))
    lib.nm_maxpool2d_bwd_f32(dy.p, idx.p, dx.p, n, c, h, w, ph, pw, sh, sw, pt, pl, oh, ow, stream())
    return dx


# ----------------------------------------------------------------- dense ----
def dense_fprop(x: DeviceArray, w: DeviceArray, b, relu: bool, holder=None, name="_y"):
    bsz, n_in = x.shape
    n_out = w.shape[1]
    y = _buf(holder, name, (bsz, n_out))
    lib.nm_dense_fprop_f32(x.p, w.p, b.p if b is not None else None, y.p, bsz, n_in, n_out, int(relu), stream())
    return y


def dense_dgrad(dy: DeviceArray, w: DeviceArray, holder=None, name="_dx"):
    bsz, n_out = dy.shape
    n_in = w.shape[0]
    dx = _buf(holder, name, (bsz, n_in))
    lib.nm_dense_dgrad_f32(dy.p, w.p, dx.p, bsz, n_in, n_out, stream())
    return dx


def dense_wgrad(x: DeviceArray, dy: DeviceArray, w: DeviceArray, dw: DeviceArray, db, scale, l1=0.0, l2=0.0):
    bsz, n_in = x.shape
    n_out = dy.shape[1]
    need = lib.nm_dense_wgrad_f32_workspace(bsz, n_in, n_out)
    ws = _scratch.get(need)
    lib.nm_dense_wgrad_f32(x.p, dy.p, w.p, dw.p, db.p if db is not None else None, ws.p, ws.nbytes,
                           bsz, n_in, n_out, float(scale), float(l1), float(l2), stream())
    return dw, db


# ------------------------------------------------------- softmax / losses ----
def softmax_fwd(x: DeviceArray, holder=None, name="_y"):
    y = _buf(holder, name, x.shape)
    lib.nm_softmax_fwd_f32(x.p, y.p, x.shape[0], x.shape[1], stream())
    return y


def softmax_bwd(probs: DeviceArray, dy: DeviceArray, holder=None, name="_dx"):
    dx = _buf(holder, name, dy.shape)
    lib.nm_softmax_bwd_f32(probs.p, dy.p, dx.p, dy.shape[0], dy.shape[1], stream())
    return dx


def softmax_ce(probs: DeviceArray, labels: DeviceArray, *, want_grad, inv_batch=None, accum=None, holder=None):
    """Per-sample CCE losses, argmax==label flags and (optionally) (p-onehot)/B."""
    bsz, n = probs.shape
    if labels.shape != (bsz,):
        raise ValueError(f"labels shape {labels.shape} does not match batch {bsz}")
    losses = _buf(holder, "_sample_losses", (bsz,))
    correct = _buf(holder, "_correct", (bsz,), np.int32)
    dlog = _buf(holder, "_dlogits", (bsz, n)) if want_grad else None
    lib.nm_softmax_ce_f32(probs.p, labels.p, dlog.p if dlog is not None else None, losses.p, correct.p,
                          accum.p if accum is not None else None, bsz, n,
                          float(1.0 / bsz if inv_batch is None else inv_batch), stream())
    return losses, correct, dlog


def cce_bwd(probs: DeviceArray, labels: DeviceArray, holder=None):
    bsz, n = probs.shape
    dx = _buf(holder, "_dx", (bsz, n))
    lib.nm_cce_bwd_f32(probs.p, labels.p, dx.p, bsz, n, float(1.0 / bsz), stream())
    return dx


# ------------------------------------------------ sigmoid / element-wise losses ----
def sigmoid_fwd(x: DeviceArray, holder=None, name="_y"):
    y = _buf(holder, name, x.shape)
    lib.nm_sigmoid_fwd_f32(x.p, y.p, x.size, stream())
    return y


def sigmoid_bwd(dy: DeviceArray, out: DeviceArray, holder=None, name="_dx"):
    dx = _buf(holder, name, dy.shape)
    lib.nm_sigmoid_bwd_f32(dy.p, out.p, dx.p, dy.size, stream())
    return dx


LOSS_MSE, LOSS_MAE, LOSS_BCE = 0, 1, 2        # NM_LOSS_* of include/neuromah_b200.h


def _loss_dims(pred: DeviceArray, target: DeviceArray):
    if pred.ndim < 2:
        raise ValueError("loss inputs must be (batch, ...) arrays")
    if target.shape != pred.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {target.shape} {pred.shape}")
    bsz = pred.shape[0]
    return bsz, pred.size // bsz


def loss_fwd(kind: int, pred: DeviceArray, target: DeviceArray, holder=None):
    """Per-sample mean over the non-batch dimensions of the MSE / MAE / BCE terms."""
    bsz, d = _loss_dims(pred, target)
    losses = _buf(holder, "_sample_losses", (bsz,))
    lib.nm_loss_fwd_f32(kind, pred.p, target.p, losses.p, bsz, d, stream())
    return losses


def loss_bwd(kind: int, pred: DeviceArray, target: DeviceArray, holder=None):
    """Gradient of the per-sample mean with respect to the predictions (divided by the outputs per sample only)."""
    bsz, d = _loss_dims(pred, target)
    dx = _buf(holder, "_dx", pred.shape)
    lib.nm_loss_bwd_f32(kind, pred.p, target.p, dx.p, bsz, d, stream())
    return dx


# ------------------------------------------------------------- optimizers ----
def adam_step(p: DeviceArray, g: DeviceArray, m: DeviceArray, v: DeviceArray, *, lr, beta_1, beta_2, eps, t,
              n_apply=1, count=None):
    n = p.size if count is None else count
    bc1 = 1.0 - beta_1 ** (t + 1)
    bc2 = 1.0 - beta_2 ** (t + 1)
    tok = _tic("adam_step_f32")
    lib.nm_adam_step_f32(p.p, g.p, m.p, v.p, n, float(lr), float(beta_1), float(beta_2), float(eps),
                         float(bc1), float(bc2), int(n_apply), stream())
    _toc(tok)


def sgd_step(p: DeviceArray, g: DeviceArray, mom, *, lr, momentum, n_apply=1, nonfinite=None, count=None):
    n = p.size if count is None else count
    lib.nm_sgd_step_f32(p.p, g.p, mom.p if mom is not None else None, n, float(lr), float(momentum),
                        int(n_apply), nonfinite.p if nonfinite is not None else None, stream())
