"""Functional NumPy restatement of the Neuromah hot path (TEST INFRASTRUCTURE).

Every function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates.  dtypes follow the reference exactly: conv and pool run
in float32 (pybind ``py::array_t<float>`` force-cast, CPU_kernels/Conv2D.cpp:83,
maxpooling_backend.cpp:24), everything else in whatever NumPy promotes to
(float64 once a float64 weight is involved).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.
"""
from __future__ import annotations

import math

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

# ----------------------------------------------------------------------------
# Conv2D  (intended math of src/layers/CPU_kernels/Conv2D.cpp)
# ----------------------------------------------------------------------------


def same_padding(kh: int, kw: int):
    """'same' zero padding amounts, src/layers/Conv_wrapepr.py:130-139.

    odd k -> ((k-1)//2, (k-1)//2); even k -> (k//2-1, k//2).
    Returns ((top, bottom), (left, right)).
    """
    def one(k):
        return ((k - 1) // 2, (k - 1) // 2) if k % 2 else (k // 2 - 1, k // 2)
    return one(kh), one(kw)


def im2col(x: np.ndarray, kh: int, kw: int) -> np.ndarray:
    """(N,C,H,W) -> (N, C*kh*kw, oh*ow); row=(c,p,q), col=(i,j).

    Index order of CPU_kernels/Conv2D.cpp:19-43 (stride 1, no padding).
    """
    n, c, h, w = x.shape
    oh, ow = h - kh + 1, w - kw + 1
    win = sliding_window_view(x, (kh, kw), axis=(2, 3))      # (N,C,oh,ow,kh,kw)
    cols = win.transpose(0, 1, 4, 5, 2, 3).reshape(n, c * kh * kw, oh * ow)
    return np.ascontiguousarray(cols)


def conv2d_cpu(x, w, *, dtype=np.float32, chunk: int = 256) -> np.ndarray:
    """Valid, stride-1 cross-correlation: out[b] = W.reshape(OC,K) @ im2col(x[b]).

    Intended math of ``conv2d_cpu`` (CPU_kernels/Conv2D.cpp:83-146, comments at
    :113 and :133-134).  Padding is applied by the caller (Conv_wrapepr.py:78-84).
    ``dtype`` defaults to float32 like the pybind boundary; float64 is accepted
    for finite-difference checks.
    """
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    if x.ndim != 4 or w.ndim != 4:
        raise RuntimeError("Input and kernel must be 4D tensors")          # Conv2D.cpp:86-87
    n, c, h, wd = x.shape
    oc, ic, kh, kw = w.shape
    if ic != c:
        raise RuntimeError("Mismatch between input channels and kernel channels")  # :101-102
    oh, ow = h - kh + 1, wd - kw + 1
    wm = w.reshape(oc, c * kh * kw)
    out = np.empty((n, oc, oh, ow), dtype=dtype)
    for s in range(0, n, chunk):
        cols = im2col(x[s:s + chunk], kh, kw)                 # (b,K,P)
        out[s:s + chunk] = np.matmul(wm, cols).reshape(-1, oc, oh, ow)
    return out


def conv2d_backward_cpu(x, w, dout, *, dtype=np.float32, chunk: int = 256):
    """(dinputs, dweights) of the valid stride-1 conv.

    dW = sum_b dOut[b].reshape(OC,P) @ cols_b^T (no batch normalisation),
    dX[b] = col2im(W.reshape(OC,K)^T @ dOut[b]).
    Intended math of ``conv2d_backward_cpu`` (CPU_kernels/Conv2D.cpp:150-240,
    col2im :48-74).
    """
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    dout = np.asarray(dout, dtype=dtype)
    n, c, h, wd = x.shape
    oc, _, kh, kw = w.shape
    oh, ow = dout.shape[2], dout.shape[3]
    k = c * kh * kw
    wm = w.reshape(oc, k)
    dwm = np.zeros((oc, k), dtype=dtype)
    dx = np.zeros_like(x)
    for s in range(0, n, chunk):
        cols = im2col(x[s:s + chunk], kh, kw)                          # (b,K,P)
        dom = dout[s:s + chunk].reshape(-1, oc, oh * ow)               # (b,OC,P)
        dwm += np.einsum("bop,bkp->ok", dom, cols, optimize=True)
        dxc = np.matmul(wm.T, dom).reshape(-1, c, kh, kw, oh, ow)       # (b,C,kh,kw,oh,ow)
        dxs = dx[s:s + chunk]
        for p in range(kh):
            for q in range(kw):
                dxs[:, :, p:p + oh, q:q + ow] += dxc[:, :, p, q]
    return dx, dwm.reshape(w.shape)


def conv_layer_forward(x, w, b, *, padding="valid", relu=False, add_bias=False,
                       dtype=np.float32):
    """Layer_Conv2D.forward (src/layers/Conv_wrapepr.py:73-98).

    'same' pads with zeros (:78-84); bias is NOT added by the reference (Q4)
    unless ``add_bias``; the embedded activation runs on the conv output (:96-98).
    Returns (output, pre_activation).
    """
    kh, kw = w.shape[2], w.shape[3]
    x = np.asarray(x)
    if padding == "same":
        (pt, pb), (pl, pr) = same_padding(kh, kw)
        xp = np.pad(x, ((0, 0), (0, 0), (pt, pb), (pl, pr)), mode="constant")
    else:
        xp = x
    pre = conv2d_cpu(xp, w, dtype=dtype)
    if add_bias:
        pre = pre + np.asarray(b, dtype=dtype).reshape(1, -1, 1, 1)
    out = np.maximum(0, pre) if relu else pre
    return out, pre


def conv_layer_backward(x, w, pre, dvalues, *, padding="valid", relu=False,
                        mode="wired", dtype=np.float32):
    """Layer_Conv2D.backward.

    mode="stub"  : the reference as shipped (Conv_wrapepr.py:100-115): ReLU
                   backward, dbiases = sum_{n,h,w} dvalues -> (OC,1), and
                   zeros for dweights / dinputs (Q3).
    mode="wired" : the intended math -- same ReLU/bias handling, then
                   conv2d_backward_cpu on the padded input with the padding
                   cropped off dinputs.
    Returns (dinputs, dweights, dbiases).
    """
    dvalues = np.asarray(dvalues)
    if relu:
        dvalues = dvalues.copy()
        dvalues[pre <= 0] = 0                                   # Relu.py:12-14
    oc = dvalues.shape[1]
    dbias = dvalues.transpose(0, 2, 3, 1).reshape(-1, oc).sum(axis=0, keepdims=True).T
    if mode == "stub":
        return np.zeros_like(x), np.zeros_like(w), dbias
    kh, kw = w.shape[2], w.shape[3]
    if padding == "same":
        (pt, pb), (pl, pr) = same_padding(kh, kw)
        xp = np.pad(np.asarray(x), ((0, 0), (0, 0), (pt, pb), (pl, pr)), mode="constant")
    else:
        pt = pb = pl = pr = 0
        xp = np.asarray(x)
    dxp, dw = conv2d_backward_cpu(xp, w, dvalues, dtype=dtype)
    h, wd = x.shape[2], x.shape[3]
    dx = dxp[:, :, pt:pt + h, pl:pl + wd]
    return np.ascontiguousarray(dx), dw, dbias


# ----------------------------------------------------------------------------
# MaxPooling2D  (src/layers/CPU_kernels/maxpooling_backend.cpp)
# ----------------------------------------------------------------------------


def _pool_geometry(in_h, in_w, ph, pw, sh, sw, padding):
    """Output size and padding, maxpooling_backend.cpp:37-52."""
    if padding == "same":
        oh = int(math.ceil(in_h / sh))
        ow = int(math.ceil(in_w / sw))
        pad_h = max((oh - 1) * sh + ph - in_h, 0)
        pad_w = max((ow - 1) * sw + pw - in_w, 0)
        pt = pad_h // 2
        pl = pad_w // 2
        return oh, ow, pt, pad_h - pt, pl, pad_w - pl
    return (in_h - ph) // sh + 1, (in_w - pw) // sw + 1, 0, 0, 0, 0


def maxpool2d_forward_cpu(x, ph, pw, sh, sw, padding="valid"):
    """(out f32[N,C,oh,ow], idx i32[N,C,oh,ow,2]).

    Window max with the (row, col) of the max in PADDED coordinates;
    'same' pads with 0.0 which participates in the max (:59, Q8); ties keep the
    first element of the row-major window scan (strict '>' at :104).
    maxpooling_backend.cpp:24-133.
    """
    x = np.asarray(x, dtype=np.float32)
    if x.ndim != 4:
        raise RuntimeError("Input must be a 4D tensor")
    n, c, h, w = x.shape
    oh, ow, pt, pb, pl, pr = _pool_geometry(h, w, ph, pw, sh, sw, padding)
    xp = np.pad(x, ((0, 0), (0, 0), (pt, pb), (pl, pr))) if padding == "same" else x
    win = sliding_window_view(xp, (ph, pw), axis=(2, 3))[:, :, ::sh, ::sw][:, :, :oh, :ow]
    flat = win.reshape(n, c, oh, ow, ph * pw)
    arg = np.argmax(flat, axis=-1)                    # first max in row-major scan
    out = np.take_along_axis(flat, arg[..., None], axis=-1)[..., 0]
    hh = (np.arange(oh) * sh)[None, None, :, None]
    ww = (np.arange(ow) * sw)[None, None, None, :]
    idx = np.stack([hh + arg // pw, ww + arg % pw], axis=-1).astype(np.int32)
    return np.ascontiguousarray(out, dtype=np.float32), idx


def maxpool2d_backward_cpu(dvalues, idx, x, ph, pw, sh, sw, padding="valid"):
    """Scatter-add dvalues to the saved max positions, then strip the padding.

    maxpooling_backend.cpp:145-238.
    """
    dvalues = np.asarray(dvalues, dtype=np.float32)
    n, c, h, w = np.asarray(x).shape
    _, _, pt, pb, pl, pr = _pool_geometry(h, w, ph, pw, sh, sw, padding)
    hp, wp = h + pt + pb, w + pl + pr
    dpad = np.zeros((n, c, hp, wp), dtype=np.float32)
    nn, cc = np.meshgrid(np.arange(n), np.arange(c), indexing="ij")
    nn = np.broadcast_to(nn[:, :, None, None], dvalues.shape)
    cc = np.broadcast_to(cc[:, :, None, None], dvalues.shape)
    np.add.at(dpad, (nn, cc, idx[..., 0], idx[..., 1]), dvalues)
    return np.ascontiguousarray(dpad[:, :, pt:pt + h, pl:pl + w])


# ----------------------------------------------------------------------------
# Dense / activations / loss
# ----------------------------------------------------------------------------


def dense_forward(x, w, b):
    """X @ W + b, src/layers/Dense.py:79-80 (NumPy promotion: f32 @ f64 -> f64)."""
    x = np.asarray(x)
    if x.shape[1] != w.shape[0]:
        raise ValueError(f"Input features mismatch. Expected {w.shape[0]}, got {x.shape[1]}")
    if np.any(np.isnan(x)) or np.any(np.isinf(x)):
        raise ValueError("Input contains NaN/Inf values")                # Dense.py:77-78
    return np.dot(x, w) + b


def dense_backward(x, w, dvalues, *, l1=0.0, l2=0.0):
    """(dinputs, dweights, dbiases), src/layers/Dense.py:95-108.

    dweights = X^T dY / B and dbiases = sum dY / B (second batch normalisation,
    Q2); L1 adds l1*sign(W), L2 adds 2*l2*W; dinputs = dY W^T (not re-divided).
    """
    bsz = dvalues.shape[0]
    dw = np.dot(x.T, dvalues) / bsz
    db = np.sum(dvalues, axis=0, keepdims=True) / bsz
    if l1 > 0:
        dw = dw + l1 * np.sign(w)
    if l2 > 0:
        dw = dw + 2 * l2 * w
    return np.dot(dvalues, w.T), dw, db


def relu_forward(x):
    """src/activations/Relu.py:7-9."""
    return np.maximum(0, x)


def relu_backward(dvalues, pre):
    """Zero where the PRE-activation input is <= 0, src/activations/Relu.py:12-14."""
    d = np.array(dvalues, copy=True)
    d[pre <= 0] = 0
    return d


def softmax_forward(x):
    """Row softmax with max subtraction, src/activations/Softmax.py:7-11."""
    e = np.exp(x - np.max(x, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def softmax_backward(out, dvalues):
    """Per-sample Jacobian product, src/activations/Softmax.py:13-18 (Q10 only).

    J = diag(s) - s s^T  =>  J d = s * (d - <s, d>).
    """
    dot = np.sum(out * dvalues, axis=1, keepdims=True)
    return out * (dvalues - dot)


def cce_forward(y_pred, y_true):
    """Per-sample -log(clip(p)[y]), src/losses/categorical_crossentropy.py:8-25."""
    n = len(y_pred)
    p = np.clip(y_pred, 1e-7, 1 - 1e-7)
    y_true = np.asarray(y_true)
    if y_true.ndim == 1:
        conf = p[np.arange(n), y_true]
    elif y_true.ndim == 2:
        conf = np.sum(p * y_true, axis=1)
    else:
        raise ValueError("y_true must be either sparse labels or one-hot encoded")
    return -np.log(conf)


def cce_backward(dvalues, y_true):
    """-y/p / B, src/losses/categorical_crossentropy.py:27-38."""
    n = len(dvalues)
    y_true = np.asarray(y_true)
    if y_true.ndim == 1:
        y_true = np.eye(dvalues.shape[1])[y_true]
    return (-y_true / dvalues) / n


def softmax_cce_backward(probs, y_true, *, batch=None):
    """(p - onehot(y)) / B, src/losses/Activation_Softmax_Loss_CategoricalCrossentropy.py:4-19.

    ``batch`` overrides the divisor (global batch under data parallelism,
    SURVEY 8e); default is len(probs) like the reference.
    """
    n = len(probs)
    y_true = np.asarray(y_true)
    if y_true.ndim == 2:
        y_true = np.argmax(y_true, axis=1)
    d = np.array(probs, copy=True)
    d[np.arange(n), y_true] -= 1
    return d / (n if batch is None else batch)


def accuracy_categorical(probs, y_true):
    """mean(argmax(p) == y), src/metrics/Accuracy_Categorical.py:13-16, Softmax.py:20-22."""
    y_true = np.asarray(y_true)
    if y_true.ndim == 2:
        y_true = np.argmax(y_true, axis=1)
    return float(np.mean(np.argmax(probs, axis=1) == y_true))


def sigmoid_forward(x):
    """src/activations/Sigmoid.py:5-7."""
    return 1 / (1 + np.exp(-x))


def sigmoid_backward(out, dvalues):
    """dvalues * (1 - out) * out on the forward OUTPUT, src/activations/Sigmoid.py:9-10."""
    return dvalues * (1 - out) * out


def sigmoid_predictions(out):
    """Threshold at 0.5, src/activations/Sigmoid.py:12-14."""
    return (out > 0.5) * 1


def _per_sample_mean(terms):
    return np.mean(terms, axis=tuple(range(1, terms.ndim)))


def mse_forward(y_pred, y_true):
    """Per-sample mean of (y - p)^2 over the non-batch axes, src/losses/MSE.py:10-14."""
    return _per_sample_mean((y_true - y_pred) ** 2)


def mse_backward(dvalues, y_true):
    """-2 (y - p) / (prod(shape) / samples): NOT divided by the batch, src/losses/MSE.py:17-28."""
    return -2 * (y_true - dvalues) / (np.prod(dvalues.shape) / len(dvalues))


def mae_forward(y_pred, y_true):
    """Per-sample mean of |y - p|, src/losses/MAE.py:11-14."""
    return _per_sample_mean(np.abs(y_true - y_pred))


def mae_backward(dvalues, y_true):
    """sign(y - p) / (prod(shape) / samples) -- the reference's sign, as shipped, src/losses/MAE.py:17-28."""
    return np.sign(y_true - dvalues) / (np.prod(dvalues.shape) / len(dvalues))


def bce_forward(y_pred, y_true):
    """Per-sample mean of -(y log p + (1 - y) log(1 - p)), p clipped to [1e-7, 1 - 1e-7],
    src/losses/binary_crossentropy.py:9-18."""
    p = np.clip(y_pred, 1e-7, 1 - 1e-7)
    return _per_sample_mean(-(y_true * np.log(p) + (1 - y_true) * np.log(1 - p)))


def bce_backward(dvalues, y_true):
    """-(y / p - (1 - y) / (1 - p)) / (prod(shape) / samples) on the clipped predictions,
    src/losses/binary_crossentropy.py:20-33."""
    p = np.clip(dvalues, 1e-7, 1 - 1e-7)
    return -(y_true / p - (1 - y_true) / (1 - p)) / (np.prod(dvalues.shape) / len(dvalues))


def accuracy_regression_precision(y):
    """std(y) / 250, src/metrics/Accuracy_Regression.py:10-12."""
    return np.std(y) / 250


def accuracy_elementwise(hits):
    """Accuracy.calculate (src/core/Accuracy.py:19-23): the batch value is the mean over EVERY compared element,
    the running sum counts elements but the running count counts samples."""
    hits = np.asarray(hits)
    return float(np.mean(hits)), float(np.sum(hits)), len(hits)


# ----------------------------------------------------------------------------
# Optimizers
# ----------------------------------------------------------------------------


def adam_update(param, grad, m, v, *, lr, beta_1=0.9, beta_2=0.999, eps=1e-7, t=0):
    """One Optimizer_Adam.update_params application (src/optimizers/Adam.py:86-100).

    m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  bias-correct with t+1;
    param -= lr * m_hat / (sqrt(v_hat) + eps)   (eps OUTSIDE the sqrt).
    ``t`` is ``optimizer.iterations`` and is NOT advanced between the duplicate
    applications of Q1.  Returns new (param, m, v).
    """
    m = beta_1 * m + (1.0 - beta_1) * grad
    v = beta_2 * v + (1.0 - beta_2) * grad ** 2
    m_hat = m / (1.0 - beta_1 ** (t + 1))
    v_hat = v / (1.0 - beta_2 ** (t + 1))
    param = param - lr * m_hat / (np.sqrt(v_hat) + eps)
    return param, m, v


def sgd_update(param, grad, mom, *, lr, momentum=0.0):
    """Optimizer_SGD.update_params (src/optimizers/SGD.py:85-97).

    momentum: v = mu v + lr g; param -= v.   plain: param -= lr g.
    NaN/Inf gradients raise ValueError (:80-83).
    """
    if np.any(np.isnan(grad)):
        raise ValueError("Gradient contains NaN values.")
    if np.any(np.isinf(grad)):
        raise ValueError("Gradient contains Inf values.")
    if momentum:
        mom = mom * momentum + lr * grad
        return param - mom, mom
    return param - lr * grad, mom


def rmsprop_update(param, grad, cache, *, lr, rho=0.9, eps=1e-7):
    """Optimizer_RMSprop.update_params (src/optimizers/RMSprop.py:103-109): cache = rho cache + (1-rho) g^2;
    param -= lr g / (sqrt(cache) + eps).  NaN/Inf gradients raise ValueError (:93-96)."""
    if np.any(np.isnan(grad)) or np.any(np.isinf(grad)):
        raise ValueError("Gradient contains NaN/Inf values.")
    cache = rho * cache + (1.0 - rho) * grad ** 2
    return param - lr * grad / (np.sqrt(cache) + eps), cache


def adagrad_update(param, grad, cache, *, lr, eps=1e-7):
    """Optimizer_Adagrad.update_params (src/optimizers/Adagrad.py, last two statements of the loop): cache += g^2;
    param -= lr g / (sqrt(cache) + eps).  NaN/Inf gradients raise ValueError."""
    if np.any(np.isnan(grad)) or np.any(np.isinf(grad)):
        raise ValueError("Gradient contains NaN/Inf values.")
    cache = cache + grad ** 2
    return param - lr * grad / (np.sqrt(cache) + eps), cache


def nadam_update(param, grad, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, t=0):
    """Optimizer_Nadam.update_params (src/optimizers/Nadam.py): eps is added to the bias-correction denominators too.
    ``t`` = optimizer.iterations, not advanced between the duplicate applications of Q1."""
    m = beta1 * m + (1.0 - beta1) * grad
    v = beta2 * v + (1.0 - beta2) * grad ** 2
    b1p = 1.0 - beta1 ** (t + 1)
    b2p = 1.0 - beta2 ** (t + 1)
    m_hat = m / (b1p + eps)
    v_hat = v / (b2p + eps)
    m_nesterov = beta1 * m_hat + (1.0 - beta1) * grad / (b1p + eps)
    return param - lr * m_nesterov / (np.sqrt(v_hat) + eps), m, v


def dropout_forward(x, mask):
    """Layer_Dropout.forward in training mode with a given mask = Bernoulli(keep)/keep (src/layers/Dropout.py:27-30)."""
    return x * mask


def dropout_backward(dvalues, mask):
    """Layer_Dropout.backward (src/layers/Dropout.py:40)."""
    return dvalues * mask


def decayed_lr(lr, decay, iterations):
    """lr / (1 + decay * t), Adam.py:45-47 / SGD.py:52-53."""
    return lr * (1.0 / (1.0 + decay * iterations)) if decay else lr


# ----------------------------------------------------------------------------
# Federated
# ----------------------------------------------------------------------------


def federated_average(updates, num_samples):
    """sum_i u_i * (n_i / N), Federated/utils.py:20-23 (returns ndarray, not list)."""
    total = sum(num_samples)
    return np.sum([np.asarray(u) * (s / total) for u, s in zip(updates, num_samples)], axis=0)
