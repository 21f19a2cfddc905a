"""Step driver over np_ref (TEST INFRASTRUCTURE): Model.train's step body.

Reproduces, for a sequential net described by a list of layer specs, exactly
what ``Model.train`` does per step (src/core/Model.py:154-174):
forward -> loss.calculate -> accuracy -> backward -> optimizer triple, including
the reference's quirks:

* Q1  every trainable layer is registered twice (Model.py:56-59, :82-83), so the
      optimizer is applied twice per step with the same gradient and the same
      ``iterations``                                   -> ``n_apply = 2``
* Q2  Dense gradients are divided by the batch a second time (Dense.py:96-98)
* Q3  Conv2D.backward as shipped is a stub             -> ``conv_backward="stub"``;
      the intended math                                -> ``conv_backward="wired"``
* Q4  Conv2D.forward never adds its bias               -> ``conv_bias=False``

Layer specs (dicts):
  {"kind": "conv", "ic":, "oc":, "k": (kh,kw), "padding": "same"|"valid", "relu": bool}
  {"kind": "pool", "size": (ph,pw), "stride": (sh,sw), "padding": "valid"|"same"}
  {"kind": "flatten"}
  {"kind": "dense", "n_in":, "n_out":, "relu": bool, "l1": 0.0, "l2": 0.0}
  {"kind": "softmax"}                (must be last; fused softmax-CE backward,
                                      Model.py:106-109, :339-356)
  {"kind": "sigmoid"} / {"kind": "linear"}   (last, with ``loss="bce" | "mse" | "mae"``: the generic
                                      loss.backward path, Model.py:362-366)
"""
from __future__ import annotations

import numpy as np

from . import np_ref as R

try:                                     # the reference's own compiled max-pool, if built
    from ._ref_loader import load_ref_maxpool
    _REF_POOL = load_ref_maxpool()
except Exception:                        # pragma: no cover - optional
    _REF_POOL = None


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def mnist_cnn_spec():
    """examples/test_Conv2D_MNIST.py:47-53."""
    return [
        {"kind": "conv", "ic": 1, "oc": 32, "k": 3, "padding": "same", "relu": True},
        {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"},
        {"kind": "conv", "ic": 32, "oc": 64, "k": 3, "padding": "same", "relu": True},
        {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"},
        {"kind": "flatten"},
        {"kind": "dense", "n_in": 64 * 7 * 7, "n_out": 10, "relu": False},
        {"kind": "softmax"},
    ]


def mlp_spec(n_in=784, hidden=128, n_out=10):
    """BASELINE.json configs[0]: clean 784-128-10 MLP."""
    return [
        {"kind": "dense", "n_in": n_in, "n_out": hidden, "relu": True},
        {"kind": "dense", "n_in": hidden, "n_out": n_out, "relu": False},
        {"kind": "softmax"},
    ]


def vgg_cifar_spec():
    """SURVEY.md 8(d) config 4 (defined there, not by the reference)."""
    s = []
    ic = 3
    for oc in (64, 128, 256):
        s.append({"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True})
        s.append({"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True})
        s.append({"kind": "pool", "size": 2, "stride": 2, "padding": "valid"})
        ic = oc
    s += [{"kind": "flatten"}, {"kind": "dense", "n_in": 256 * 4 * 4, "n_out": 10, "relu": False},
          {"kind": "softmax"}]
    return s


def init_params(spec, rng: np.random.RandomState, init="random_normal"):
    """Host-side initial parameters in the reference's shapes/dtypes (float64).

    conv W (OC,IC,kh,kw), b (OC,1) (Conv_wrapepr.py:62-66); dense W (in,out),
    b (1,out) (Dense.py:54-56).  RandomNormal: randn*0.01; he: N(0, sqrt(2/fan_in))
    (src/initializers/Initializer.py:90-128).
    """
    params = []
    for L in spec:
        if L["kind"] == "conv":
            kh, kw = _pair(L["k"])
            shape = (L["oc"], L["ic"], kh, kw)
            fan_in = L["ic"] * kh * kw
            b = np.zeros((L["oc"], 1))
        elif L["kind"] == "dense":
            shape = (L["n_in"], L["n_out"])
            fan_in = L["n_in"]
            b = np.zeros((1, L["n_out"]))
        else:
            params.append(None)
            continue
        if init == "he":
            w = rng.normal(0, np.sqrt(2.0 / fan_in), size=shape)
        else:
            w = rng.randn(*shape) * 0.01
        params.append({"weights": w, "biases": b})
    return params


class RefNet:
    """NumPy reference net + optimizer, stepped like Model.train (Model.py:154-174)."""

    def __init__(self, spec, params, *, optimizer="adam", lr=0.001, decay=0.0,
                 eps=1e-7, beta_1=0.9, beta_2=0.999, momentum=0.0,
                 n_apply=2, conv_backward="wired", conv_bias=False,
                 first_layer_dinputs=True, use_ref_pool=True, loss="cce"):
        self.spec = spec
        self.loss = loss                # "cce" (fused with a final softmax) | "mse" | "mae" | "bce" (generic path, Model.py:362-366)
        self.precision = None           # Accuracy_Regression.precision, set by init_accuracy(y)
        self.params = [None if p is None else {k: np.array(v, dtype=np.float64, copy=True)
                                               for k, v in p.items()} for p in params]
        self.opt = dict(kind=optimizer, lr=lr, decay=decay, eps=eps, beta_1=beta_1,
                        beta_2=beta_2, momentum=momentum)
        self.iterations = 0
        self.current_lr = lr
        self.n_apply = n_apply
        self.conv_backward = conv_backward
        self.conv_bias = conv_bias
        self.first_layer_dinputs = first_layer_dinputs
        self.pool = _REF_POOL if (use_ref_pool and _REF_POOL is not None) else None
        self.state = [None if p is None else {k: (np.zeros_like(v), np.zeros_like(v))
                                              for k, v in p.items()} for p in self.params]
        self.grads = [None] * len(spec)
        self.dinputs = [None] * len(spec)
        self.dropout_masks = {}         # layer index -> mask for the next training forward (Layer_Dropout.binary_mask)

    # -- forward ------------------------------------------------------------
    def forward(self, x, training=True):
        self.x = x
        self.cache = []
        self.outs = []          # every layer's output, for per-layer parity
        a = x
        for L, p in zip(self.spec, self.params):
            k = L["kind"]
            if k == "conv":
                out, pre = R.conv_layer_forward(a, p["weights"], p["biases"], padding=L["padding"],
                                                relu=L.get("relu", False), add_bias=self.conv_bias)
                self.cache.append((a, pre))
                a = out
            elif k == "pool":
                ph, pw = _pair(L["size"])
                sh, sw = _pair(L.get("stride") or L["size"])
                fwd = self.pool.maxpool2d_forward_cpu if self.pool else R.maxpool2d_forward_cpu
                out, idx = fwd(np.asarray(a, dtype=np.float32), ph, pw, sh, sw, L.get("padding", "valid"))
                self.cache.append((a, idx))
                a = out
            elif k == "flatten":
                self.cache.append((a.shape,))
                a = a.reshape(a.shape[0], -1)
            elif k == "dense":
                pre = R.dense_forward(a, p["weights"], p["biases"])
                self.cache.append((a, pre))
                a = R.relu_forward(pre) if L.get("relu", False) else pre
            elif k == "softmax":
                self.cache.append((a,))
                a = R.softmax_forward(a)
            elif k == "sigmoid":
                self.cache.append((a,))
                a = R.sigmoid_forward(a)
            elif k == "linear":
                self.cache.append((a,))                 # Activation_Linear.forward: output = inputs (Linear.py:5-7)
            elif k == "dropout":
                # Dropout.py:21-30; the mask (Bernoulli(keep)/keep) is supplied by the caller: self.dropout_masks[layer index]
                mask = self.dropout_masks[len(self.cache)] if training else None
                self.cache.append((mask,))
                a = R.dropout_forward(a, mask) if training else a.copy()
            else:
                raise ValueError(k)
            self.outs.append(a)
        self.output = a
        return a

    # -- backward -----------------------------------------------------------
    def backward(self, output, y, *, global_batch=None):
        spec = self.spec
        last = spec[-1]["kind"]
        if self.loss == "cce":
            assert last == "softmax", "fused softmax-CE path only"
            d = R.softmax_cce_backward(output, y, batch=global_batch)  # Model.py:339-348
        else:
            # generic path (Model.py:362-366): loss.backward, then every layer in reverse -- the output activation first
            assert global_batch is None and last in ("sigmoid", "linear")
            d = {"mse": R.mse_backward, "mae": R.mae_backward, "bce": R.bce_backward}[self.loss](output, y)
            d = R.sigmoid_backward(output, d) if last == "sigmoid" else d.copy()
        self.dinputs[len(spec) - 1] = d
        for i in range(len(spec) - 2, -1, -1):                         # Model.py:353-354
            L, p, c = spec[i], self.params[i], self.cache[i]
            k = L["kind"]
            if k == "dense":
                a, pre = c
                if L.get("relu", False):
                    d = R.relu_backward(d, pre)
                if global_batch is not None:
                    # data-parallel rule: divide by the GLOBAL batch
                    dx = np.dot(d, p["weights"].T)
                    dw = np.dot(a.T, d) / global_batch
                    db = np.sum(d, axis=0, keepdims=True) / global_batch
                    if L.get("l1", 0) > 0:
                        dw = dw + L["l1"] * np.sign(p["weights"])
                    if L.get("l2", 0) > 0:
                        dw = dw + 2 * L["l2"] * p["weights"]
                else:
                    dx, dw, db = R.dense_backward(a, p["weights"], d, l1=L.get("l1", 0.0), l2=L.get("l2", 0.0))
                self.grads[i] = {"weights": dw, "biases": db}
                d = dx
            elif k == "flatten":
                d = d.reshape(c[0])
            elif k == "dropout":
                d = R.dropout_backward(d, c[0])
            elif k == "pool":
                a, idx = c
                ph, pw = _pair(L["size"])
                sh, sw = _pair(L.get("stride") or L["size"])
                bwd = self.pool.maxpool2d_backward_cpu if self.pool else R.maxpool2d_backward_cpu
                d = bwd(np.asarray(d, dtype=np.float32), idx, np.asarray(a, dtype=np.float32),
                        ph, pw, sh, sw, L.get("padding", "valid"))
            elif k == "conv":
                a, pre = c
                dx, dw, db = R.conv_layer_backward(a, p["weights"], pre, d, padding=L["padding"],
                                                   relu=L.get("relu", False), mode=self.conv_backward)
                self.grads[i] = {"weights": dw, "biases": db}
                d = dx
            else:
                raise ValueError(k)
            self.dinputs[i] = d

    # -- optimizer ----------------------------------------------------------
    def update(self):
        o = self.opt
        self.current_lr = R.decayed_lr(o["lr"], o["decay"], self.iterations)   # pre_update_params
        for i, p in enumerate(self.params):
            if p is None:
                continue
            for _ in range(self.n_apply):                                     # Q1
                for name in ("weights", "biases"):
                    g = self.grads[i][name]
                    m, v = self.state[i][name]
                    if o["kind"] == "adam":
                        p[name], m, v = R.adam_update(p[name], g, m, v, lr=self.current_lr,
                                                      beta_1=o["beta_1"], beta_2=o["beta_2"],
                                                      eps=o["eps"], t=self.iterations)
                    elif o["kind"] == "sgd":
                        p[name], m = R.sgd_update(p[name], g, m, lr=self.current_lr,
                                                  momentum=o["momentum"])
                    # the three below keep per-tensor state here; the reference keys it by the bare parameter name, which
                    # is the same thing for the single-trainable-layer models it can run them on
                    elif o["kind"] == "rmsprop":
                        p[name], m = R.rmsprop_update(p[name], g, m, lr=self.current_lr, rho=o.get("rho", 0.9), eps=o["eps"])
                    elif o["kind"] == "adagrad":
                        p[name], m = R.adagrad_update(p[name], g, m, lr=self.current_lr, eps=o["eps"])
                    elif o["kind"] == "nadam":
                        p[name], m, v = R.nadam_update(p[name], g, m, v, lr=self.current_lr, beta1=o["beta_1"],
                                                       beta2=o["beta_2"], eps=o["eps"], t=self.iterations)
                    else:
                        raise ValueError(o["kind"])
                    self.state[i][name] = (m, v)
        self.iterations += 1                                                   # post_update_params

    # -- one Model.train step -------------------------------------------------
    def step(self, x, y, *, global_batch=None, update=True):
        out = self.forward(x, training=True)
        if self.loss == "cce":
            losses = R.cce_forward(out, y)
            acc = R.accuracy_categorical(out, y)
        else:
            losses = {"mse": R.mse_forward, "mae": R.mae_forward, "bce": R.bce_forward}[self.loss](out, y)
            if self.spec[-1]["kind"] == "sigmoid":          # Accuracy_Categorical(binary=True) on Sigmoid.predictions
                hits = R.sigmoid_predictions(out) == y
            else:                                           # Accuracy_Regression on Linear.predictions
                hits = np.absolute(out - y) < self.precision
            acc = R.accuracy_elementwise(hits)[0]
        loss = float(np.mean(losses))
        self.backward(out, y, global_batch=global_batch)
        if update:
            self.update()
        return {"loss": loss, "loss_sum": float(np.sum(losses)), "acc": acc, "output": out}

    def init_accuracy(self, y):
        """Accuracy_Regression.init (Accuracy_Regression.py:10-12); Model.train calls it before the first epoch (:119)."""
        if self.precision is None:
            self.precision = R.accuracy_regression_precision(y)

    # -- flat views (arena order: per trainable layer, weights then biases) --
    def flat(self, what="params"):
        chunks = []
        for i, p in enumerate(self.params):
            if p is None:
                continue
            src = p if what == "params" else self.grads[i]
            chunks += [np.asarray(src["weights"], dtype=np.float64).ravel(),
                       np.asarray(src["biases"], dtype=np.float64).ravel()]
        return np.concatenate(chunks)

    def set_flat(self, vec):
        off = 0
        for p in self.params:
            if p is None:
                continue
            for name in ("weights", "biases"):
                n = p[name].size
                p[name] = np.asarray(vec[off:off + n], dtype=np.float64).reshape(p[name].shape).copy()
                off += n
