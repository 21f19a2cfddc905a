"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Only usable in the authoring container (the GPU box has no /root/reference).
The reference does not import as shipped: src/layers/__init__.py:9
imports a missing ``.RNN`` and Conv_wrapepr.py:3 / MaxPooling2D.py:2 import a
missing ``.compiled`` package.  We pre-register those two sub-modules in
``sys.modules`` and expose the tree under the package name ``NeuromahRef`` via a
symlink in a temp dir (all intra-package imports of the reference are relative,
so the top-level name is free).  Nothing of the reference is copied.

``compiled._MaxPooling2DBackend_cpp``  = the reference's own C++ built by
                                          oracle/build_ref.py
``compiled._Conv2DBackend_cpp``        = oracle.np_ref restatement (the real
                                          Conv2D.cpp needs Eigen)
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np

from . import np_ref as R
from .build_ref import build as _build_ref
from ._ref_loader import load_ref_maxpool

REF_ROOT = "/root/reference"
PKG = "NeuromahRef"
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


def load():
    """Return the imported reference package (``NeuromahRef``)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present (expected only in the authoring container)")
    _build_ref()
    pool = load_ref_maxpool()
    tmp = tempfile.mkdtemp(prefix="nmref_")
    os.symlink(REF_ROOT, os.path.join(tmp, PKG))
    sys.path.insert(0, tmp)

    conv_backend = types.SimpleNamespace(
        conv2d_cpu=lambda x, w: R.conv2d_cpu(x, w),
        conv2d_backward_cpu=lambda x, w, d: R.conv2d_backward_cpu(x, w, d))
    compiled = types.ModuleType(f"{PKG}.src.layers.compiled")
    compiled._Conv2DBackend_cpp = conv_backend
    compiled._MaxPooling2DBackend_cpp = pool
    rnn = types.ModuleType(f"{PKG}.src.layers.RNN")
    rnn.Layer_RNN = type("Layer_RNN", (), {})
    sys.modules[compiled.__name__] = compiled
    sys.modules[rnn.__name__] = rnn

    import importlib
    pkg = importlib.import_module(PKG)
    for sub in ("layers", "activations", "losses", "optimizers", "core", "metrics", "initializers"):
        importlib.import_module(f"{PKG}.src.{sub}")
    _loaded = pkg
    return pkg


def wired_conv_class():
    """Layer_Conv2D subclass whose backward actually calls conv2d_backward_cpu.

    The shipped backward is a stub (Conv_wrapepr.py:100-115); this is
    the "wired" oracle of SURVEY 8c: same activation / bias-gradient lines, then
    the backend's dgrad/wgrad on the padded input with the padding cropped.
    """
    load()
    from NeuromahRef.src.layers import Layer_Conv2D
    from NeuromahRef.src.layers.compiled import _Conv2DBackend_cpp as be

    class Layer_Conv2D_Wired(Layer_Conv2D):
        def backward(self, dvalues):
            super().backward(dvalues)              # activation bwd + dbiases (+ zeros)
            if self.activation:
                dvalues = self.activation.dinputs
            x = self.inputs
            _, _, h, w = x.shape
            _, _, pad_h, pad_w = self._calculate_output_shape(h, w)
            xp = np.pad(x, ((0, 0), (0, 0), pad_h, pad_w)) if self.padding == "same" else x
            dxp, dw = be.conv2d_backward_cpu(xp, self.weights, dvalues)
            self.weight_gradients = dw
            self.dinputs = np.ascontiguousarray(dxp[:, :, pad_h[0]:pad_h[0] + h, pad_w[0]:pad_w[0] + w])

    return Layer_Conv2D_Wired


def build_model(spec, params, *, optimizer="adam", lr=0.001, momentum=0.0, decay=0.0,
                conv_backward="wired"):
    """Build a reference ``Model`` from an oracle.ref_net spec with given initial params."""
    load()
    from NeuromahRef.src.core import Model
    from NeuromahRef.src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten, Layer_Dropout
    from NeuromahRef.src.activations import Activation_ReLU, Activation_Softmax
    from NeuromahRef.src.losses import Loss_CategoricalCrossentropy
    from NeuromahRef.src.optimizers import (Optimizer_Adam, Optimizer_SGD, Optimizer_RMSprop, Optimizer_Adagrad,
                                            Optimizer_Nadam)
    from NeuromahRef.src.metrics import Accuracy_Categorical

    Conv = wired_conv_class() if conv_backward == "wired" else Layer_Conv2D
    model = Model()
    for L, p in zip(spec, params):
        k = L["kind"]
        if k == "conv":
            layer = Conv(L["ic"], L["oc"], L["k"], padding=L["padding"],
                         activation=Activation_ReLU() if L.get("relu") else None)
        elif k == "pool":
            layer = Layer_MaxPooling2D(L["size"], strides=L.get("stride"), padding=L.get("padding", "valid"))
        elif k == "flatten":
            layer = Layer_Flatten()
        elif k == "dense":
            layer = Layer_Dense(L["n_in"], L["n_out"], activation=Activation_ReLU() if L.get("relu") else None,
                                weight_regularizer_l1=L.get("l1", 0), weight_regularizer_l2=L.get("l2", 0))
        elif k == "softmax":
            layer = Activation_Softmax()
        elif k == "dropout":
            layer = Layer_Dropout(L["rate"])
        else:
            raise ValueError(k)
        if p is not None:
            layer.weights = np.array(p["weights"], dtype=np.float64, copy=True)
            layer.biases = np.array(p["biases"], dtype=np.float64, copy=True)
        model.add(layer)
    opt = {"adam": lambda: Optimizer_Adam(learning_rate=lr, decay=decay),
           "sgd": lambda: Optimizer_SGD(learning_rate=lr, decay=decay, momentum=momentum),
           "rmsprop": lambda: Optimizer_RMSprop(learning_rate=lr, decay=decay),
           "adagrad": lambda: Optimizer_Adagrad(learning_rate=lr, decay=decay),
           "nadam": lambda: Optimizer_Nadam(learning_rate=lr, decay=decay)}[optimizer]()
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=opt, accuracy=Accuracy_Categorical)
    model.finalize()
    model.loss.new_pass()          # Model.train does this at the start of every epoch (:133-134)
    model.accuracy.new_pass()
    return model


def train_step(model, x, y):
    """One step exactly as Model.train's loop body (src/core/Model.py:154-174)."""
    output = model.forward(x, training=True)
    losses = model.loss.calculate(output, y, include_regularization=True)
    predictions = model.output_layer_activation.predictions(output)
    acc = model.accuracy.calculate(predictions, y)
    model.backward(output, y)
    model.optimizer.pre_update_params()
    for layer in model.trainable_layers:
        model.optimizer.update_params(layer)
    model.optimizer.post_update_params()
    return {"loss": float(losses["data_loss"]), "acc": float(acc), "output": output}
