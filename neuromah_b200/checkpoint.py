"""Whole-model checkpoints: Model.save / Model.load (reference: src/core/Model.py:396-417).

The reference pickles a deep copy of the live Model with the activations stripped (Model.py:398-411).  A model here
owns device memory, streams, CUDA graphs and (data parallel) IPC mappings, none of which survive a pickle, so the file
holds a HOST-SIDE description instead -- layer classes + constructor arguments + parameters, the optimizer's class,
hyper-parameters, iteration count and state slabs, the loss / accuracy classes and the kernel mode -- and ``Model.load``
rebuilds and finalizes an equivalent model from it.  ``save_parameters`` / ``load_parameters`` keep the reference's
own pickle layout (list of ``{"weights": (W, dW), "biases": (b, db)}``, Model.py:369-394).
"""
import pickle

import numpy as np

FORMAT = "neuromah_b200.model.v1"


def _activation_name(act):
    return None if act is None else type(act).__name__


def _activation(name):
    if name is None:
        return None
    from .src import activations
    return getattr(activations, name)()


def layer_state(layer):
    from .src.layers import Layer_Conv2D, Layer_Dense, Layer_Dropout, Layer_Flatten, Layer_MaxPooling2D
    from .src.activations import Activation_Linear, Activation_ReLU, Activation_Sigmoid, Activation_Softmax
    if isinstance(layer, Layer_Dense):
        n_in, n_out = layer.weights.shape
        st = {"cls": "Layer_Dense", "args": dict(n_inputs=int(n_in), n_neurons=int(n_out),
                                                  weight_regularizer_l1=layer.weight_regularizer_l1,
                                                  weight_regularizer_l2=layer.weight_regularizer_l2),
              "activation": _activation_name(layer.activation)}
    elif isinstance(layer, Layer_Conv2D):
        oc, ic = layer.weights.shape[:2]
        st = {"cls": "Layer_Conv2D", "args": dict(in_channels=int(ic), out_channels=int(oc), kernel_size=tuple(layer.kernel_size),
                                                   stride=tuple(layer.stride), padding=layer.padding),
              "activation": _activation_name(layer.activation)}
    elif isinstance(layer, Layer_MaxPooling2D):
        return {"cls": "Layer_MaxPooling2D", "args": dict(pool_size=tuple(layer.pool_size), strides=tuple(layer.strides),
                                                          padding=layer.padding)}
    elif isinstance(layer, Layer_Flatten):
        return {"cls": "Layer_Flatten", "args": {}}
    elif isinstance(layer, Layer_Dropout):
        return {"cls": "Layer_Dropout", "args": dict(rate=1 - layer.rate, seed=layer.seed), "calls": layer.calls}
    elif isinstance(layer, (Activation_ReLU, Activation_Softmax, Activation_Sigmoid, Activation_Linear)):
        return {"cls": type(layer).__name__, "args": {}}
    else:
        raise TypeError(f"Model.save: no checkpoint recipe for {type(layer).__name__}")
    st["weights"] = np.asarray(layer.weights)
    st["biases"] = np.asarray(layer.biases)
    return st


def build_layer(st):
    from .src import activations, layers
    cls = getattr(layers, st["cls"], None) or getattr(activations, st["cls"])
    args = dict(st["args"])
    if "activation" in st:
        args["activation"] = _activation(st["activation"])
    layer = cls(**args)
    if "weights" in st:
        layer.set_parameters(st["weights"], st["biases"])
    if "calls" in st:
        layer.calls = st["calls"]
    return layer


_OPT_ATTRS = ("learning_rate", "current_learning_rate", "decay", "iterations", "epsilon", "beta_1", "beta_2", "beta1", "beta2",
              "rho", "momentum_factor", "check_gradients")


_QUIRKS = ("duplicate_trainable_registration", "dense_double_batch_division", "conv_bias_in_forward", "conv_backward")


def _quirks():
    """The reference-compatibility switches (neuromah_b200/config.py) the model was finalized under: they change the
    arithmetic of a step (e.g. how many times the optimizer is applied per step), so a checkpoint carries them."""
    from . import config
    return {k: getattr(config, k) for k in _QUIRKS if hasattr(config, k)}


def model_state(model):
    opt = model.optimizer
    return {
        "format": FORMAT,
        "mode": getattr(model, "mode", "fp32"),
        "quirks": _quirks(),
        "layers": [layer_state(l) for l in model.layers],
        "loss": type(model.loss).__name__ if not isinstance(model.loss, type) else model.loss.__name__,
        "accuracy": {"cls": type(model.accuracy).__name__ if not isinstance(model.accuracy, type) else model.accuracy.__name__,
                     "binary": getattr(model.accuracy, "binary", False),
                     "precision": getattr(model.accuracy, "precision", None)},      # Accuracy_Regression: std(y) / 250 of the run
        "optimizer": {"cls": type(opt).__name__, "attrs": {k: getattr(opt, k) for k in _OPT_ATTRS if hasattr(opt, k)},
                      "state": [np.asarray(s) for s in (model.arena.state if model.arena else [])]},
    }


def model_from_state(st):
    from .src import losses, metrics, optimizers
    from .src.core import Model
    if st.get("format") != FORMAT:
        raise ValueError("not a neuromah_b200 model checkpoint")
    model = Model()
    for ls in st["layers"]:
        model.add(build_layer(ls))
    o = st["optimizer"]
    opt = getattr(optimizers, o["cls"])()
    for k, v in o["attrs"].items():
        setattr(opt, k, v)
    acc_cls = getattr(metrics, st["accuracy"]["cls"])
    model.set(loss=getattr(losses, st["loss"]), optimizer=opt, accuracy=acc_cls)
    from . import config
    saved = {k: getattr(config, k) for k in st.get("quirks", {}) if hasattr(config, k)}
    try:                                  # finalize under the quirk settings the checkpoint was written with
        for k, v in st.get("quirks", {}).items():
            if hasattr(config, k):
                setattr(config, k, v)
        model.finalize(mode=st["mode"])
    finally:
        for k, v in saved.items():
            setattr(config, k, v)
    if hasattr(model.accuracy, "binary"):
        model.accuracy.binary = st["accuracy"].get("binary", False)
    if hasattr(model.accuracy, "precision"):
        model.accuracy.precision = st["accuracy"].get("precision")
    if o["state"]:
        model.arena.ensure_state(len(o["state"]))
        for slab, host in zip(model.arena.state, o["state"]):
            slab.copy_from(host)
    return model


def save(model, path):
    with open(path, "wb") as f:
        pickle.dump(model_state(model), f)


def load(path):
    with open(path, "rb") as f:
        return model_from_state(pickle.load(f))
