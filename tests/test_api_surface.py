"""CPU, authoring container only (needs /root/reference): the mirror package's public surface against the UNMODIFIED
reference classes, imported through oracle/ref_shim.py -- every class the reference exports around the Model.train path:
same constructor parameters (names, order, defaults, kinds) and, for every public method of the reference class, a
method of the same name on the mirror whose parameters start with the reference's (the mirror may append keyword
arguments such as ``mode`` or ``stream_batches``)."""
import importlib
import inspect

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree only in the authoring container")

SURFACE = {
    "layers": ["Layer_Conv2D", "Layer_Dense", "Layer_Dropout", "Layer_Flatten", "Layer_Input", "Layer_MaxPooling2D"],
    "activations": ["Activation_ReLU", "Activation_Softmax", "Activation_Sigmoid", "Activation_Linear"],
    "losses": ["Loss_CategoricalCrossentropy", "Loss_BinaryCrossentropy", "Loss_MeanSquaredError", "Loss_MeanAbsoluteError"],
    "optimizers": ["Optimizer_Adam", "Optimizer_SGD", "Optimizer_RMSprop", "Optimizer_Adagrad", "Optimizer_Nadam"],
    "metrics": ["Accuracy_Categorical", "Accuracy_Regression"],
    "core": ["Model", "Loss", "Accuracy", "Activation"],
    "utils": ["TensorMonitor", "CustomInitializer", "HeInitializer", "XavierInitializer", "RandomNormalInitializer"],
}

def _params(fn):
    return [(k, p.default, p.kind) for k, p in inspect.signature(fn).parameters.items() if k != "self"]


def _cases():
    return [(sub, name) for sub, names in SURFACE.items() for name in names]


@pytest.mark.parametrize("sub,name", _cases())
def test_class_surface_matches_the_reference(sub, name):
    ref_shim.load()
    ref_cls = getattr(importlib.import_module(f"{ref_shim.PKG}.src.{sub}"), name)
    cls = getattr(importlib.import_module(f"neuromah_b200.src.{sub}"), name)
    ref_ctor, ctor = _params(ref_cls.__init__), _params(cls.__init__)
    assert ctor[:len(ref_ctor)] == ref_ctor, f"{name}.__init__: {ctor} vs reference {ref_ctor}"
    assert all(p[1] is not inspect.Parameter.empty for p in ctor[len(ref_ctor):]), "extra constructor arguments need defaults"
    for meth, fn in vars(ref_cls).items():
        if meth.startswith("_") or not callable(fn):
            continue
        assert hasattr(cls, meth), f"{name}.{meth} missing from the mirror"
        mine = getattr(cls, meth)
        ref_p = [(k, kind) for k, _, kind in _params(fn)]
        my_p = [(k, kind) for k, _, kind in _params(mine)]
        # positional parameters keep name and order; keyword-only ones keep their names
        ref_pos = [p for p in ref_p if p[1] != inspect.Parameter.KEYWORD_ONLY]
        my_pos = [p for p in my_p if p[1] != inspect.Parameter.KEYWORD_ONLY]
        assert my_pos[:len(ref_pos)] == ref_pos, f"{name}.{meth}: {my_p} vs reference {ref_p}"
        ref_kw = {k for k, kind in ref_p if kind == inspect.Parameter.KEYWORD_ONLY}
        my_kw = {k for k, kind in my_p if kind == inspect.Parameter.KEYWORD_ONLY}
        assert ref_kw <= my_kw, f"{name}.{meth}: keyword arguments {sorted(ref_kw - my_kw)} missing"


def test_package_exports_match_the_reference():
    """Every name the reference's package __init__s export (layers / activations / losses / optimizers / metrics / core /
    utils) resolves on the mirror -- except Layer_RNN, which the reference imports from a file it does not ship."""
    ref_shim.load()
    for sub in SURFACE:
        ref_mod = importlib.import_module(f"{ref_shim.PKG}.src.{sub}")
        mod = importlib.import_module(f"neuromah_b200.src.{sub}")
        exported = [k for k, v in vars(ref_mod).items() if inspect.isclass(v) and v.__module__.startswith(ref_shim.PKG)]
        missing = [k for k in exported if not hasattr(mod, k) and k != "Layer_RNN"]
        assert not missing, f"src.{sub}: {missing}"


@pytest.mark.parametrize("name", ["Optimizer_Adam", "Optimizer_SGD", "Optimizer_RMSprop", "Optimizer_Adagrad", "Optimizer_Nadam"])
def test_optimizer_constructor_validation_is_the_reference(name):
    """Differential: every constructor parameter of every optimizer, set to each of 13 good / bad values, gives the same
    outcome as the unmodified reference -- the same exception type, or an object with the same hyper-parameter
    attributes (Adam.py:24-35, SGD.py:28-46, RMSprop.py / Adagrad.py / Nadam.py constructors)."""
    ref_shim.load()
    ref_cls = getattr(importlib.import_module(f"{ref_shim.PKG}.src.optimizers"), name)
    cls = getattr(importlib.import_module("neuromah_b200.src.optimizers"), name)
    attrs = ("learning_rate", "current_learning_rate", "decay", "iterations", "epsilon", "beta_1", "beta_2", "beta1", "beta2",
             "rho", "momentum", "momentum_factor")

    def outcome(c, kw):
        try:
            o = c(**kw)
        except Exception as e:          # noqa: BLE001 - the exception TYPE is the thing compared
            return type(e).__name__
        return {k: getattr(o, k, None) for k in attrs}

    params = [k for k in inspect.signature(ref_cls.__init__).parameters if k != "self"]
    assert params
    for p in params:
        for v in (0, -1, -0.5, 1, 1.0, 1.5, 2, "x", None, True, 0.0, 1e-9, 0.5):
            assert outcome(cls, {p: v}) == outcome(ref_cls, {p: v}), f"{name}({p}={v!r})"
