import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def synth(seed, n, shape, classes=10):
    """Synthetic MNIST-shaped data exactly as tests/golden/make_golden.py draws it."""
    rng = np.random.default_rng(seed)
    x = rng.random((n,) + tuple(shape), dtype=np.float32)
    y = np.eye(classes)[rng.integers(0, classes, n)]
    return x, y


def synth_learnable(seed, n, classes=10):
    """MNIST-shaped inputs = uniform noise + a fixed binary 28x28 template per class, so a loss curve
    actually falls within 200 steps (random labels would pin it at ln 10)."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, classes, n)
    t = (np.random.default_rng(12345).random((classes, 1, 28, 28), dtype=np.float32) > 0.7).astype(np.float32)
    x = 0.6 * rng.random((n, 1, 28, 28), dtype=np.float32) + 0.4 * t[lab]
    return x.astype(np.float32), np.eye(classes)[lab]


def losses_family_inputs():
    """Operands of the per-op vectors in golden/losses_family.npz (sigmoid, MSE / MAE / binary cross-entropy)."""
    rng = np.random.default_rng(71)
    return {"x": rng.standard_normal((16, 5)).astype(np.float32) * 3,
            "dv": rng.standard_normal((16, 5)).astype(np.float32),
            "pred": rng.random((16, 5), dtype=np.float32),
            "t_bin": (rng.random((16, 5)) > 0.5).astype(np.float32),
            "t_reg": rng.standard_normal((16, 5)).astype(np.float32),
            "pred4": rng.random((6, 2, 3, 4), dtype=np.float32),
            "t4": rng.random((6, 2, 3, 4), dtype=np.float32)}


# the three small models of golden/losses_family.npz: Dense(8,16,ReLU) -> Dense(16,n_out) -> output activation
LOSSES_FAMILY_CASES = {
    "mse": {"n_out": 3, "act": "linear", "loss": "mse", "acc": "regression", "opt": "adam", "lr": 0.01, "steps": 8},
    "mae": {"n_out": 3, "act": "linear", "loss": "mae", "acc": "regression", "opt": "sgd", "lr": 0.05, "steps": 6},
    "bce": {"n_out": 2, "act": "sigmoid", "loss": "bce", "acc": "binary", "opt": "adam", "lr": 0.01, "steps": 8},
}


def losses_family_case(name, batch=32):
    """(initial weights, inputs, targets) of one of the small models, drawn from fixed seeds."""
    c = LOSSES_FAMILY_CASES[name]
    rs = np.random.RandomState({"mse": 81, "mae": 82, "bce": 83}[name])
    w = [rs.randn(8, 16) * 0.3, rs.randn(16, c["n_out"]) * 0.3]
    n = batch * c["steps"]
    x = rs.rand(n, 8).astype(np.float32)
    if c["loss"] == "bce":
        y = (x[:, :2] + 0.2 * rs.randn(n, 2) > 0.5).astype(np.float64)
    else:
        y = np.stack([x[:, 0] - x[:, 1], x[:, 2] * x[:, 3], np.sin(3 * x[:, 4])], axis=1) + 0.05 * rs.randn(n, 3)
    return w, x, y


# ---- constructor differential (golden/ctor_surface.json): the same constructor calls on the reference and the mirror ----
# values are Python literals; "@relu" / "@xavier" / "@he" / "@object" stand for Activation_ReLU(), XavierInitializer(),
# HeInitializer() and a bare object() of whichever package is under test
CTOR_CASES = (
    [("Layer_Dense", {"n_inputs": a, "n_neurons": b}) for a, b in
     [(4, 3), (0, 3), (-1, 3), (2.5, 3), ("4", 3), (None, 3), (4, 0), (4, -2), (4, 1.0), (True, 3)]] +
    [("Layer_Dense", dict(n_inputs=4, n_neurons=3, **kw)) for kw in
     [{"weight_initializer": "xavier"}, {"weight_initializer": "HE"}, {"weight_initializer": "glorot"},
      {"weight_initializer": "@xavier"}, {"weight_initializer": "@object"}, {"weight_initializer": 3},
      {"activation": "@relu"}, {"weight_regularizer_l1": 0.01, "weight_regularizer_l2": 0.5}]] +
    [("Layer_Conv2D", dict(in_channels=a, out_channels=b, kernel_size=k)) for a, b, k in
     [(3, 5, 3), (3, 5, (2, 4)), (0, 5, 3), (3, -1, 3), (3.0, 5, 3), (3, "5", 3), (3, 5, [3, 3]), (3, 5, (3,)), (3, 5, (3, 3, 3)),
      (3, 5, 2.0), (3, 5, None)]] +
    [("Layer_Conv2D", dict(in_channels=3, out_channels=5, kernel_size=3, **kw)) for kw in
     [{"stride": 2}, {"stride": (1, 2)}, {"stride": 0}, {"stride": (1, 0)}, {"stride": -1}, {"stride": 1.5}, {"stride": (1, 1, 1)},
      {"padding": "same"}, {"padding": "SAME"}, {"padding": "full"}, {"padding": None},
      {"weight_initializer": "he"}, {"weight_initializer": "Xavier"}, {"weight_initializer": "lecun"},
      {"weight_initializer": "@he"}, {"weight_initializer": "@object"}, {"weight_initializer": 1.0}, {"activation": "@relu"}]] +
    [("Layer_MaxPooling2D", kw) for kw in
     [{"pool_size": 2}, {"pool_size": (2, 3)}, {"pool_size": [2, 2]}, {"pool_size": (2,)}, {"pool_size": 2.0}, {"pool_size": None},
      {"pool_size": 2, "strides": 1}, {"pool_size": 2, "strides": (1, 2)}, {"pool_size": 2, "strides": [1, 1]},
      {"pool_size": 2, "strides": 1.0}, {"pool_size": 3, "strides": None, "padding": "same"},
      {"pool_size": 2, "padding": "full"}, {"pool_size": 2, "padding": "VALID"}]] +
    [("Layer_Dropout", {"rate": r}) for r in (0.0, 0.1, 0.5, 0.9)] +
    [("Layer_Flatten", {})]
)


def ctor_outcome(cls, kwargs, ns):
    """Exception type name, or {attribute: descriptor} of the constructed object: arrays by shape, plain values by value,
    anything else by type name.  ``ns`` resolves the "@..." tokens to objects of the package under test."""
    kw = {k: (ns[v[1:]]() if isinstance(v, str) and v.startswith("@") else v) for k, v in kwargs.items()}
    try:
        obj = cls(**kw)
    except Exception as e:          # noqa: BLE001 - the exception TYPE is what is compared
        return type(e).__name__
    out = {}
    for k, v in vars(obj).items():
        if isinstance(getattr(v, "shape", None), tuple) and hasattr(v, "dtype"):
            out[k] = ["array", [int(d) for d in v.shape]]
        elif v is None or isinstance(v, (bool, int, float, str)):
            out[k] = ["value", v]
        elif isinstance(v, tuple):
            out[k] = ["value", list(v)]
        else:
            out[k] = ["object", type(v).__name__]
    return out


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir("/root/reference/src")
