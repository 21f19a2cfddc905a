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


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir("/root/reference/src")
