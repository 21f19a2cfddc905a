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


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir("/root/reference/src")
