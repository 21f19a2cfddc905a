"""Host-side weight initialisers with the reference's draws (src/initializers/Initializer.py:90-128).

Weights are created on the host from the GLOBAL ``np.random`` stream -- exactly
the draws the reference makes, so ``np.random.seed(s)`` followed by the same
layer construction order yields identical initial weights -- and uploaded to HBM
by the layer.  ``xp`` is accepted for signature compatibility and ignored.
"""
import numpy as np


class Initializer:
    def __init__(self, xp=np):
        self.xp = np

    def initialize(self, shape):
        raise NotImplementedError("Derived classes must implement the 'initialize' method.")


def _fans(shape):
    if len(shape) == 2:                      # Dense (in, out)
        return shape[0], shape[1]
    if len(shape) == 4:                      # Conv2D (OC, IC, kH, kW)
        rf = shape[2] * shape[3]
        return shape[1] * rf, shape[0] * rf
    return None


class RandomNormalInitializer(Initializer):
    """randn(*shape) * stddev   (Initializer.py:90-96)."""

    def __init__(self, stddev=0.01, xp=np):
        super().__init__(xp)
        self.stddev = stddev

    def initialize(self, shape):
        return np.random.randn(*shape) * self.stddev


class XavierInitializer(Initializer):
    """U(-l, l), l = sqrt(6 / (fan_in + fan_out))   (Initializer.py:98-113)."""

    def initialize(self, shape):
        fans = _fans(shape)
        if fans is None:
            raise ValueError("Unsupported shape for Xavier initialization.")
        limit = np.sqrt(6.0 / (fans[0] + fans[1]))
        return np.random.uniform(-limit, limit, size=shape)


class HeInitializer(Initializer):
    """N(0, sqrt(2 / fan_in))   (Initializer.py:115-128)."""

    def initialize(self, shape):
        fans = _fans(shape)
        if fans is None:
            raise ValueError("Unsupported shape for He initialization.")
        return np.random.normal(0, np.sqrt(2.0 / fans[0]), size=shape)


class CustomInitializer(Initializer):
    """User callable ``f(shape, xp=np) -> ndarray``   (Initializer.py:130-143)."""

    def __init__(self, init_func, xp=np):
        super().__init__(xp)
        if not callable(init_func):
            raise TypeError("init_func must be a callable function.")
        self.init_func = init_func

    def initialize(self, shape):
        try:
            w = self.init_func(shape, xp=np)
            if not isinstance(w, np.ndarray) or w.shape != tuple(shape):
                raise ValueError(f"Custom init_func must return a numpy array of shape {shape}.")
            return w
        except Exception as e:
            raise RuntimeError(f"Error in custom init_func: {e}")


def resolve(weight_initializer, bad_type_exc=ValueError):
    """None | 'xavier' | 'he' | Initializer -> Initializer (Dense.py:40-52, Conv_wrapepr.py:47-59)."""
    if weight_initializer is None:
        return RandomNormalInitializer()
    if isinstance(weight_initializer, str):
        key = weight_initializer.lower()
        if key == "xavier":
            return XavierInitializer()
        if key == "he":
            return HeInitializer()
        raise ValueError(f"Unknown initializer: {weight_initializer} \n implemented initializers are 'xavier' and 'he' ")
    if isinstance(weight_initializer, Initializer):
        return weight_initializer
    raise bad_type_exc("weight_initializer must be a string or an Initializer instance.")
