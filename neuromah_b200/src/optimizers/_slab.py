"""Shared plumbing of the slab optimizers that keep their state keyed by the BARE parameter name, as the reference's
RMSprop / Adagrad / Nadam do (RMSprop.py:99-101, Adagrad.py, Nadam.py: ``self.caches[param_name]``).  With several
differently shaped layers the reference's shared buffers stop broadcasting and it raises; ``update_params`` here fails
the same way.  The arena path (``update_arena``, what ``Model.train`` uses) keeps one state slab laid out like the
parameter arena, i.e. the per-tensor state the reference intended."""
import numpy as np

from ...device import DeviceArray


def checked_parameters(layer, require_nonempty=True):
    if not hasattr(layer, "get_parameters") or not callable(getattr(layer, "get_parameters")):
        raise AttributeError("The layer must have a 'get_parameters' method.")
    parameters = layer.get_parameters()
    if not isinstance(parameters, dict):
        raise TypeError("The 'get_parameters' method must return a dictionary.")
    if require_nonempty and not parameters:
        raise ValueError("The parameters dictionary is empty.")
    out = []
    for name, values in parameters.items():
        if not isinstance(name, str):
            raise TypeError(f"Parameter name '{name}' must be a string.")
        if not isinstance(values, tuple) or len(values) != 2:
            raise ValueError(f"Parameter values for '{name}' must be a tuple of (parameter, gradient).")
        param, grad = values
        if not isinstance(param, DeviceArray) or not isinstance(grad, DeviceArray):
            raise TypeError(f"Parameter and gradient for '{name}' must be device arrays.")
        if param.shape != grad.shape:
            raise ValueError(f"Parameter and gradient for '{name}' must have the same shape.")
        out.append((name, param, grad))
    return out


def named_state(store, name, param):
    """State buffer keyed by the bare name (reference behaviour); a shape clash is the reference's broadcast error."""
    if name not in store:
        store[name] = DeviceArray.zeros(param.shape)
    if store[name].shape != param.shape:
        raise ValueError(f"operands could not be broadcast together: state '{name}' {store[name].shape} vs parameter "
                         f"{param.shape} (state is keyed by the bare parameter name, as in the reference)")
    return store[name]


class NonFiniteFlag:
    """Device flag raised by the kernel when a gradient is NaN/Inf -> ValueError, as RMSprop.py:93-96 / Adagrad.py."""

    def __init__(self):
        self.flag = None

    def ptr(self):
        if self.flag is None:
            self.flag = DeviceArray.zeros((1,), np.int32)
        return self.flag.p

    def check(self, what):
        if self.flag is not None and int(self.flag.numpy()[0]):
            self.flag.fill_zero()
            raise ValueError(f"Gradient for '{what}' contains NaN/Inf values.")
