"""Optimizer_SGD on the device (reference: src/optimizers/SGD.py:4-98).

Momentum buffers are keyed by the bare parameter name exactly as in the reference
(SGD.py:85-91), so momentum with several differently shaped layers fails the
same way; the arena path keeps one momentum slab for the whole model instead.
NaN/Inf gradients raise ValueError (SGD.py:80-83) -- checked by a device flag.
"""
import numpy as np

from ... import ops
from ...device import DeviceArray


class Optimizer_SGD:
    def __init__(self, learning_rate=1., decay=0., momentum=0.):
        if not isinstance(learning_rate, (int, float)):
            raise TypeError(f"learning_rate must be a number, got {type(learning_rate)}")
        if learning_rate <= 0:
            raise ValueError(f"learning_rate should be a positive number, got {learning_rate}")
        if not isinstance(decay, (int, float)):
            raise TypeError(f"decay must be a number, got {type(decay)}")
        if decay < 0:
            raise ValueError(f"decay must be non-negative, got {decay}")
        if not isinstance(momentum, (int, float)):
            raise TypeError(f"momentum must be a number, got {type(momentum)}")
        if momentum < 0:
            raise ValueError(f"momentum must be non-negative, got {momentum}")
        self.learning_rate = learning_rate
        self.current_learning_rate = learning_rate
        self.decay = decay
        self.iterations = 0
        self.momentum_factor = momentum
        self.momentums = {}
        self.check_gradients = True
        self._flag = None

    def pre_update_params(self):
        if self.decay:
            self.current_learning_rate = self.learning_rate * (1. / (1. + self.decay * self.iterations))

    def _nonfinite_flag(self):
        if self._flag is None:
            self._flag = DeviceArray.zeros((1,), np.int32)
        return self._flag

    def _raise_if_nonfinite(self, what):
        if self.check_gradients and int(self._flag.numpy()[0]):
            self._flag.fill_zero()
            raise ValueError(f"Gradient for '{what}' contains NaN/Inf values.")

    def update_params(self, layer):
        if not hasattr(layer, "get_parameters") or not callable(getattr(layer, "get_parameters")):
            raise AttributeError(f"The layer: {layer} does not implement get_parameters method")
        parameters = layer.get_parameters()
        if not isinstance(parameters, dict):
            raise TypeError("The get_parameters() method should return a Dict")
        for name, values in parameters.items():
            if not isinstance(name, str):
                raise TypeError(f"The parameter name for the layer:{layer} should be a string")
            if not isinstance(values, tuple) or len(values) != 2:
                raise ValueError("The parameter values must be a tuple: (parameter, gradients)")
            param, grad = values
            if not isinstance(param, DeviceArray) or not isinstance(grad, DeviceArray):
                raise TypeError("The parameters and gradients for the layer must be device arrays")
            if param.shape != grad.shape:
                raise ValueError("Parameter and gradient must have the same shape")
            mom = None
            if self.momentum_factor:
                if name not in self.momentums:
                    self.momentums[name] = DeviceArray.zeros(param.shape)
                mom = self.momentums[name]
                if mom.shape != param.shape:
                    raise ValueError("momentum buffer shape mismatch (momentums are keyed by bare name, SGD.py:85-91)")
            flag = self._nonfinite_flag() if self.check_gradients else None
            ops.sgd_step(param, grad, mom, lr=self.current_learning_rate, momentum=self.momentum_factor,
                         n_apply=1, nonfinite=flag)
            self._raise_if_nonfinite(name)

    def bind_arena(self, arena):
        if self.momentum_factor:
            arena.ensure_state(1)

    def update_arena(self, arena, n_apply=1):
        mom = arena.state[0] if self.momentum_factor else None
        flag = self._nonfinite_flag() if self.check_gradients else None
        ops.sgd_step(arena.params, arena.grads, mom, lr=self.current_learning_rate,
                     momentum=self.momentum_factor, n_apply=n_apply, nonfinite=flag)

    def check_nonfinite(self):
        """Arena path: the kernel raised the device flag if any gradient of the slab was NaN/Inf (SGD.py:80-83).  Reading
        it costs a device synchronisation, so Model.train calls this once per epoch beside the loss read-back."""
        if self.check_gradients and self._flag is not None:
            self._raise_if_nonfinite("arena")

    def post_update_params(self):
        self.iterations += 1
