"""Loss_BinaryCrossentropy on the device (reference: src/losses/binary_crossentropy.py:4-33)."""
from ... import ops
from ...device import as_device
from ..core.Loss import Loss


class Loss_BinaryCrossentropy(Loss):
    def __init__(self, model):
        super().__init__(model)

    def forward(self, y_pred, y_true):
        """Per-sample mean of -(y log p + (1 - y) log(1 - p)), p clipped to [1e-7, 1 - 1e-7] (:9-18)."""
        return ops.loss_fwd(ops.LOSS_BCE, as_device(y_pred), as_device(y_true), holder=self)

    def backward(self, dvalues, y_true):
        """-(y / p - (1 - y) / (1 - p)) / outputs-per-sample on the clipped predictions (:20-33)."""
        self.dinputs = ops.loss_bwd(ops.LOSS_BCE, as_device(dvalues), as_device(y_true), holder=self)
