"""Loss_MeanSquaredError on the device (reference: src/losses/MSE.py:4-28)."""
from ... import ops
from ...device import as_device
from ..core.Loss import Loss


class Loss_MeanSquaredError(Loss):  # L2 loss
    def __init__(self, model):
        super().__init__(model)

    def forward(self, y_pred, y_true):
        """Per-sample mean of (y_true - y_pred)**2 over the non-batch dimensions (:13)."""
        return ops.loss_fwd(ops.LOSS_MSE, as_device(y_pred), as_device(y_true), holder=self)

    def backward(self, dvalues, y_true):
        """-2 (y_true - dvalues) / outputs-per-sample (:23-28; not divided by the number of samples)."""
        self.dinputs = ops.loss_bwd(ops.LOSS_MSE, as_device(dvalues), as_device(y_true), holder=self)
