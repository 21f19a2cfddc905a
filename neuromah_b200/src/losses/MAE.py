"""Loss_MeanAbsoluteError on the device (reference: src/losses/MAE.py:4-28)."""
from ... import ops
from ...device import as_device
from ..core.Loss import Loss


class Loss_MeanAbsoluteError(Loss):  # L1 loss
    def __init__(self, model):
        super().__init__(model)

    def forward(self, y_pred, y_true):
        """Per-sample mean of |y_true - y_pred| over the non-batch dimensions (:13)."""
        return ops.loss_fwd(ops.LOSS_MAE, as_device(y_pred), as_device(y_true), holder=self)

    def backward(self, dvalues, y_true):
        """sign(y_true - dvalues) / outputs-per-sample, with the reference's sign (:23-28)."""
        self.dinputs = ops.loss_bwd(ops.LOSS_MAE, as_device(dvalues), as_device(y_true), holder=self)
