"""Loss_CategoricalCrossentropy on the device (reference: src/losses/categorical_crossentropy.py:4-38)."""
from ... import ops
from ...device import as_device, labels_to_device
from ..core.Loss import Loss


class Loss_CategoricalCrossentropy(Loss):
    def __init__(self, model):
        super().__init__(model)

    def forward(self, y_pred, y_true):
        """Per-sample -log(clip(p, 1e-7, 1-1e-7)[y]) (:8-25); one-hot rows are read via argmax."""
        probs = as_device(y_pred)
        losses, self._correct, _ = ops.softmax_ce(probs, labels_to_device(y_true), want_grad=False, holder=self)
        return losses

    def backward(self, dvalues, y_true):
        """-onehot/p / samples (:27-38)."""
        self.dinputs = ops.cce_bwd(as_device(dvalues), labels_to_device(y_true), holder=self)
