"""Fused softmax + cross-entropy gradient (p - onehot)/B on the device
(reference: src/losses/Activation_Softmax_Loss_CategoricalCrossentropy.py:3-19)."""
from ... import ops
from ...device import as_device, labels_to_device


class Activation_Softmax_Loss_CategoricalCrossentropy:
    def __init__(self):
        self.grad_batch = None       # global batch under data parallelism

    def backward(self, dvalues, y_true):
        probs = as_device(dvalues)
        inv = None if self.grad_batch is None else 1.0 / self.grad_batch
        _, _, self.dinputs = ops.softmax_ce(probs, labels_to_device(y_true), want_grad=True,
                                            inv_batch=inv, holder=self)
