"""Activation_Sigmoid on the device (reference: src/activations/Sigmoid.py:4-14)."""
import numpy as np

from ...device import as_device
from ... import ops
from ._base import Activation


class Activation_Sigmoid(Activation):
    def forward(self, inputs, training=False):
        self.inputs = as_device(inputs)
        self.output = ops.sigmoid_fwd(self.inputs, holder=self)            # 1 / (1 + exp(-x)) (:7)

    def backward(self, dvalues):
        self.dinputs = ops.sigmoid_bwd(as_device(dvalues), self.output, holder=self)     # dy * (1 - out) * out (:10)

    def predictions(self, outputs):
        return (np.asarray(outputs) > 0.5) * 1                              # threshold at 0.5 (:13-14)
