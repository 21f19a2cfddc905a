"""Activation_Softmax on the device (reference: src/activations/Softmax.py:7-22)."""
import numpy as np

from ...device import as_device
from ... import ops


class Activation_Softmax:
    def forward(self, inputs, training=False):
        self.inputs = as_device(inputs)
        if self.inputs.ndim != 2:
            raise ValueError("Activation_Softmax expects a (batch, classes) input")
        self.output = ops.softmax_fwd(self.inputs, holder=self)

    def backward(self, dvalues):
        # closed form of the per-sample Jacobian loop (Softmax.py:13-18): s * (d - <s, d>)
        self.dinputs = ops.softmax_bwd(self.output, as_device(dvalues), holder=self)

    def predictions(self, outputs):
        return np.argmax(np.asarray(outputs), axis=1)
