"""Activation_ReLU on the device (reference: src/activations/Relu.py:7-17)."""
from ...device import as_device
from ... import ops


class Activation_ReLU:
    def forward(self, inputs, training=False):
        self.inputs = as_device(inputs)
        self.output = ops.relu_fwd(self.inputs, holder=self)

    def backward(self, dvalues):
        # mask from the pre-activation inputs: inputs <= 0 -> 0 (Relu.py:12-14)
        self.dinputs = ops.relu_bwd(as_device(dvalues), self.inputs, holder=self)

    def predictions(self, outputs):
        return outputs
