"""Layer_Flatten: zero-copy reshape views in HBM (reference: src/layers/Flatten.py:21-53).
NCHW C-order is kept, so feature index = c*H*W + h*W + w lines up with Dense weights."""
import numpy as np

from ...device import as_device


class Layer_Flatten:
    def __init__(self, xp=np):
        self.xp = np
        self.input_shape = self.output_shape = None
        self.dinputs = self.output = None

    def forward(self, inputs, training):
        inputs = as_device(inputs)
        self.input_shape = inputs.shape
        self.output = inputs.reshape(inputs.shape[0], -1)
        self.output_shape = self.output.shape
        return self.output

    def backward(self, dvalues):
        self.dinputs = as_device(dvalues).reshape(self.input_shape)
        return self.dinputs

    def get_parameters(self):
        return {}

    def set_parameters(self, params):
        pass

    def __repr__(self):
        return f"Layer_Flatten(input_shape={self.input_shape})"
