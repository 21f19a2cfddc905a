"""Layer_Input (reference: src/layers/Input.py): the upload point -- host batches become DeviceArrays here."""
from ...device import as_device


class Layer_Input:
    def forward(self, inputs, training):
        self.output = as_device(inputs)
