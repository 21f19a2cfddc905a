"""Layer_Input (reference: src/layers/Input.py): the upload point -- host batches become DeviceArrays here.

Raw ``uint8`` pixel batches are accepted as well: they cross PCIe as bytes and are normalised on the device with the
example scripts' preprocessing ``X.astype(np.float32) / 255.0`` (examples/test_Conv2D_MNIST.py:27; IEEE division, bit
identical to NumPy).  The bf16 tensor-core plans consume the bytes directly (``keep_uint8``)."""
import numpy as np

from ... import ops
from ...device import input_to_device


class Layer_Input:
    keep_uint8 = False            # set by Model.finalize when a tensor-core plan reads uint8 inputs itself
    uint8_denominator = 255.0

    def forward(self, inputs, training):
        x = input_to_device(inputs)
        if x.dtype == np.uint8 and not self.keep_uint8:
            x = ops.u8_to_f32(x, self.uint8_denominator, holder=self)
        self.output = x
