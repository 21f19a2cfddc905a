"""Layer_Dropout on the device (reference: src/layers/Dropout.py:3-40).

Same attributes: ``rate`` is the KEEP probability (Dropout.py:12: ``self.rate = 1 - rate``), ``binary_mask`` holds
0 or 1/keep per element, ``output`` / ``dinputs`` as usual.  The reference draws the mask with NumPy's global RNG
(``np.random.binomial``, Dropout.py:27); here it comes from a counter-based Philox4x32-10 stream on the device
(``seed`` + one offset per forward call), so the masks are reproducible but not the same bits as NumPy's.  For
bit-for-bit comparisons with the reference set ``fixed_mask`` to the reference's ``binary_mask``.
"""
import numpy as np

from ..._lib import lib
from ...device import DeviceArray, as_device, stream


class Layer_Dropout:
    _next_seed = 0x5EED

    def __init__(self, rate, seed=None):
        self.rate = 1 - rate
        if not (0 < self.rate <= 1):
            raise ValueError("dropout rate must be in [0, 1)")
        if seed is None:
            seed = Layer_Dropout._next_seed
            Layer_Dropout._next_seed += 1
        self.seed = int(seed)
        self.calls = 0                  # Philox stream offset: one per training forward
        self.fixed_mask = None
        self.inputs = self.output = self.binary_mask = self.dinputs = None

    def forward(self, inputs, training):
        inputs = as_device(inputs)
        self.inputs = inputs
        if not training:
            self.output = inputs.copy()                  # Dropout.py:23-25
            return self.output
        out = DeviceArray(inputs.shape, np.float32)
        if self.fixed_mask is not None:
            self.binary_mask = as_device(np.ascontiguousarray(self.fixed_mask, dtype=np.float32)
                                         if not isinstance(self.fixed_mask, DeviceArray) else self.fixed_mask)
            if self.binary_mask.shape != inputs.shape:
                raise ValueError("fixed_mask must have the shape of the inputs")
            lib.nm_mul_f32(out.p, inputs.p, self.binary_mask.p, inputs.size, stream())
        else:
            if self.binary_mask is None or self.binary_mask.shape != inputs.shape:
                self.binary_mask = DeviceArray(inputs.shape, np.float32)
            lib.nm_dropout_fwd_f32(inputs.p, out.p, self.binary_mask.p, inputs.size, float(self.rate), self.seed, self.calls,
                                   stream())
            self.calls += 1
        self.output = out
        return self.output

    def backward(self, dvalues):
        dvalues = as_device(dvalues)
        self.dinputs = DeviceArray(dvalues.shape, np.float32)
        lib.nm_mul_f32(self.dinputs.p, dvalues.p, self.binary_mask.p, dvalues.size, stream())   # Dropout.py:40
        return self.dinputs

    def get_parameters(self):
        return {}

    def set_parameters(self, params):
        pass
