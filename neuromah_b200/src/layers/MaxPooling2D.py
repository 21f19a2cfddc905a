"""Layer_MaxPooling2D on the B200 (reference: src/layers/MaxPooling2D.py:4-84 and
CPU_kernels/maxpooling_backend.cpp).  ``max_indices`` keeps the reference's
int32 [N,C,OH,OW,2] (row, col)-in-padded-coordinates format."""
from ... import ops
from ...device import as_device, is_array


class Layer_MaxPooling2D:
    def __init__(self, pool_size, strides=None, padding="valid"):
        if isinstance(pool_size, int):
            self.pool_size = (pool_size, pool_size)
        elif isinstance(pool_size, tuple) and len(pool_size) == 2:
            self.pool_size = pool_size
        else:
            raise ValueError("pool_size must be an int or a tuple of two ints")
        if strides is None:
            self.strides = self.pool_size
        elif isinstance(strides, int):
            self.strides = (strides, strides)
        elif isinstance(strides, tuple) and len(strides) == 2:
            self.strides = strides
        else:
            raise ValueError("strides must be an int or a tuple of two ints")
        if padding not in ("valid", "same"):
            raise ValueError("padding must be 'valid' or 'same'")
        self.padding = padding

    def forward(self, inputs, training=True):
        if not is_array(inputs) or inputs.ndim != 4:
            raise RuntimeError("Input must be a 4D tensor")
        self.inputs = as_device(inputs)
        self.output, self.max_indices = ops.maxpool_fwd(self.inputs, *self.pool_size, *self.strides,
                                                        self.padding, holder=self)
        return self.output

    def backward(self, dvalues):
        dvalues = as_device(dvalues)
        if dvalues.ndim != 4:
            raise RuntimeError("dvalues must be a 4D tensor")
        self.dinputs = ops.maxpool_bwd(dvalues, self.max_indices, self.inputs.shape, *self.pool_size,
                                       *self.strides, self.padding, holder=self)
        return self.dinputs
