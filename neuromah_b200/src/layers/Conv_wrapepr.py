"""Layer_Conv2D on the B200 (reference: src/layers/Conv_wrapepr.py:10-149).

forward  = nm_conv2d_fprop_f32: implicit GEMM with the 'same' zero padding folded in
           (no np.pad copy, no materialised im2col) and the ReLU fused in the epilogue.
backward = ReLU mask + nm_conv2d_wgrad_f32 (+ bias gradient) + nm_conv2d_dgrad_f32.
The reference's shipped backward is a stub (zeros); ``config.conv_backward = "stub"``
reproduces it, the default "wired" computes the intended gradients.
Stride is validated and otherwise ignored, as in the reference (Q5: backend is stride 1).
"""
import numpy as np

from ... import config, ops
from ...device import DeviceArray, as_device, is_array
from ..initializers.Initializer import resolve
from ..activations.Relu import Activation_ReLU


def _pair(v, what, positive=False):
    if isinstance(v, int):
        if positive and v <= 0:
            raise ValueError(f"{what} must be positive integer or tuple of positive integers")
        return (v, v)
    if isinstance(v, tuple) and len(v) == 2:
        if positive and not all(s > 0 for s in v):
            raise ValueError(f"{what} must be positive integer or tuple of positive integers")
        return v
    raise ValueError(f"{what} must be int or tuple of two ints")


class Layer_Conv2D:
    def __init__(self, in_channels, out_channels, kernel_size, weight_initializer=None,
                 activation=None, stride=1, padding="valid", xp=np):
        if not isinstance(in_channels, int) or in_channels <= 0:
            raise ValueError("in_channels must be a positive integer")
        if not isinstance(out_channels, int) or out_channels <= 0:
            raise ValueError("out_channels must be a positive integer")
        self.kernel_size = _pair(kernel_size, "kernel_size")
        if not all(isinstance(k, int) and k > 0 for k in self.kernel_size):
            raise ValueError("kernel_size must be positive")      # tests/layers/test_Conv2dBackend.py:68-70
        self.xp = np
        self.stride = _pair(stride, "stride", positive=True)
        if padding not in ("valid", "same"):
            raise ValueError("padding must be 'valid' or 'same'")
        self.padding = padding
        self.weight_initializer = resolve(weight_initializer, TypeError)
        shape = (out_channels, in_channels, *self.kernel_size)
        self.weights = DeviceArray.from_numpy(self.weight_initializer.initialize(shape), np.float32)
        self.biases = DeviceArray.zeros((out_channels, 1))
        self.weight_gradients = DeviceArray.zeros(shape)
        self.bias_gradients = DeviceArray.zeros((out_channels, 1))
        self.weight_momentums = np.zeros(shape)              # legacy attributes (Conv_wrapepr.py:69-70)
        self.bias_momentums = np.zeros((out_channels, 1))
        self.activation = activation
        self.compute_dinputs = True      # a Model clears this on its first layer (no consumer)

    def _pads(self):
        if self.padding == "same":
            (pt, pb), (pl, pr) = ops.same_padding(self.kernel_size[0]), ops.same_padding(self.kernel_size[1])
            return pt, pb, pl, pr
        return 0, 0, 0, 0

    def _calculate_padding(self, in_h, in_w):
        pt, pb, pl, pr = self._pads()
        return (pt, pb), (pl, pr)

    def _calculate_output_shape(self, in_h, in_w):
        pt, pb, pl, pr = self._pads()
        out_h = (in_h + pt + pb - self.kernel_size[0]) // self.stride[0] + 1
        out_w = (in_w + pl + pr - self.kernel_size[1]) // self.stride[1] + 1
        return out_h, out_w, (pt, pb), (pl, pr)

    def _fused_relu(self):
        return type(self.activation) is Activation_ReLU

    def forward(self, inputs, training):
        if not is_array(inputs) or inputs.ndim != 4:
            raise RuntimeError("Input and kernel must be 4D tensors")           # Conv2D.cpp:86-87
        self.inputs = as_device(inputs)
        bias = self.biases if config.conv_bias_in_forward else None          # Q4
        fused = self._fused_relu()
        self.output = ops.conv2d_fprop(self.inputs, self.weights, bias, self._pads(), fused, holder=self)
        if fused:
            self.activation.inputs = self.activation.output = self.output
        elif self.activation:
            self.activation.forward(self.output)
            self.output = self.activation.output

    def backward(self, dvalues):
        dvalues = as_device(dvalues)
        if self._fused_relu():
            dvalues = ops.relu_bwd(dvalues, self.output, holder=self)
            self.activation.dinputs = dvalues
        elif self.activation:
            self.activation.backward(dvalues)
            dvalues = as_device(self.activation.dinputs)
        if config.conv_backward == "stub":                      # the reference as shipped (Q3)
            ones = ops._buf(self, "_zero_in", self.inputs.shape).fill_zero()
            ops.conv2d_wgrad(ones, dvalues, self.weight_gradients, self.bias_gradients, self._pads())
            self.weight_gradients.fill_zero()
            self.dinputs = ones
            return
        ops.conv2d_wgrad(self.inputs, dvalues, self.weight_gradients, self.bias_gradients, self._pads())
        if self.compute_dinputs:
            self.dinputs = ops.conv2d_dgrad(dvalues, self.weights, self.inputs.shape, self._pads(), holder=self)
        else:
            self.dinputs = None

    def get_parameters(self):
        return {"weights": (self.weights, self.weight_gradients), "biases": (self.biases, self.bias_gradients)}

    def set_parameters(self, weights, biases):
        self.weights.copy_from(weights)
        self.biases.copy_from(biases)
