"""Layer_Dense on the B200 (reference: src/layers/Dense.py:6-121).

Same constructor, attributes and call shapes as the reference; ``weights``
(in,out), ``biases`` (1,out), ``dweights``, ``dbiases``, ``inputs``, ``output``
and ``dinputs`` are DeviceArrays (float32, reference layout).  Arithmetic:
nm_dense_fprop_f32 / nm_dense_dgrad_f32 / nm_dense_wgrad_f32.
"""
import numpy as np

from ...device import DeviceArray, as_device, is_array
from ... import ops
from ..initializers.Initializer import resolve
from ..activations.Relu import Activation_ReLU


class Layer_Dense:
    def __init__(self, n_inputs, n_neurons, weight_initializer=None, activation=None,
                 weight_regularizer_l1=0, weight_regularizer_l2=0, xp=np):
        self.xp = np
        if not isinstance(n_inputs, int) or n_inputs <= 0:
            raise ValueError("n_inputs must be a positive integer")
        if not isinstance(n_neurons, int) or n_neurons <= 0:
            raise ValueError("n_neurons must be a positive integer")
        # anything that is neither None, a name nor an Initializer: TypeError, as the reference actually raises -- its
        # isinstance test against the Initializer MODULE (Dense.py:3, :49) fails before the ValueError of :52 is reached.
        # (For the same reason the reference cannot take an Initializer INSTANCE at all; the mirror does.)
        self.weight_initializer = resolve(weight_initializer, TypeError)
        self.weights = DeviceArray.from_numpy(self.weight_initializer.initialize((n_inputs, n_neurons)), np.float32)
        self.biases = DeviceArray.zeros((1, n_neurons))
        self.dweights = DeviceArray.zeros((n_inputs, n_neurons))
        self.dbiases = DeviceArray.zeros((1, n_neurons))
        self.activation = activation
        self.weight_regularizer_l1 = weight_regularizer_l1
        self.weight_regularizer_l2 = weight_regularizer_l2
        # divisor of the parameter gradients; a data-parallel trainer sets this to the
        # GLOBAL batch so that a SUM all-reduce reproduces the single-GPU result
        self.grad_batch = None

    def _fused_relu(self):
        return type(self.activation) is Activation_ReLU

    def forward(self, inputs, training):
        if not is_array(inputs):
            raise TypeError("inputs must be a numpy array")
        if inputs.ndim != 2 or inputs.shape[1] != self.weights.shape[0]:
            raise ValueError(f"Input features mismatch. Expected {self.weights.shape[0]}, got {inputs.shape[-1]}")
        if isinstance(inputs, np.ndarray) and (np.any(np.isnan(inputs)) or np.any(np.isinf(inputs))):
            raise ValueError("Input contains NaN/Inf values")
        self.inputs = as_device(inputs)
        if self._fused_relu():                      # bias + ReLU in the GEMM epilogue
            self.output = ops.dense_fprop(self.inputs, self.weights, self.biases, True, holder=self)
            self.activation.inputs = self.activation.output = self.output
            return
        self.output = ops.dense_fprop(self.inputs, self.weights, self.biases, False, holder=self)
        if self.activation is not None:
            self.activation.forward(self.output, training=training)
            self.output = self.activation.output

    def backward(self, dvalues):
        dvalues = as_device(dvalues)
        if self._fused_relu():
            dvalues = ops.relu_bwd(dvalues, self.output, holder=self)      # out>0 <=> pre-activation>0
            self.activation.dinputs = dvalues
        elif self.activation is not None:
            self.activation.backward(dvalues)
            dvalues = self.activation.dinputs
        batch = dvalues.shape[0] if self.grad_batch is None else self.grad_batch
        ops.dense_wgrad(self.inputs, dvalues, self.weights, self.dweights, self.dbiases, 1.0 / batch,
                        self.weight_regularizer_l1, self.weight_regularizer_l2)
        self.dinputs = ops.dense_dgrad(dvalues, self.weights, holder=self)

    def get_parameters(self):
        return {"weights": (self.weights, self.dweights), "biases": (self.biases, self.dbiases)}

    def set_parameters(self, weights, biases):
        if tuple(weights.shape) != self.weights.shape:
            raise ValueError(f"Expected weights shape {self.weights.shape}, got {weights.shape}")
        if tuple(biases.shape) != self.biases.shape:
            raise ValueError(f"Expected biases shape {self.biases.shape}, got {biases.shape}")
        self.weights.copy_from(weights)       # in place: the arena / optimizer state stay bound
        self.biases.copy_from(biases)
