"""The activation interface (reference: src/core/Activation.py:1-38).  It lives here so that the activations do not
import the ``core`` package (whose ``Model`` imports the layers, which import the activations); ``core.Activation``
re-exports it under the reference's path."""


class Activation:
    def forward(self, inputs, training):
        raise NotImplementedError("Derived classes must implement the 'forward' method.")

    def backward(self, dvalues):
        raise NotImplementedError("Derived classes must implement the 'backward' method.")

    def predictions(self, outputs):
        raise NotImplementedError("Derived classes must implement the 'predictions' method.")
