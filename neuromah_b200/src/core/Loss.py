"""Loss base class (reference: src/core/Loss.py:5-98): data loss, running sums, L1/L2 term."""
import numpy as np


class Loss:
    def __init__(self, model):
        self.model = model

    def regularization_loss(self):
        """L1/L2 weight penalties over model.trainable_layers (Loss.py:8-34).  The list holds
        every layer twice, so the reference's value is doubled -- reproduced as is."""
        total = 0
        for layer in self.model.trainable_layers:
            l1 = getattr(layer, "weight_regularizer_l1", 0)
            l2 = getattr(layer, "weight_regularizer_l2", 0)
            if l1 > 0 or l2 > 0:
                w = np.asarray(layer.weights, dtype=np.float64)
                if l1 > 0:
                    total += l1 * np.sum(np.abs(w))
                if l2 > 0:
                    total += l2 * np.sum(w * w)
        return total

    def calculate(self, output, y, *, include_regularization=False):
        sample_losses = np.asarray(self.forward(output, y), dtype=np.float64)
        data_loss = np.mean(sample_losses)
        self.accumulated_sum += np.sum(sample_losses)
        self.accumulated_count += len(sample_losses)
        if not include_regularization:
            return {"data_loss": data_loss}
        return {"data_loss": data_loss, "regularization_loss": self.regularization_loss()}

    def calculate_accumulated(self, *, include_regularization=False):
        data_loss = self.accumulated_sum / self.accumulated_count
        if not include_regularization:
            return {"data_loss": data_loss}
        return {"data_loss": data_loss, "regularization_loss": self.regularization_loss()}

    def new_pass(self):
        self.accumulated_sum = 0
        self.accumulated_count = 0

    def forward(self, y_pred, y_true):
        raise NotImplementedError("Derived classes must implement the 'forward' method")

    def backward(self, dvalues, y_true):
        raise NotImplementedError("Derived classes must implement the 'backward' method.")
