"""Categorical accuracy (reference: src/metrics/Accuracy_Categorical.py:4-16): a hit is a predicted class id equal to
the target class id; one-hot targets are reduced with argmax unless ``binary`` is set."""
import numpy as np

from ..core.Accuracy import Accuracy


class Accuracy_Categorical(Accuracy):
    def __init__(self, *, binary=False, model):
        super().__init__(model)
        self.binary = bool(binary)

    def init(self, y):
        """Hook for metrics that must see the targets before training (nothing to prepare for class ids)."""
        return None

    def compare(self, predictions, y):
        targets = np.asarray(y)
        if targets.ndim == 2 and not self.binary:
            targets = targets.argmax(axis=1)
        return np.asarray(predictions) == targets
