"""Regression accuracy (reference: src/metrics/Accuracy_Regression.py:4-17): a hit is a prediction within
``precision = std(y) / 250`` of its target; the precision is fixed by the first ``init`` unless ``reinit``."""
import numpy as np

from ..core.Accuracy import Accuracy


class Accuracy_Regression(Accuracy):
    def __init__(self, model):
        super().__init__(model)
        self.precision = None

    def init(self, y, reinit=False):
        if self.precision is None or reinit:
            self.precision = np.std(np.asarray(y)) / 250

    def compare(self, predictions, y):
        return np.absolute(np.asarray(predictions) - np.asarray(y)) < self.precision
