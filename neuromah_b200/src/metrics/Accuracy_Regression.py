"""Regression accuracy (reference: src/metrics/Accuracy_Regression.py:4-17).

A prediction counts as a hit when it lies within ``precision`` of its target, where ``precision`` is 1/250 of the
standard deviation of the targets the metric was initialised with (``Model.train`` calls ``init(y)`` first, Model.py:119).
The tolerance is frozen by the first ``init``; ``reinit=True`` recomputes it.  Host arithmetic over the B x D flags.
"""
import numpy as np

from ..core.Accuracy import Accuracy

_STD_FRACTION = 250


class Accuracy_Regression(Accuracy):
    def __init__(self, model):
        super().__init__(model)
        self.precision = None

    def init(self, y, reinit=False):
        if reinit or self.precision is None:
            self.precision = np.std(np.asarray(y)) / _STD_FRACTION

    def compare(self, predictions, y):
        gap = np.abs(np.asarray(predictions) - np.asarray(y))
        return gap < self.precision
