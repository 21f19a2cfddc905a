"""Running accuracy over a training / validation pass -- the base class the metrics derive from.

Mirrors the public surface of the reference's src/core/Accuracy.py (``calculate``, ``calculate_accumulated``,
``new_pass``, ``compare`` and the two ``accumulated_*`` attributes that Model.train reads); predictions and targets
may be device arrays, the comparison itself is a host-side reduction over B flags.
"""
import numpy as np


class Accuracy:
    def __init__(self, model):
        self.model = model
        self.new_pass()

    def new_pass(self):
        """Start a new pass: forget the hits and the sample count gathered so far."""
        self.accumulated_sum, self.accumulated_count = 0, 0

    def compare(self, predictions, y):
        raise NotImplementedError("Derived classes must implement the 'compare' method.")

    def calculate(self, predictions, y):
        """Fraction of hits in this batch; also folded into the running totals of the pass."""
        hits = np.asarray(self.compare(np.asarray(predictions), np.asarray(y)))
        self.accumulated_sum += hits.sum()
        self.accumulated_count += len(hits)        # samples, also when a sample has several outputs (Accuracy.py:22)
        return hits.mean()                         # over every compared element (Accuracy.py:20)

    def calculate_accumulated(self):
        return self.accumulated_sum / self.accumulated_count
