"""Accuracy base class (reference: src/core/Accuracy.py:3-52)."""
import numpy as np


class Accuracy:
    def __init__(self, model):
        self.model = model
        self.accumulated_sum = 0
        self.accumulated_count = 0

    def calculate(self, predictions, y):
        comparisons = self.compare(np.asarray(predictions), np.asarray(y))
        accuracy = np.mean(comparisons)
        self.accumulated_sum += np.sum(comparisons)
        self.accumulated_count += len(comparisons)
        return accuracy

    def calculate_accumulated(self):
        return self.accumulated_sum / self.accumulated_count

    def new_pass(self):
        self.accumulated_sum = 0
        self.accumulated_count = 0

    def compare(self, predictions, y):
        raise NotImplementedError("Derived classes must implement the 'compare' method.")
