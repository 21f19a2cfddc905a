"""Metrics (reference package: src/metrics)."""
from .Accuracy_Categorical import Accuracy_Categorical
from .Accuracy_Regression import Accuracy_Regression

__all__ = ["Accuracy_Categorical", "Accuracy_Regression"]
