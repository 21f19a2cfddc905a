"""Metrics (reference package: src/metrics)."""
from .Accuracy_Categorical import Accuracy_Categorical

__all__ = ["Accuracy_Categorical"]
