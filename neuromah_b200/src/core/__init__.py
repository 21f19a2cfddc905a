"""Model driver, loss base and accuracy base (reference package: src/core)."""
from .Accuracy import Accuracy
from .Loss import Loss
from .Model import Model

__all__ = ["Accuracy", "Loss", "Model"]
