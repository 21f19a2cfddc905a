"""Model driver, loss base and accuracy base (reference package: src/core)."""
from .Accuracy import Accuracy
from .Activation import Activation
from .Loss import Loss
from .Model import Model

__all__ = ["Accuracy", "Activation", "Loss", "Model"]
