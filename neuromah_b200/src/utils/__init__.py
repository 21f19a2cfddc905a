"""Utilities (reference package: src/utils): the initializers re-exported under their reference names and TensorMonitor."""
from .TensorMonitor import TensorMonitor
from ..initializers.Initializer import (CustomInitializer, HeInitializer, RandomNormalInitializer,
                                        XavierInitializer)

__all__ = ["TensorMonitor", "CustomInitializer", "HeInitializer", "RandomNormalInitializer", "XavierInitializer"]
