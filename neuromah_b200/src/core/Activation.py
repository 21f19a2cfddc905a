"""Activation base class under the reference's import path (src/core/Activation.py)."""
from ..activations._base import Activation

__all__ = ["Activation"]
