"""Layers of the hot path under the reference's class names (reference package: src/layers).
Layer_RNN and BatchNormalization2D are not part of this package (the reference imports an RNN module it does not ship;
its batch-norm file is unreachable -- SURVEY.md section 2)."""
from .Conv_wrapepr import Layer_Conv2D
from .Dense import Layer_Dense
from .Dropout import Layer_Dropout
from .Flatten import Layer_Flatten
from .Input import Layer_Input
from .MaxPooling2D import Layer_MaxPooling2D

__all__ = ["Layer_Conv2D", "Layer_Dense", "Layer_Dropout", "Layer_Flatten", "Layer_Input", "Layer_MaxPooling2D"]
