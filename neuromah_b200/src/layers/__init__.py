from .Conv_wrapepr import Layer_Conv2D
from .Dense import Layer_Dense
from .Dropout import Layer_Dropout
from .Input import Layer_Input
from .Flatten import Layer_Flatten
from .MaxPooling2D import Layer_MaxPooling2D

__all__ = ["Layer_Dense", "Layer_Dropout", "Layer_Input", "Layer_Conv2D", "Layer_Flatten", "Layer_MaxPooling2D"]
