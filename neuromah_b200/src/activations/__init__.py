from .Relu import Activation_ReLU
from .Softmax import Activation_Softmax
