from .Relu import Activation_ReLU
from .Softmax import Activation_Softmax
from .Sigmoid import Activation_Sigmoid
from .Linear import Activation_Linear
