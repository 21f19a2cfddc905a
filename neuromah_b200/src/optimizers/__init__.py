from .SGD import Optimizer_SGD
from .Adam import Optimizer_Adam
from .RMSprop import Optimizer_RMSprop
from .Adagrad import Optimizer_Adagrad
from .Nadam import Optimizer_Nadam
