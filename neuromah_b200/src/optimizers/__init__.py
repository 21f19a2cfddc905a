from .SGD import Optimizer_SGD
from .Adam import Optimizer_Adam
