"""Device-side optimizers under the reference's class names (src/optimizers/*.py)."""
from .Adagrad import Optimizer_Adagrad
from .Adam import Optimizer_Adam
from .Nadam import Optimizer_Nadam
from .RMSprop import Optimizer_RMSprop
from .SGD import Optimizer_SGD

__all__ = ["Optimizer_Adagrad", "Optimizer_Adam", "Optimizer_Nadam", "Optimizer_RMSprop", "Optimizer_SGD"]
