from .Model import Model
from .Loss import Loss
from .Accuracy import Accuracy
