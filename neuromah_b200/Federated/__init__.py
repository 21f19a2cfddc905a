from .utils import federated_average, federated_average_device
from .FLTrainer import FLTrainer
