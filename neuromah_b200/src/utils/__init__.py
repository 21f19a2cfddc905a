from ..initializers.Initializer import CustomInitializer, HeInitializer, XavierInitializer, RandomNormalInitializer
from .TensorMonitor import TensorMonitor
