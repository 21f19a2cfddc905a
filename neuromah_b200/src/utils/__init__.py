from ..initializers.Initializer import CustomInitializer, HeInitializer, XavierInitializer, RandomNormalInitializer
