from .Initializer import (Initializer, RandomNormalInitializer, XavierInitializer, HeInitializer,
                          CustomInitializer)
