from .pconv import PConv2d
from .generator import PConvUNet
from .discriminator import Discriminator

__all__ = ["PConv2d", "PConvUNet", "Discriminator"]
