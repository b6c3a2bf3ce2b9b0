from .WeightNormalizedConv import WeightNormalizedConv2d, WeightNormalizedConvTranspose2d
from .WeightNormalizedLinear import WeightNormalizedLinear
from .TPReLU import TPReLU
from .View import View

__all__ = ["WeightNormalizedConv2d", "WeightNormalizedConvTranspose2d", "WeightNormalizedLinear",
           "TPReLU", "View"]
