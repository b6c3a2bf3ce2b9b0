"""``TPReLU`` — drop-in for the reference's common/modules/TPReLU.py:8-21.

Translated PReLU ``prelu(x - b, clamp(a, 0, 1)) + b`` with one slope ``weight`` (init .25)
and one translation ``bias`` (init 0) per channel (dim 1).  Forward and backward are single
sm_100a kernels (``glis_tprelu_forward`` / ``glis_tprelu_backward``); inside the fused
training step the same maths lives in the GEMM epilogues instead.
"""
import torch
import torch.nn as nn

from glis_b200 import ops

__all__ = ["TPReLU"]


class TPReLU(nn.Module):
    def __init__(self, num_parameters=1, init=0.25):
        super(TPReLU, self).__init__()
        self.num_parameters = num_parameters
        self.weight = nn.Parameter(torch.full((num_parameters,), float(init)))
        self.bias = nn.Parameter(torch.zeros(num_parameters))

    def forward(self, input):
        return ops.tprelu(input, self.weight, self.bias)

    def __repr__(self):
        return "%s (%d)" % (self.__class__.__name__, self.num_parameters)
