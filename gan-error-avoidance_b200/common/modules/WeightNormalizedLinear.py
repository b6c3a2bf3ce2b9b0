"""``WeightNormalizedLinear`` — drop-in for common/modules/WeightNormalizedLinear.py:7-46.

``y = (x W^T) / sqrt(sum_in W^2 + 1e-6) [* scale] [+ bias]`` with ``weight`` (out, in),
``scale`` / ``bias`` (1, out).  Runs as the 1x1 case of the convolution kernels.
"""
import math

import torch
import torch.nn as nn

from glis_b200 import ops

__all__ = ["WeightNormalizedLinear"]


class WeightNormalizedLinear(nn.Module):
    def __init__(self, in_features, out_features, scale=True, bias=True, init_factor=1, init_scale=1):
        super(WeightNormalizedLinear, self).__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.zeros(1, out_features))
        else:
            self.register_parameter("bias", None)
        if scale:
            self.scale = nn.Parameter(torch.full((1, out_features), float(init_scale)))
        else:
            self.register_parameter("scale", None)
        self._spec = ops.ContractionSpec(False, (1, 1), (1, 1), (0, 0), (1, 1), linear=True)
        self.reset_parameters(init_factor)

    def reset_parameters(self, factor):
        bound = 1.0 * factor / math.sqrt(self.weight.size(1))
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def weight_norm(self):
        """(out, 1) row norms — the legacy keep-dim shape of the reference (:30-31)."""
        norm, _, _ = ops.wn_prepare(self.weight, None, self._spec, False, False)
        return norm.view(-1, 1)

    def norm_scale_bias(self, input):
        """Compatibility helper (:33-39); ``forward`` fuses this into the GEMM instead."""
        output = input / self.weight_norm().view(1, -1)
        if self.scale is not None:
            output = output * self.scale
        if self.bias is not None:
            output = output + self.bias
        return output

    def forward(self, input):
        return ops.wn_contraction(input, self.weight, self.scale, self.bias, self._spec)

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_features, self.out_features)
