"""Weight-normalized (transposed) convolutions — drop-ins for
common/modules/WeightNormalizedConv.py:9-99 of the reference.

``y = conv(x, w) / n [* scale] [+ bias]`` with ``n[o] = sqrt(c * sum w[o]^2 + 1e-6)`` and
``c = 1`` (conv) or ``1 / prod(stride)`` (transposed).  Parameter names and shapes are the
reference's: ``weight`` (Cout,Cin,kh,kw) or, transposed, (Cin,Cout,kh,kw); ``scale`` and
``bias`` (1,Cout,1,1).  The arithmetic is one packed-weight pass plus one gather-GEMM with
the bias in its epilogue; gradients of ``weight`` / ``scale`` come out of the raw filter
gradient through the closed-form projection.
"""
import math

import torch
import torch.nn as nn

from glis_b200 import ops

__all__ = ["WeightNormalizedConv2d", "WeightNormalizedConvTranspose2d"]


def _two(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class _WeightNormalizedConvNd(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation,
                 transposed, output_padding, scale, bias, init_factor, init_scale):
        super(_WeightNormalizedConvNd, self).__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = kernel_size, stride
        self.padding, self.dilation = padding, dilation
        self.transposed, self.output_padding = transposed, output_padding
        self.groups = 1
        lead = (in_channels, out_channels) if transposed else (out_channels, in_channels)
        self.weight = nn.Parameter(torch.empty(*(lead + tuple(kernel_size))))
        ones = (1,) * len(kernel_size)
        if scale:
            self.scale = nn.Parameter(torch.full((1, out_channels) + ones, float(init_scale)))
        else:
            self.register_parameter("scale", None)
        if bias:
            self.bias = nn.Parameter(torch.zeros((1, out_channels) + ones))
        else:
            self.register_parameter("bias", None)
        # the 2017 `_ConvNd.reset_parameters`: fan = in_channels * prod(kernel) in both directions
        fan = in_channels
        for k in kernel_size:
            fan *= k
        bound = 1.0 / math.sqrt(fan)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound).mul_(init_factor)
        self.weight_norm_factor = 1.0
        if transposed:
            for s in stride:
                self.weight_norm_factor = self.weight_norm_factor / s

    def _spec(self, output_padding=None):
        return ops.ContractionSpec(self.transposed, self.kernel_size, self.stride, self.padding,
                                   self.dilation,
                                   self.output_padding if output_padding is None else output_padding)

    def weight_norm(self):
        """(Cout,1,1,1) for conv, (1,Cout,1,1) transposed — the reference's keep-dim shapes (:29-38)."""
        norm, _, _ = ops.wn_prepare(self.weight, None, self._spec(), False, False)
        return norm.view(1, -1, 1, 1) if self.transposed else norm.view(-1, 1, 1, 1)

    def norm_scale_bias(self, input):
        """Compatibility helper (:40-49); ``forward`` fuses this into the GEMM instead."""
        output = input / self.weight_norm().view(1, -1, 1, 1)
        if self.scale is not None:
            output = output * self.scale
        if self.bias is not None:
            output = output + self.bias
        return output

    def __repr__(self):
        s = "%s(%d, %d, kernel_size=%s, stride=%s" % (self.__class__.__name__, self.in_channels,
                                                      self.out_channels, self.kernel_size, self.stride)
        if any(p != 0 for p in self.padding):
            s += ", padding=%s" % (self.padding,)
        if any(d != 1 for d in self.dilation):
            s += ", dilation=%s" % (self.dilation,)
        if any(p != 0 for p in self.output_padding):
            s += ", output_padding=%s" % (self.output_padding,)
        if self.scale is None:
            s += ", scale=False"
        if self.bias is None:
            s += ", bias=False"
        return s + ")"


class WeightNormalizedConv2d(_WeightNormalizedConvNd):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 scale=True, bias=True, init_factor=1, init_scale=1):
        super(WeightNormalizedConv2d, self).__init__(
            in_channels, out_channels, _two(kernel_size), _two(stride), _two(padding), _two(dilation),
            False, (0, 0), scale, bias, init_factor, init_scale)

    def forward(self, input):
        return ops.wn_contraction(input, self.weight, self.scale, self.bias, self._spec())


class WeightNormalizedConvTranspose2d(_WeightNormalizedConvNd):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0,
                 scale=True, bias=True, dilation=1, init_factor=1, init_scale=1):
        super(WeightNormalizedConvTranspose2d, self).__init__(
            in_channels, out_channels, _two(kernel_size), _two(stride), _two(padding), _two(dilation),
            True, _two(output_padding), scale, bias, init_factor, init_scale)

    def _output_padding(self, input, output_size):
        if output_size is None:
            return self.output_padding
        want = list(output_size)[-2:]
        pads = []
        for d in range(2):
            smallest = ((input.size(d + 2) - 1) * self.stride[d] - 2 * self.padding[d]
                        + self.dilation[d] * (self.kernel_size[d] - 1) + 1)
            extra = want[d] - smallest
            if extra < 0 or extra >= max(self.stride[d], self.dilation[d]):
                raise ValueError("requested output size %s is not reachable from input %s"
                                 % (tuple(want), tuple(input.shape[2:])))
            pads.append(extra)
        return tuple(pads)

    def forward(self, input, output_size=None):
        spec = self._spec(self._output_padding(input, output_size))
        return ops.wn_contraction(input, self.weight, self.scale, self.bias, spec)
