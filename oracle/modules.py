"""ORACLE (test infrastructure) — weight-normalized layers, restated.

Follows the reference's ``common/modules/WeightNormalizedConv.py``,
``WeightNormalizedLinear.py``, ``TPReLU.py`` and ``View.py`` with the legacy
(mid-2017 PyTorch) semantics spelled out in SURVEY.md App. B: ``sum(dim)`` keeps
the reduced dimension, no implicit broadcasting (``expand_as`` everywhere).
The maths is written here as explicit closed forms; nothing is copied.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_EPS = 1e-6  # added under the square root: WeightNormalizedConv.py:37, WeightNormalizedLinear.py:31


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def conv_weight_norm(weight, transposed, stride):
    """Per-output-channel norm as a (1, Cout, 1, 1) tensor.

    Reference: WeightNormalizedConv.py:29-38.  conv weights are (Cout, Cin, kh, kw)
    and reduce over dims 1,2,3; transposed-conv weights are (Cin, Cout, kh, kw) and
    reduce over dims 0,2,3.  ``weight_norm_factor`` = 1 / prod(stride) for the
    transposed case (:23-27), applied inside the square root, then + 1e-6.
    """
    sq = weight * weight
    if transposed:
        s = sq.sum(dim=(0, 2, 3)).view(1, -1, 1, 1)
        c = 1.0
        for t in stride:
            c = c / t
    else:
        s = sq.sum(dim=(1, 2, 3)).view(1, -1, 1, 1)
        c = 1.0
    return (s * c + _EPS).sqrt()


def affine_epilogue(z, norm, scale, bias):
    """``z / norm [* scale] [+ bias]`` — WeightNormalizedConv.py:40-49, ...Linear.py:33-39."""
    y = z / norm
    if scale is not None:
        y = y * scale
    if bias is not None:
        y = y + bias
    return y


class _WNConvBase(nn.Module):
    """Parameter set of ``_WeightNormalizedConvNd.__init__`` (WeightNormalizedConv.py:11-27)."""

    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation,
                 output_padding, scale, bias, init_factor, init_scale):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.output_padding = _pair(output_padding)
        kh, kw = self.kernel_size
        if self.transposed:
            shape = (in_channels, out_channels, kh, kw)
        else:
            shape = (out_channels, in_channels, kh, kw)
        # 2017 `_ConvNd.reset_parameters`: U(+-1/sqrt(in_channels * prod(k))) for both
        # directions, then `weight.data.mul_(init_factor)` (:22).
        bound = 1.0 / math.sqrt(in_channels * kh * kw)
        self.weight = nn.Parameter(torch.empty(shape).uniform_(-bound, bound) * init_factor)
        if scale:
            self.scale = nn.Parameter(torch.full((1, out_channels, 1, 1), float(init_scale)))
        else:
            self.register_parameter("scale", None)
        if bias:
            self.bias = nn.Parameter(torch.zeros(1, out_channels, 1, 1))
        else:
            self.register_parameter("bias", None)
        self.weight_norm_factor = 1.0
        if self.transposed:
            for t in self.stride:
                self.weight_norm_factor = self.weight_norm_factor / t

    def weight_norm(self):
        return conv_weight_norm(self.weight, self.transposed, self.stride)

    def norm_scale_bias(self, z):
        return affine_epilogue(z, self.weight_norm(), self.scale, self.bias)


class WeightNormalizedConv2d(_WNConvBase):
    """WeightNormalizedConv.py:67-81."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 scale=True, bias=True, init_factor=1, init_scale=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, 0,
                         scale, bias, init_factor, init_scale)

    def forward(self, x):
        z = F.conv2d(x, self.weight, None, self.stride, self.padding, self.dilation, 1)
        return self.norm_scale_bias(z)


class WeightNormalizedConvTranspose2d(_WNConvBase):
    """WeightNormalizedConv.py:83-99 (``output_size`` resolves to an output padding)."""

    transposed = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 output_padding=0, scale=True, bias=True, dilation=1, init_factor=1, init_scale=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation,
                         output_padding, scale, bias, init_factor, init_scale)

    def _resolve_output_padding(self, x, output_size):
        if output_size is None:
            return self.output_padding
        output_size = list(output_size)[-2:]
        res = []
        for d in range(2):
            lo = ((x.size(d + 2) - 1) * self.stride[d] - 2 * self.padding[d]
                  + self.dilation[d] * (self.kernel_size[d] - 1) + 1)
            extra = output_size[d] - lo
            if extra < 0 or extra >= max(self.stride[d], self.dilation[d]):
                raise ValueError("requested output size is not reachable")
            res.append(extra)
        return tuple(res)

    def forward(self, x, output_size=None):
        op = self._resolve_output_padding(x, output_size)
        z = F.conv_transpose2d(x, self.weight, None, self.stride, self.padding, op, 1, self.dilation)
        return self.norm_scale_bias(z)


class WeightNormalizedLinear(nn.Module):
    """WeightNormalizedLinear.py:7-42: rows of W normalised, optional (1,out) scale/bias."""

    def __init__(self, in_features, out_features, scale=True, bias=True, init_factor=1, init_scale=1):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        bound = 1.0 * init_factor / math.sqrt(in_features)  # reset_parameters, :24-28
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-bound, bound))
        if bias:
            self.bias = nn.Parameter(torch.empty(1, out_features).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)
        if scale:
            self.scale = nn.Parameter(torch.full((1, out_features), float(init_scale)))
        else:
            self.register_parameter("scale", None)

    def weight_norm(self):
        # legacy keep-dim: (out, 1)
        return ((self.weight * self.weight).sum(dim=1, keepdim=True) + _EPS).sqrt()

    def norm_scale_bias(self, z):
        return affine_epilogue(z, self.weight_norm().transpose(0, 1), self.scale, self.bias)

    def forward(self, x):
        return self.norm_scale_bias(F.linear(x, self.weight))


def tprelu(x, a_raw, b):
    """Translated PReLU, TPReLU.py:16-18: ``prelu(x - b, clamp(a, 0, 1)) + b`` along dim 1."""
    shape = (1, -1) + (1,) * (x.dim() - 2)
    a = a_raw.clamp(0, 1).view(shape)
    bb = b.view(shape)
    t = x - bb
    return torch.where(t > 0, t, a * t) + bb


class TPReLU(nn.Module):
    """TPReLU.py:8-21; parameters ``weight`` (slope, init .25) and ``bias`` (translation, init 0)."""

    def __init__(self, num_parameters=1, init=0.25):
        super().__init__()
        self.num_parameters = num_parameters
        self.weight = nn.Parameter(torch.full((num_parameters,), float(init)))
        self.bias = nn.Parameter(torch.zeros(num_parameters))

    def forward(self, x):
        return tprelu(x, self.weight, self.bias)


class View(nn.Module):
    """View.py:4-11."""

    def __init__(self, *target_size):
        super().__init__()
        self.target_size = target_size

    def forward(self, x):
        return x.contiguous().view(x.size(0), *self.target_size)
