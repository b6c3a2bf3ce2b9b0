"""Network builders of the G-LIS path — drop-ins for the reference's common/model.py.

Same call signatures, parameter shapes and ``state_dict`` keys as the reference
(SURVEY.md App. C): ``build_discriminator`` (:10-63), ``build_generator`` (:65-140),
``GeneratorLearnedInputSpace`` (:142-314) and ``build_reverser`` (:316-368).  The layers
are the sm_100a-backed modules of ``common.modules``.  ``norm`` must be ``'weight'`` or
``'weight-affine'`` — the batch-norm / un-normalized variants of the reference are plain
torch modules and not part of this path.
"""
import random

import torch.nn as nn

from glis_b200 import ops
from glis_b200.naming import DottedSequential as _DottedSequential
from .modules import (TPReLU, View, WeightNormalizedConv2d, WeightNormalizedConvTranspose2d,
                      WeightNormalizedLinear)
from .modules.WeightNormalizedConv import _WeightNormalizedConvNd

__all__ = ["build_discriminator", "build_generator", "GeneratorLearnedInputSpace", "build_reverser"]


def _planes_suffice(consumer, follower, out_shape):
    """Does the module that consumes a fused pair's output read it through its bf16 planes only?  (Then
    the pair does not write the fp32 copy of that activation: ops.consumes_planes_only.)"""
    if not isinstance(consumer, _WeightNormalizedConvNd) or len(out_shape) != 4:
        return False
    return ops.consumes_planes_only(consumer._spec(), consumer.weight, out_shape, isinstance(follower, TPReLU))


def run_layers(layers, x, following=(), out=None):
    """Apply ``layers`` in order, running every (weight-normalized layer, TPReLU) pair as ONE
    operator: the TPReLU moves into the contraction's epilogue (``ops.wn_contraction_tprelu``).
    Module structure, parameters and results are those of calling the modules one by one.
    ``following``: the modules the caller applies to the result next (only looked at, not run).
    ``out``: caller-owned buffer for the result of a trailing (WN layer, Sigmoid) pair (no_grad only)."""
    layers = list(layers)
    n_run = len(layers)
    layers = layers + list(following)
    i = 0
    while i < n_run:
        m = layers[i]
        nxt = layers[i + 1] if i + 1 < len(layers) else None
        fusable = (isinstance(m, (_WeightNormalizedConvNd, WeightNormalizedLinear)) and isinstance(nxt, TPReLU)
                   and x.is_cuda and x.dtype == ops.torch.float32 and x.dim() in (2, 4))
        after = layers[i + 2] if i + 2 < len(layers) else None
        # linear -> View(C, h, w) -> TPReLU(C): the generators' head, one kernel writing the map in NHWC
        head = (isinstance(m, WeightNormalizedLinear) and isinstance(nxt, View) and isinstance(after, TPReLU)
                and m.bias is None and x.is_cuda and x.dtype == ops.torch.float32 and x.dim() == 2
                and len(nxt.target_size) == 3 and after.weight.numel() == nxt.target_size[0]
                and m.out_features == nxt.target_size[0] * nxt.target_size[1] * nxt.target_size[2])
        if head and i + 3 <= n_run:
            consumer = layers[i + 3] if i + 3 < len(layers) else None
            follower = layers[i + 4] if i + 4 < len(layers) else None
            suffice = _planes_suffice(consumer, follower, (x.shape[0],) + tuple(nxt.target_size))
            x = ops.wn_linear_view_tprelu(x, m.weight, m.scale, after.weight, after.bias, nxt.target_size, suffice)
            i += 3
        elif fusable and i + 2 <= n_run:
            spec = m._spec() if isinstance(m, _WeightNormalizedConvNd) else m._spec
            consumer = layers[i + 2] if i + 2 < len(layers) else None
            follower = layers[i + 3] if i + 3 < len(layers) else None
            suffice = x.dim() == 4 and _planes_suffice(consumer, follower, ops.layer_out_shape(x.shape, m.weight, spec))
            x = ops.wn_contraction_tprelu(x, m.weight, m.scale, m.bias, nxt.weight, nxt.bias, spec, suffice)
            i += 2
        elif (isinstance(m, _WeightNormalizedConvNd) and type(nxt) is nn.Sigmoid and i + 2 <= n_run
              and x.is_cuda and x.dtype == ops.torch.float32 and x.dim() == 4):
            # the generators' last two modules: the sigmoid moves into the contraction's epilogue
            x = ops.wn_contraction_sigmoid(x, m.weight, m.scale, m.bias, m._spec(),
                                           out=out if i + 2 == n_run else None)
            i += 2
        else:
            x = m(x)
            i += 1
    return x


def lis_residual(block, x):
    """``x + block(x)`` for one LIS block (model.py:281-297); one cluster kernel when the block is the
    norm='weight' form ``WN linear -> TPReLU -> WN linear`` (no scale / bias) and the code size fits."""
    mods = list(block.children())
    if (len(mods) == 3 and isinstance(mods[0], WeightNormalizedLinear) and isinstance(mods[1], TPReLU)
            and isinstance(mods[2], WeightNormalizedLinear) and x.is_cuda and x.dtype == ops.torch.float32
            and x.dim() == 2 and all(m.scale is None and m.bias is None for m in (mods[0], mods[2]))
            and mods[0].in_features == mods[0].out_features == mods[2].in_features == mods[2].out_features
            == x.shape[1] == mods[1].weight.numel() and ops.lis_supported(x.shape[1])):
        return ops.lis_module(x, mods[0].weight, mods[2].weight, mods[1].weight, mods[1].bias,
                              mods[0]._spec, mods[2]._spec)
    return x + block(x)


class DottedSequential(_DottedSequential):
    """Sequential container with reference-compatible dotted child names whose forward fuses
    (WN layer, TPReLU) pairs."""

    def forward(self, input, out=None):
        return run_layers(self._modules.values(), input, out=out)


class PhiloxDropout(nn.Dropout):
    """``nn.Dropout`` (the reference's module at model.py:52-53) whose training-mode mask comes from the
    counter-based generator of ``glis_dropout``: capturable into the iteration's CUDA graph (a replay draws a
    fresh mask), nothing stored for backward, reproducible on the host from (seed, counter, call)."""

    def forward(self, input):
        if not self.training or self.p == 0 or not input.is_cuda:
            return super(PhiloxDropout, self).forward(input)
        return ops.dropout(input, self.p, channel_mode=False)


class PhiloxDropout2d(nn.Dropout2d):
    """``nn.Dropout2d`` (model.py:344-346) on the same generator: one mask element per (image, channel)."""

    def forward(self, input):
        if not self.training or self.p == 0 or not input.is_cuda:
            return super(PhiloxDropout2d, self).forward(input)
        return ops.dropout(input, self.p, channel_mode=True)


def _require_even(w, h, what):
    if (w % 2 != 0) or (h % 2 != 0):
        raise ValueError("%s width and height must be even numbers" % what)


def _is_affine(norm):
    if norm == "weight":
        return False
    if norm == "weight-affine":
        return True
    raise NotImplementedError("glis_b200 builds the weight-normalized networks only "
                              "(norm='weight' or 'weight-affine'), got %r" % (norm,))


def _activation(norm, channels):
    """('tprelu', TPReLU) under norm='weight', ('prelu', nn.PReLU) under 'weight-affine'."""
    if norm == "weight":
        return "tprelu", TPReLU(channels)
    return "prelu", nn.PReLU(channels)


def _extra_pad(size, is_last):
    """One extra pixel of padding where halving would otherwise give an odd size (model.py:19-30)."""
    return 1 if (not is_last and size % 4 == 2) else 0


def _encoder_levels(net, w, h, f_first, num_levels, norm, dropout2d=0):
    """The stride-2 4x4 stack shared by D (:14-50) and R (:320-357). Returns (channels, w, h)."""
    affine = _is_affine(norm)
    f_prev, f = 3, f_first
    for level in range(num_levels):
        last = level == num_levels - 1
        pw, ph = _extra_pad(w, last), _extra_pad(h, last)
        net.add_module("level.{0}.conv".format(level),
                       WeightNormalizedConv2d(f_prev, f, 4, 2, (1 + ph, 1 + pw), scale=affine, bias=affine))
        if level >= 1 and dropout2d > 0:
            net.add_module("level.{0}.sd".format(level), PhiloxDropout2d(dropout2d))
        kind, act = _activation(norm, f)
        net.add_module("level.{0}.{1}".format(level, kind), act)
        f_prev, f = f, 2 * f
        w, h = (w + 2 * pw) // 2, (h + 2 * ph) // 2
    return f_prev, w, h


def build_discriminator(w_in, h_in, f_first, num_down_layers, norm, p_dropout=0):
    _require_even(w_in, h_in, "input")
    net = DottedSequential()
    f_prev, w, h = _encoder_levels(net, w_in, h_in, f_first, num_down_layers, norm)
    if p_dropout > 0:
        net.add_module("final.dropout", PhiloxDropout(p_dropout))
    net.add_module("final.conv", WeightNormalizedConv2d(f_prev, 1, (h, w)))
    net.add_module("final.sigmoid", nn.Sigmoid())
    net.add_module("final.view", View(1))
    return net


def build_reverser(w_in, h_in, f_first, num_down_layers, code_size, norm, spatial_dropout_r=0):
    _require_even(w_in, h_in, "input")
    net = DottedSequential()
    f_prev, w, h = _encoder_levels(net, w_in, h_in, f_first, num_down_layers, norm,
                                   dropout2d=spatial_dropout_r)
    net.add_module("final.conv", WeightNormalizedConv2d(f_prev, code_size, (h, w)))
    net.add_module("final.view", View(code_size))
    return net


class DecoderPlan(object):
    """Geometry of the up-sampling stack (model.py:69-91, :149-170): the (w, h, f) of the
    tensor produced by the initial linear layer and each level's extra padding."""

    def __init__(self, w_out, h_out, f_last, num_up_layers):
        _require_even(w_out, h_out, "output")
        self.pad_w, self.pad_h = [], []
        w, h, f = w_out, h_out, f_last
        for _ in range(num_up_layers - 1):
            pw, ph = (1 if w % 4 == 2 else 0), (1 if h % 4 == 2 else 0)
            self.pad_w.append(pw)
            self.pad_h.append(ph)
            w, h, f = (w + 2 * pw) // 2, (h + 2 * ph) // 2, 2 * f
        self.pad_w.append(0)
        self.pad_h.append(0)
        self.w, self.h, self.f = w // 2, h // 2, f
        self.levels = num_up_layers

    def padding(self, level):
        return (1 + self.pad_h[level], 1 + self.pad_w[level])


def build_generator(w_out, h_out, f_last, num_up_layers, code_size, norm):
    plan = DecoderPlan(w_out, h_out, f_last, num_up_layers)
    affine = _is_affine(norm)
    net = DottedSequential()
    f = plan.f
    net.add_module("initial.linear",
                   WeightNormalizedLinear(code_size, f * plan.h * plan.w, init_factor=0.01,
                                          scale=affine, bias=affine))
    net.add_module("initial.view", View(f, plan.h, plan.w))
    kind, act = _activation(norm, f)
    net.add_module("initial." + kind, act)
    for level in range(num_up_layers - 1, 0, -1):
        net.add_module("level.{0}.conv".format(level),
                       WeightNormalizedConvTranspose2d(f, f // 2, 4, 2, plan.padding(level),
                                                       scale=affine, bias=affine))
        kind, act = _activation(norm, f // 2)
        net.add_module("level.{0}.{1}".format(level, kind), act)
        f //= 2
    net.add_module("level.0.conv", WeightNormalizedConvTranspose2d(f, 3, 4, 2, plan.padding(0)))
    net.add_module("level.0.sigmoid", nn.Sigmoid())
    return net


class GeneratorLearnedInputSpace(nn.Module):
    """Generator preceded by ``n_lis_layers`` residual LIS modules with stochastic depth.

    ``forward(x, n_execute_lis_layers=None)`` returns ``(image, [lis outputs])`` exactly as
    model.py:275-314.  ``self.rng`` (default: the global ``random`` module, as in the
    reference) supplies the depth draws; data-parallel ranks share its seed.
    """

    def __init__(self, w_out, h_out, f_last, num_up_layers, code_size, norm, n_lis_layers, upscaling):
        super(GeneratorLearnedInputSpace, self).__init__()
        plan = DecoderPlan(w_out, h_out, f_last, num_up_layers)
        affine = _is_affine(norm)
        self.w, self.h, self.f = plan.w, plan.h, plan.f
        self.code_size = code_size
        self.rng = random

        def lis_linear():
            return WeightNormalizedLinear(code_size, code_size, init_factor=0.01, scale=affine, bias=affine)

        blocks = []
        for i in range(n_lis_layers):
            block = DottedSequential()
            block.add_module("lis.{0}-1.linear".format(i), lis_linear())
            block.add_module("lis.{0}-1.act".format(i), _activation(norm, code_size)[1])
            block.add_module("lis.{0}-2.linear".format(i), lis_linear())
            blocks.append(block)
        self.lis_layers = nn.ModuleList(blocks)

        f = plan.f
        self.initial_linear = nn.ModuleList([
            WeightNormalizedLinear(code_size, f * plan.h * plan.w, init_factor=0.01, scale=affine, bias=affine),
            View(f, plan.h, plan.w),
            _activation(norm, f)[1],
        ])

        layers = []
        for level in range(num_up_layers - 1, 0, -1):
            if upscaling == "fractional":
                layers.append(WeightNormalizedConvTranspose2d(f, f // 2, 4, 2, plan.padding(level),
                                                              scale=affine, bias=affine))
            elif upscaling in ("nearest", "bilinear"):
                layers.append(nn.UpsamplingNearest2d(scale_factor=2) if upscaling == "nearest"
                              else nn.UpsamplingBilinear2d(scale_factor=2))
                layers.append(WeightNormalizedConv2d(f, f // 2, 3, 1, plan.padding(level),
                                                     scale=affine, bias=affine))
            else:
                raise Exception("Unknown upscaling, must be fractional|nearest|bilinear, got %s" % (upscaling,))
            layers.append(_activation(norm, f // 2)[1])
            f //= 2
        layers.append(WeightNormalizedConvTranspose2d(f, 3, 4, 2, plan.padding(0)))
        layers.append(nn.Sigmoid())
        self.conv_layers = nn.ModuleList(layers)

    def lis_depth(self, n_execute_lis_layers=None):
        """Number of LIS modules the next forward executes (break rule of model.py:281-297);
        consumes one ``rng.random()`` per module visited, like the reference."""
        n = len(self.lis_layers)
        for i in range(n):
            p = (1.0 / 2) ** (n - i) if self.training else 0
            if n_execute_lis_layers is not None:
                run = n_execute_lis_layers == "all" or (i + 1) <= n_execute_lis_layers
                p = 0 if run else 1
            if self.rng.random() < p:
                return i
        return n

    def forward(self, x, n_execute_lis_layers=None, out=None):
        """``out`` (additive, no_grad only): NHWC buffer the generated images are written into."""
        lis_results = []
        for i in range(self.lis_depth(n_execute_lis_layers)):
            x = lis_residual(self.lis_layers[i], x)
            lis_results.append(x)
        x = run_layers(self.initial_linear, x, following=self.conv_layers)
        x = run_layers(self.conv_layers, x, out=out)
        return x, lis_results
