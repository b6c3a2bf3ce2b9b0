"""ORACLE (test infrastructure) — network topologies of the G-LIS path, restated.

Follows ``common/model.py`` of the reference: discriminator (:10-63), plain generator
(:65-140), generator with learned input space (:142-314) and reverser (:316-368).
Only ``norm='weight'`` / ``'weight-affine'`` are restated.

Module names in the reference contain dots (``'level.0.conv'``); modern torch rejects
them, so :class:`DottedSequential` registers children under a mangled name and
translates ``state_dict`` keys so that they read exactly as in SURVEY App. C.
"""
import random as _random

import torch
import torch.nn as nn

from .modules import (TPReLU, View, WeightNormalizedConv2d,
                      WeightNormalizedConvTranspose2d, WeightNormalizedLinear)

_DOT = "·"  # stands in for '.' inside registered child names


class DottedSequential(nn.Sequential):
    """``nn.Sequential`` that accepts child names containing dots."""

    def __init__(self):
        super().__init__()
        self._register_state_dict_hook(DottedSequential._export_keys)
        self._register_load_state_dict_pre_hook(self._import_keys)

    def add_module(self, name, module):
        super().add_module(name.replace(".", _DOT), module)

    def child(self, name):
        return getattr(self, name.replace(".", _DOT))

    @staticmethod
    def _export_keys(module, state, prefix, _meta):
        for k in [k for k in state if k.startswith(prefix) and _DOT in k[len(prefix):]]:
            state[prefix + k[len(prefix):].replace(_DOT, ".")] = state.pop(k)

    def _import_keys(self, state, prefix, *_):
        for name in self._modules:
            if _DOT not in name:
                continue
            dotted = prefix + name.replace(_DOT, ".") + "."
            for k in [k for k in state if k.startswith(dotted)]:
                state[prefix + name + "." + k[len(dotted):]] = state.pop(k)


def _check_even(w, h, what):
    if w % 2 or h % 2:
        raise ValueError("%s width and height must be even numbers" % what)


def _affine(norm):
    if norm not in ("weight", "weight-affine"):
        raise NotImplementedError("oracle restates norm='weight' / 'weight-affine' only")
    return norm == "weight-affine"


def _act(norm, channels):
    return ("tprelu", TPReLU(channels)) if norm == "weight" else ("prelu", nn.PReLU(channels))


def _down_stack(net, w, h, f_first, levels, norm, channel_dropout=0.0):
    """Shared body of D and R: model.py:14-50 and :320-357.

    Non-final levels pad one extra pixel on an axis whose size is 2 mod 4 so that the
    halved size stays even; the last level never does.  Returns (f_prev, w, h).
    """
    aff = _affine(norm)
    f_prev, f = 3, f_first
    for i in range(levels):
        last = i == levels - 1
        pw = 1 if (not last and w % 4 == 2) else 0
        ph = 1 if (not last and h % 4 == 2) else 0
        net.add_module("level.%d.conv" % i,
                       WeightNormalizedConv2d(f_prev, f, 4, 2, (1 + ph, 1 + pw), scale=aff, bias=aff))
        if channel_dropout > 0 and i >= 1:  # reverser only, between conv and activation (:344-346)
            net.add_module("level.%d.sd" % i, nn.Dropout2d(channel_dropout))
        kind, act = _act(norm, f)
        net.add_module("level.%d.%s" % (i, kind), act)
        f_prev, f = f, f * 2
        w, h = (w + 2 * pw) // 2, (h + 2 * ph) // 2
    return f_prev, w, h


def build_discriminator(w_in, h_in, f_first, num_down_layers, norm, p_dropout=0):
    """model.py:10-63.  ``p_dropout`` defaults to 0 (5-arg call sites)."""
    _check_even(w_in, h_in, "input")
    net = DottedSequential()
    f_prev, w, h = _down_stack(net, w_in, h_in, f_first, num_down_layers, norm)
    if p_dropout > 0:
        net.add_module("final.dropout", nn.Dropout(p_dropout))
    net.add_module("final.conv", WeightNormalizedConv2d(f_prev, 1, (h, w)))
    net.add_module("final.sigmoid", nn.Sigmoid())
    net.add_module("final.view", View(1))
    return net


def build_reverser(w_in, h_in, f_first, num_down_layers, code_size, norm, spatial_dropout_r=0):
    """model.py:316-368: D's body, a ``code_size``-channel head and no sigmoid."""
    _check_even(w_in, h_in, "input")
    net = DottedSequential()
    f_prev, w, h = _down_stack(net, w_in, h_in, f_first, num_down_layers, norm,
                               channel_dropout=spatial_dropout_r)
    net.add_module("final.conv", WeightNormalizedConv2d(f_prev, code_size, (h, w)))
    net.add_module("final.view", View(code_size))
    return net


def up_plan(w_out, h_out, f_last, num_up_layers):
    """Mirror of the down-stack geometry: model.py:69-91 / :149-170.

    Returns (w0, h0, f0, pad_w, pad_h): the spatial size and width of the tensor the
    initial linear produces and the extra padding of each level (index = level).
    """
    pad_w, pad_h = [], []
    w, h, f = w_out, h_out, f_last
    for _ in range(num_up_layers - 1):
        if w % 4 == 2:
            pad_w.append(1)
            w = (w + 2) // 2
        else:
            pad_w.append(0)
            w //= 2
        if h % 4 == 2:
            pad_h.append(1)
            h = (h + 2) // 2
        else:
            pad_h.append(0)
            h //= 2
        f *= 2
    pad_w.append(0)
    pad_h.append(0)
    return w // 2, h // 2, f, pad_w, pad_h


def build_generator(w_out, h_out, f_last, num_up_layers, code_size, norm):
    """model.py:65-140 — the LIS-free generator used by the R-iterative trainer."""
    _check_even(w_out, h_out, "output")
    aff = _affine(norm)
    w, h, f, pad_w, pad_h = up_plan(w_out, h_out, f_last, num_up_layers)
    net = DottedSequential()
    net.add_module("initial.linear",
                   WeightNormalizedLinear(code_size, f * h * w, init_factor=0.01, scale=aff, bias=aff))
    net.add_module("initial.view", View(f, h, w))
    kind, act = _act(norm, f)
    net.add_module("initial.%s" % kind, act)
    for level in range(num_up_layers - 1, 0, -1):
        net.add_module("level.%d.conv" % level,
                       WeightNormalizedConvTranspose2d(f, f // 2, 4, 2,
                                                       (1 + pad_h[level], 1 + pad_w[level]),
                                                       scale=aff, bias=aff))
        kind, act = _act(norm, f // 2)
        net.add_module("level.%d.%s" % (level, kind), act)
        f //= 2
    net.add_module("level.0.conv",
                   WeightNormalizedConvTranspose2d(f, 3, 4, 2, (1 + pad_h[0], 1 + pad_w[0])))
    net.add_module("level.0.sigmoid", nn.Sigmoid())
    return net


class GeneratorLearnedInputSpace(nn.Module):
    """model.py:142-314.  ``rng`` (default: Python's global ``random``, as in :294) is
    injectable so that data-parallel ranks and parity tests can share depth decisions."""

    def __init__(self, w_out, h_out, f_last, num_up_layers, code_size, norm, n_lis_layers,
                 upscaling="fractional", rng=None):
        super().__init__()
        _check_even(w_out, h_out, "output")
        aff = _affine(norm)
        w, h, f, pad_w, pad_h = up_plan(w_out, h_out, f_last, num_up_layers)
        self.w, self.h, self.f = w, h, f
        self.rng = rng if rng is not None else _random

        lis = []
        for i in range(n_lis_layers):  # :176-201
            seq = DottedSequential()
            seq.add_module("lis.%d-1.linear" % i,
                           WeightNormalizedLinear(code_size, code_size, init_factor=0.01, scale=aff, bias=aff))
            seq.add_module("lis.%d-1.act" % i,
                           TPReLU(code_size) if norm == "weight" else nn.PReLU(code_size))
            seq.add_module("lis.%d-2.linear" % i,
                           WeightNormalizedLinear(code_size, code_size, init_factor=0.01, scale=aff, bias=aff))
            lis.append(seq)
        self.lis_layers = nn.ModuleList(lis)

        self.initial_linear = nn.ModuleList([  # :203-219
            WeightNormalizedLinear(code_size, f * h * w, init_factor=0.01, scale=aff, bias=aff),
            View(f, h, w),
            _act(norm, f)[1],
        ])

        conv = []
        for level in range(num_up_layers - 1, 0, -1):  # :222-260
            pad = (1 + pad_h[level], 1 + pad_w[level])
            if upscaling == "fractional":
                conv.append(WeightNormalizedConvTranspose2d(f, f // 2, 4, 2, pad, scale=aff, bias=aff))
            elif upscaling == "nearest":
                conv.append(nn.UpsamplingNearest2d(scale_factor=2))
                conv.append(WeightNormalizedConv2d(f, f // 2, 3, 1, pad, scale=aff, bias=aff))
            elif upscaling == "bilinear":
                conv.append(nn.UpsamplingBilinear2d(scale_factor=2))
                conv.append(WeightNormalizedConv2d(f, f // 2, 3, 1, pad, scale=aff, bias=aff))
            else:
                raise Exception("Unknown upscaling, must be fractional|nearest|bilinear, got %s" % (upscaling,))
            conv.append(_act(norm, f // 2)[1])
            f //= 2
        conv.append(WeightNormalizedConvTranspose2d(f, 3, 4, 2, (1 + pad_h[0], 1 + pad_w[0])))  # :262-267
        conv.append(nn.Sigmoid())
        self.conv_layers = nn.ModuleList(conv)

    def lis_depth(self, n_execute_lis_layers=None):
        """How many LIS modules this forward runs — the break rule of model.py:281-297.

        One ``rng.random()`` draw per module visited, stopping at the first break.
        """
        n = len(self.lis_layers)
        for i in range(n):
            p = 0.5 ** (n - i) if self.training else 0
            if n_execute_lis_layers is not None:
                p = 0 if (n_execute_lis_layers == "all" or (i + 1) <= n_execute_lis_layers) else 1
            if self.rng.random() < p:
                return i
        return n

    def forward(self, x, n_execute_lis_layers=None):
        depth = self.lis_depth(n_execute_lis_layers)
        lis_results = []
        for i in range(depth):
            x = x + self.lis_layers[i](x)  # :303
            lis_results.append(x)
        for layer in self.initial_linear:
            x = layer(x)
        for layer in self.conv_layers:
            x = layer(x)
        return x, lis_results
