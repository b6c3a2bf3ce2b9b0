#!/usr/bin/env python
"""What would single-pass reduced-precision BACKWARD contractions cost in gradient accuracy?  (CPU, oracle only.)

VERDICT r1 item 4 asks for `kind::tf32` data / weight gradients (the north star allows 1e-3 on gradients) "or justify
why not".  This study answers with numbers before any kernel is written: the fp64 oracle runs whole G-LIS iterations
twice — exactly, and with every backward contraction (data gradient of conv / transposed conv / linear and their
weight gradients) fed operands rounded to the candidate format, accumulating exactly (what a tensor core with fp32
accumulators does to 1e-7) — and the gradients of all parameters are compared in the max-norm-relative measure the
parity tests use.  The forward pass and the TPReLU masks are identical in both runs, so what is measured is the
arithmetic alone.

Formats: tf32rn (operands rounded to nearest at 10 mantissa bits — needs the producers to pre-round),
tf32tr (truncated: what `tcgen05.mma.kind::tf32` does to raw fp32 operands), bf16 (single pass), bf16x2 (hi*hi + lo*hi +
hi*lo without ... i.e. the current three-product scheme for reference), bf16x3.

usage: python tools/tf32_gradient_study.py [W nf nl B seeds]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import oracle.modules as om  # noqa: E402
from oracle.step import GLISOracleTrainer  # noqa: E402

MODE = ["exact"]


def _round_bits(x, keep, nearest=True):
    """fp64 -> fp32 -> `keep` explicit mantissa bits (RN-even or truncate) -> fp64."""
    f = x.to(torch.float32).contiguous()
    i = f.view(torch.int32)
    drop = 23 - keep
    if nearest:
        bias = ((i >> drop) & 1) + ((1 << (drop - 1)) - 1)
        i = i + bias
    i = i & ~((1 << drop) - 1)
    return i.view(torch.float32).to(torch.float64)


def _terms(x, mode):
    """The list of (operand-a part, operand-b part) products a mode issues is built by the caller; this returns the
    parts of one operand: [hi] or [hi, lo]."""
    if mode == "tf32rn":
        return [_round_bits(x, 10, True)]
    if mode == "tf32tr":
        return [_round_bits(x, 10, False)]
    if mode == "bf16":
        return [_round_bits(x, 7, True)]
    if mode in ("bf16x3", "bf16x2"):
        hi = _round_bits(x, 7, True)
        lo = _round_bits(x - hi, 7, True)
        return [hi, lo]
    raise ValueError(mode)


def _pairs(a, b, mode):
    A, B = _terms(a, mode), _terms(b, mode)
    if mode == "bf16x3":
        return [(A[0], B[0]), (A[0], B[1]), (A[1], B[0])]
    if mode == "bf16x2":     # drop the hi*lo product of the SECOND operand: it stays single-plane
        return [(A[0], B[0]), (A[1], B[0])]
    return [(A[0], B[0])]


class _Contraction(torch.autograd.Function):
    """y = op(x, w) exactly; backward contractions on rounded operands when MODE != exact."""

    @staticmethod
    def forward(ctx, x, w, kind, args):
        ctx.kind, ctx.args = kind, args
        ctx.save_for_backward(x, w)
        return _Contraction.apply_op(kind, x, w, args)

    @staticmethod
    def apply_op(kind, x, w, args):
        if kind == "conv":
            return _ORIG["conv2d"](x, w, None, *args)
        if kind == "tconv":
            return _ORIG["conv_transpose2d"](x, w, None, *args)
        return _ORIG["linear"](x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        kind, args = ctx.kind, ctx.args
        mode = MODE[0]

        def grads(xx, ww, gg, need_x, need_w):
            with torch.enable_grad():
                xx = xx.detach().requires_grad_(need_x)
                ww = ww.detach().requires_grad_(need_w)
                y = _Contraction.apply_op(kind, xx, ww, args)
                return torch.autograd.grad(y, [t for t, n in ((xx, need_x), (ww, need_w)) if n], gg)

        if mode == "exact":
            gx, gw = grads(x, w, gy, True, True)
            return gx, gw, None, None
        gx = sum(grads(x, wp, gp, True, False)[0] for gp, wp in _pairs(gy, w, mode))
        gw = sum(grads(xp, w, gp, False, True)[0] for gp, xp in _pairs(gy, x, mode))
        return gx, gw, None, None


_ORIG = {"conv2d": F.conv2d, "conv_transpose2d": F.conv_transpose2d, "linear": F.linear}


class _PatchedF(object):
    def __getattr__(self, name):
        return getattr(F, name)

    @staticmethod
    def conv2d(x, w, b, stride, padding, dilation, groups):
        return _Contraction.apply(x, w, "conv", (stride, padding, dilation, groups))

    @staticmethod
    def conv_transpose2d(x, w, b, stride, padding, output_padding, groups, dilation):
        return _Contraction.apply(x, w, "tconv", (stride, padding, output_padding, groups, dilation))

    @staticmethod
    def linear(x, w, b=None):
        assert b is None
        return _Contraction.apply(x, w, "linear", ())


def rel(a, b):
    d = b.abs().max().item()
    return (a - b).abs().max().item() / d if d > 0 else 0.0


def main():
    a = [int(v) for v in sys.argv[1:]]
    W, nf, nl, B, seeds = (a + [32, 32, 3, 16, 3][len(a):])[:5]
    code = 64
    om.F = _PatchedF()
    import oracle.model as omodel
    if hasattr(omodel, "F"):
        omodel.F = om.F
    modes = ["bf16x3", "tf32rn", "tf32tr", "bf16x2", "bf16"]
    worst = {m: (0.0, "") for m in modes}
    med = {m: [] for m in modes}
    for seed in range(seeds):
        ref = None
        for mode in ["exact"] + modes:
            torch.manual_seed(seed)
            g = oracle.GeneratorLearnedInputSpace(W, W, nf, nl, code, "weight", 1, "fractional").double()
            d = oracle.build_discriminator(W, W, nf, nl, "weight", 0).double()
            tr = GLISOracleTrainer(g, d, lr=2e-5)
            real, zd, zg = torch.rand(B, 3, W, W).double(), torch.randn(B, code).double(), torch.randn(B, code).double()
            MODE[0] = mode
            grads = {}
            # D's gradients are consumed by its update inside the iteration: capture them with hooks
            hooks = []
            for net, tag in ((g, "G."), (d, "D.")):
                for name, p in net.named_parameters():
                    hooks.append(p.register_hook(lambda gr, k=tag + name: grads.__setitem__(
                        k, grads[k] + gr.detach().clone() if k in grads else gr.detach().clone())))
            tr.step(real, zd, zg, 1, 1)
            for h in hooks:
                h.remove()
            if mode == "exact":
                ref = grads
                continue
            errs = [(rel(grads[k], ref[k]), k) for k in ref if ref[k].abs().max().item() > 0]
            e, k = max(errs)
            if e > worst[mode][0]:
                worst[mode] = (e, k)
            med[mode].append(sorted(x for x, _ in errs)[len(errs) // 2])
    print("config: %dx%d nfeature %d levels %d batch %d, %d seeds; max-norm relative error of parameter gradients"
          % (W, W, nf, nl, B, seeds))
    for m in modes:
        print("%-7s worst %.2e (%s)   median over parameters %.2e" % (m, worst[m][0], worst[m][1],
                                                                      sum(med[m]) / len(med[m])))


if __name__ == "__main__":
    main()
