"""Autograd-visible operators of the G-LIS path, backed by the sm_100a kernels.

Each ``torch.autograd.Function`` here replaces a group of stock-PyTorch calls the
reference makes from its modules (file:line cited per class).  Activations cross this
boundary as fp32 CUDA tensors; 4-D tensors are handled in NHWC memory order
(``torch.channels_last``) and returned that way, which leaves their logical
``(N, C, H, W)`` shape — what the reference API promises — untouched.
"""
import ctypes as C

import torch

from . import _lib as L


def _nhwc(x):
    """fp32, CUDA, NHWC-dense view/copy of a 2-D (B,F) or 4-D (N,C,H,W) tensor."""
    if not x.is_cuda:
        raise RuntimeError("glis_b200: CUDA tensor required (no CPU fallback)")
    if x.dtype != torch.float32:
        raise RuntimeError("glis_b200: fp32 tensors required, got %s" % x.dtype)
    if x.dim() == 4:
        return x.contiguous(memory_format=torch.channels_last)
    if x.dim() == 2:
        return x.contiguous()
    raise RuntimeError("glis_b200: expected a 2-D or 4-D tensor, got %d-D" % x.dim())


def _empty_nhwc(n, c, h, w, like):
    return torch.empty((n, c, h, w), device=like.device, dtype=torch.float32,
                       memory_format=torch.channels_last)


class ContractionSpec(object):
    """Static description of one WN layer's contraction (conv / transposed conv / linear)."""

    def __init__(self, transposed, kernel_size, stride, padding, dilation, output_padding=(0, 0),
                 linear=False, precision=None):
        self.transposed, self.linear = bool(transposed), bool(linear)
        self.kernel_size, self.stride, self.padding = tuple(kernel_size), tuple(stride), tuple(padding)
        self.dilation, self.output_padding = tuple(dilation), tuple(output_padding)
        self._precision = precision

    @property
    def precision(self):
        return L.default_precision if self._precision is None else self._precision

    @property
    def norm_factor(self):
        c = 1.0
        if self.transposed:
            for s in self.stride:
                c /= s
        return c

    def out_hw(self, h, w):
        res = []
        for d, size in enumerate((h, w)):
            k, s, p, dl = self.kernel_size[d], self.stride[d], self.padding[d], self.dilation[d]
            if self.transposed:
                res.append((size - 1) * s - 2 * p + dl * (k - 1) + self.output_padding[d] + 1)
            else:
                res.append((size + 2 * p - dl * (k - 1) - 1) // s + 1)
        return tuple(res)

    def geom(self, relation, n, hi, wi, ci, ho, wo, co):
        g = L.Geom()
        g.relation = relation
        g.N, g.Hi, g.Wi, g.Ci, g.Ho, g.Wo, g.Co = n, hi, wi, ci, ho, wo, co
        g.KH, g.KW = self.kernel_size
        g.stride_h, g.stride_w = self.stride
        g.pad_h, g.pad_w = self.padding
        g.dil_h, g.dil_w = self.dilation
        return g


def wn_prepare(weight, scale, spec, want_io=True, want_oi=True):
    """norm [Cout], pack_io [T][Cin][Cout], pack_oi [T][Cout][Cin] of the effective weights."""
    out_axis = 1 if spec.transposed else 0
    cout = weight.shape[out_axis]
    cin = weight.shape[1 - out_axis]
    t = weight.numel() // (cout * cin)
    w = weight.detach().contiguous()
    norm = torch.empty(cout, device=w.device, dtype=torch.float32)
    io = torch.empty(t, cin, cout, device=w.device, dtype=torch.float32) if want_io else None
    oi = torch.empty(t, cout, cin, device=w.device, dtype=torch.float32) if want_oi else None
    sc = None if scale is None else scale.detach().contiguous()
    L.call("glis_wn_prepare", L.ptr(w), L.ptr(sc), out_axis, cout, cin, t, spec.norm_factor,
           L.ptr(norm), L.ptr(io), L.ptr(oi), L.stream(), kernels=2 if (want_io or want_oi) else 1)
    return norm, io, oi


def wn_prepare_bf16(weight, scale, spec, want_fwd=True, want_bwd=True, lo=True):
    """norm [Cout] and K-major bf16 (hi, lo) packs: fwd [T][Cout][Cin], bwd [T][Cin][Cout]."""
    out_axis = 1 if spec.transposed else 0
    cout, cin = weight.shape[out_axis], weight.shape[1 - out_axis]
    t = weight.numel() // (cout * cin)
    w = weight.detach().contiguous()
    dev = w.device
    norm = torch.empty(cout, device=dev, dtype=torch.float32)

    def plane(shape, want):
        return torch.empty(shape, device=dev, dtype=torch.bfloat16) if want else None

    fh, fl = plane((t, cout, cin), want_fwd), plane((t, cout, cin), want_fwd and lo)
    bh, bl = plane((t, cin, cout), want_bwd), plane((t, cin, cout), want_bwd and lo)
    sc = None if scale is None else scale.detach().contiguous()
    L.call("glis_wn_prepare_bf16", L.ptr(w), L.ptr(sc), out_axis, cout, cin, t, spec.norm_factor, L.ptr(norm),
           L.ptr16(fh), L.ptr16(fl), L.ptr16(bh), L.ptr16(bl), L.stream(),
           kernels=2 if (want_fwd or want_bwd) else 1)
    return norm, (fh, fl), (bh, bl)


def split_bf16(x, lo=True):
    """(hi, lo) bf16 planes of a dense fp32 tensor, same memory order."""
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo_t = torch.empty_like(x, dtype=torch.bfloat16) if lo else None
    L.call("glis_split_bf16", L.ptr(x), L.ptr16(hi), L.ptr16(lo_t), x.numel(), L.stream())
    return hi, lo_t


def tc_supported(g):
    return bool(L.load().glis_conv_tc_supported(C.byref(g)))


def _launch_geom(spec, relation, x_nhwc, out_shape_nchw):
    if x_nhwc.dim() == 4:
        n, ci, hi, wi = x_nhwc.shape
        _, co, ho, wo = out_shape_nchw
        out = _empty_nhwc(n, co, ho, wo, x_nhwc)
    else:
        n, ci = x_nhwc.shape
        hi = wi = ho = wo = 1
        co = out_shape_nchw[1]
        out = torch.empty((n, co), device=x_nhwc.device, dtype=torch.float32)
    return spec.geom(relation, n, hi, wi, ci, ho, wo, co), out


def _tag(relation, g):
    if relation == L.CONV:
        return "conv_forward M=%d N=%d K=%d" % (g.N * g.Ho * g.Wo, g.Co, g.Ci * g.KH * g.KW)
    return "tconv_forward M=%d N=%d K=%d" % (g.N * g.Hi * g.Wi, g.Co * g.KH * g.KW, g.Ci)


def conv_forward(spec, relation, x_nhwc, wpack, out_shape_nchw, bias=None, act=L.ACT_NONE,
                 act_a=None, act_b=None, preact=None):
    """fp32 FFMA gather-GEMM; ``x_nhwc`` and the result are NHWC-dense (or 2-D)."""
    g, out = _launch_geom(spec, relation, x_nhwc, out_shape_nchw)
    ep = L.Epilogue(L.ptr(bias), act, L.ptr(act_a), L.ptr(act_b), L.ptr(preact))
    with L.timed(_tag(relation, g) + " fp32"):
        L.call("glis_conv_forward", C.byref(g), L.ptr(x_nhwc), L.ptr(wpack), C.byref(ep), L.ptr(out),
               L.PREC_FP32, L.stream())
    return out


def conv_forward_tc(spec, relation, x_planes, wpack_planes, out_shape_nchw, like, precision, bias=None,
                    act=L.ACT_NONE, act_a=None, act_b=None, preact=None):
    """tcgen05 implicit GEMM on split-bf16 planes; returns the fp32 NHWC result."""
    g, out = _launch_geom(spec, relation, like, out_shape_nchw)
    ep = L.Epilogue(L.ptr(bias), act, L.ptr(act_a), L.ptr(act_b), L.ptr(preact))
    with L.timed(_tag(relation, g) + " tc"):
        L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(x_planes[0]), L.ptr16(x_planes[1]),
               L.ptr16(wpack_planes[0]), L.ptr16(wpack_planes[1]), C.byref(ep), L.ptr(out), None, None,
               precision, L.stream())
    return out


def _use_tc(spec, relation, in_shape, out_shape):
    """Whether the launch reading NCHW-shaped ``in_shape`` and writing ``out_shape`` runs on tcgen05."""
    if spec.precision == L.PREC_FP32 or len(in_shape) != 4:
        return False  # batch-sized linears stay on the FFMA kernel
    n, ci, hi, wi = in_shape
    _, co, ho, wo = out_shape
    return tc_supported(spec.geom(relation, n, hi, wi, ci, ho, wo, co))


class WNContraction(torch.autograd.Function):
    """``norm_scale_bias(F.conv2d / F.conv_transpose2d / F.linear (x, w))``.

    Reference: common/modules/WeightNormalizedConv.py:79-81, :96-99, :29-49 and
    common/modules/WeightNormalizedLinear.py:30-42.  The per-channel ``scale/norm`` is
    folded into the packed weights (one pass over the parameters), the bias into the GEMM
    epilogue; backward applies the closed-form projection of SURVEY.md App. E.
    """

    @staticmethod
    def forward(ctx, x, weight, scale, bias, spec):
        xc = _nhwc(x)
        need_dx = ctx.needs_input_grad[0]
        out_axis = 1 if spec.transposed else 0
        cout = weight.shape[out_axis]
        if xc.dim() == 4:
            n, ci, h, w = xc.shape
            ho, wo = spec.out_hw(h, w)
            shape = (n, cout, ho, wo)
        else:
            shape = (xc.shape[0], cout)
            ci = xc.shape[1]
        if ci != weight.shape[1 - out_axis]:
            raise RuntimeError("glis_b200: input has %d channels, weight expects %d"
                               % (ci, weight.shape[1 - out_axis]))
        b = None if bias is None else bias.detach().reshape(-1).contiguous()
        rel_f = L.TCONV if spec.transposed else L.CONV
        rel_b = L.CONV if spec.transposed else L.TCONV
        prec = spec.precision
        tc_f = _use_tc(spec, rel_f, tuple(xc.shape), shape)
        tc_b = need_dx and _use_tc(spec, rel_b, shape, tuple(xc.shape))
        lo = prec == L.PREC_BF16X3
        pack_oi = bwd_planes = None
        if tc_f or tc_b:
            norm, fwd_planes, bwd_planes = wn_prepare_bf16(weight, scale, spec, tc_f, tc_b, lo)
        if not tc_f or (need_dx and not tc_b):
            norm, pack_io, pack_oi = wn_prepare(weight, scale, spec, not tc_f, need_dx and not tc_b)
        x_planes = (None, None)
        if tc_f:
            x_planes = split_bf16(xc, lo)
            out = conv_forward_tc(spec, rel_f, x_planes, fwd_planes, shape, xc, prec, bias=b)
        else:
            out = conv_forward(spec, rel_f, xc, pack_io, shape, bias=b)
        ctx.spec, ctx.tc_b, ctx.prec = spec, tc_b, prec
        ctx.has_scale, ctx.has_bias = scale is not None, bias is not None
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        saved_b = bwd_planes if tc_b else (None, None)
        ctx.save_for_backward(xc, weight, scale, norm, pack_oi, saved_b[0], saved_b[1], x_planes[0], x_planes[1])
        return out

    @staticmethod
    def backward(ctx, dy):
        xc, weight, scale, norm, pack_oi, bwd_hi, bwd_lo, x_hi, x_lo = ctx.saved_tensors
        spec = ctx.spec
        dyc = _nhwc(dy)
        out_axis = 1 if spec.transposed else 0
        cout, cin = weight.shape[out_axis], weight.shape[1 - out_axis]
        t = weight.numel() // (cout * cin)
        if xc.dim() == 4:
            n, _, h, w = xc.shape
            ho, wo = dyc.shape[2], dyc.shape[3]
        else:
            n, h, w, ho, wo = xc.shape[0], 1, 1, 1, 1

        dx = None
        want_lo = ctx.prec == L.PREC_BF16X3
        dy_planes = None
        if ctx.needs_input_grad[0]:
            # conv layer: dx gathers dy through the transposed relation; transposed layer: the direct one
            rel = L.CONV if spec.transposed else L.TCONV
            if ctx.tc_b:
                dy_planes = split_bf16(dyc, want_lo)
                dx = conv_forward_tc(spec, rel, dy_planes, (bwd_hi, bwd_lo), tuple(xc.shape), dyc, ctx.prec)
            else:
                dx = conv_forward(spec, rel, dyc, pack_oi, tuple(xc.shape))

        dw = dscale = dbias = None
        if ctx.needs_input_grad[1] or (ctx.has_scale and ctx.needs_input_grad[2]):
            graw = torch.zeros_like(weight, memory_format=torch.contiguous_format)
            if spec.transposed:   # small = x (Cin), big = dy (Cout)
                g = spec.geom(L.CONV, n, ho, wo, cout, h, w, cin)
                small, big = xc, dyc
            else:                 # small = dy (Cout), big = x (Cin)
                g = spec.geom(L.CONV, n, h, w, cin, ho, wo, cout)
                small, big = dyc, xc
            tag = "conv_wgrad M=%d N=%d K=%d" % (g.Co, g.Ci * t, n * g.Ho * g.Wo)
            if ctx.prec != L.PREC_FP32 and xc.dim() == 4 and L.load().glis_wgrad_tc_supported(C.byref(g)):
                if dy_planes is None:
                    dy_planes = split_bf16(dyc, want_lo)
                xp = (x_hi, x_lo) if x_hi is not None else split_bf16(xc, want_lo)
                sp, bp = (xp, dy_planes) if spec.transposed else (dy_planes, xp)
                with L.timed(tag + " tc"):
                    L.call("glis_conv_wgrad_bf16", C.byref(g), L.ptr16(sp[0]), L.ptr16(sp[1]), L.ptr16(bp[0]),
                           L.ptr16(bp[1]), L.ptr(graw), ctx.prec, L.stream())
            else:
                with L.timed(tag + " fp32"):
                    L.call("glis_conv_wgrad", C.byref(g), L.ptr(small), L.ptr(big), L.ptr(graw), L.PREC_FP32,
                           L.stream())
            dw = torch.empty_like(graw)
            if ctx.has_scale:
                dscale = torch.empty(cout, device=dw.device, dtype=torch.float32)
            wc = weight.detach().contiguous()
            sc = None if scale is None else scale.detach().contiguous()
            L.call("glis_wn_project", L.ptr(graw), L.ptr(wc), L.ptr(sc), L.ptr(norm), out_axis, cout, cin, t,
                   spec.norm_factor, L.ptr(dw), L.ptr(dscale), 0, L.stream())
            if dscale is not None:
                dscale = dscale.view_as(scale)
        if ctx.has_bias and ctx.needs_input_grad[3]:
            dbias = torch.empty(cout, device=dyc.device, dtype=torch.float32)
            L.call("glis_channel_sum", L.ptr(dyc), L.ptr(dbias), dyc.numel(), cout, 1, 0, L.stream())
            dbias = dbias.view(ctx.bias_shape)
        return dx, dw, dscale, dbias, None


def wn_contraction(x, weight, scale, bias, spec):
    out = WNContraction.apply(x, weight, scale, bias, spec)
    return out


def _channel_layout(x):
    """(dense tensor, inner) such that channel(i) = (i // inner) % C over its storage order."""
    if x.dim() == 2:
        return x.contiguous(), 1
    if x.dim() == 4:
        if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
            return x, 1
        xc = x.contiguous()
        return xc, xc.shape[2] * xc.shape[3]
    raise RuntimeError("glis_b200: TPReLU expects a 2-D or 4-D tensor")


class TPReLUFunction(torch.autograd.Function):
    """``F.prelu(x - b, a.clamp(0, 1)) + b`` — common/modules/TPReLU.py:16-18."""

    @staticmethod
    def forward(ctx, x, a_raw, b):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("glis_b200: fp32 CUDA tensor required (no CPU fallback)")
        xc, inner = _channel_layout(x)
        c = a_raw.numel()
        if xc.shape[1] != c:
            raise RuntimeError("glis_b200: TPReLU has %d channels, input has %d" % (c, xc.shape[1]))
        out = torch.empty_like(xc)
        L.call("glis_tprelu_forward", L.ptr(xc), L.ptr(a_raw.detach()), L.ptr(b.detach()), L.ptr(out),
               xc.numel(), c, inner, L.stream())
        ctx.inner = inner
        ctx.save_for_backward(xc, a_raw, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, a_raw, b = ctx.saved_tensors
        c = a_raw.numel()
        if xc.dim() == 4 and ctx.inner == 1:
            dc = dout.contiguous(memory_format=torch.channels_last)
        else:
            dc = dout.contiguous()
        dx = torch.empty_like(xc)
        da = torch.zeros(c, device=xc.device, dtype=torch.float32)
        db = torch.zeros(c, device=xc.device, dtype=torch.float32)
        L.call("glis_tprelu_backward", L.ptr(xc), L.ptr(a_raw.detach()), L.ptr(b.detach()), L.ptr(dc),
               L.ptr(dx), L.ptr(da), L.ptr(db), xc.numel(), c, ctx.inner, L.stream())
        return dx, da, db


def tprelu(x, a_raw, b):
    return TPReLUFunction.apply(x, a_raw, b)


def rmsprop_(p_flat, g_flat, v_flat, lr, alpha=0.9, eps=1e-6, gscale=1.0):
    """Fused RMSprop over flat parameter / gradient / square-average buffers (g_lis/main.py:313-314)."""
    L.call("glis_rmsprop", L.ptr(p_flat), L.ptr(g_flat), L.ptr(v_flat), p_flat.numel(), lr, alpha, eps, gscale,
           L.stream())


def randn_(out, seed, offset=0):
    L.call("glis_randn", L.ptr(out), out.numel(), seed, offset, L.stream())
    return out


def uniform_(out, seed, offset=0):
    L.call("glis_uniform", L.ptr(out), out.numel(), seed, offset, L.stream())
    return out


def bce_logits(logit, target, gscale=1.0, want_grad=True, want_prob=False):
    """(loss[1], dlogit or None, prob or None) for mean BCE of sigmoid(logit) against a constant target."""
    lg = logit.detach().reshape(-1).contiguous()
    loss = torch.empty(1, device=lg.device, dtype=torch.float32)
    dl = torch.empty_like(lg) if want_grad else None
    pr = torch.empty_like(lg) if want_prob else None
    L.call("glis_bce_logits", L.ptr(lg), float(target), lg.numel(), float(gscale), L.ptr(loss), L.ptr(dl),
           L.ptr(pr), L.stream())
    return loss, dl, pr


def mse_scaled(u, z, lam, du=None, accumulate=False):
    uc, zc = u.detach().contiguous(), z.detach().contiguous()
    loss = torch.empty(1, device=uc.device, dtype=torch.float32)
    L.call("glis_mse_scaled", L.ptr(uc), L.ptr(zc), uc.numel(), float(lam), L.ptr(loss), L.ptr(du),
           1 if accumulate else 0, L.stream())
    return loss
