"""Autograd-visible operators of the G-LIS path, backed by the sm_100a kernels.

Each ``torch.autograd.Function`` here replaces a group of stock-PyTorch calls the
reference makes from its modules (file:line cited per class).  Activations cross this
boundary as fp32 CUDA tensors; 4-D tensors are handled in NHWC memory order
(``torch.channels_last``) and returned that way, which leaves their logical
``(N, C, H, W)`` shape — what the reference API promises — untouched.
"""
import ctypes as C
import os

import torch

from . import _lib as L


def _nhwc(x):
    """fp32, CUDA, NHWC-dense view/copy of a 2-D (B,F) or 4-D (N,C,H,W) tensor."""
    if not x.is_cuda:
        raise RuntimeError("glis_b200: CUDA tensor required (no CPU fallback)")
    if x.dtype != torch.float32:
        raise RuntimeError("glis_b200: fp32 tensors required, got %s" % x.dtype)
    if x.dim() == 4:
        xc = x.contiguous(memory_format=torch.channels_last)
    elif x.dim() == 2:
        xc = x.contiguous()
    else:
        raise RuntimeError("glis_b200: expected a 2-D or 4-D tensor, got %d-D" % x.dim())
    if xc is not x:
        _require_f32(x, "a layout conversion")
    return xc


def _require_f32(t, what):
    """Inside a fused chain an activation may exist ONLY as bf16 planes (its fp32 buffer is allocated but
    never written: ``launch(..., want_f32=False)``); anything that needs the fp32 values must not get one."""
    if getattr(t, "_glis_f32_invalid", False):
        raise RuntimeError("glis_b200: %s needs the fp32 values of a planes-only activation "
                           "(internal planning error)" % what)


def _empty_nhwc(n, c, h, w, like):
    return torch.empty((n, c, h, w), device=like.device, dtype=torch.float32,
                       memory_format=torch.channels_last)


class ContractionSpec(object):
    """Static description of one WN layer's contraction (conv / transposed conv / linear)."""

    def __init__(self, transposed, kernel_size, stride, padding, dilation, output_padding=(0, 0),
                 linear=False, precision=None, perm=None):
        self.transposed, self.linear = bool(transposed), bool(linear)
        self.kernel_size, self.stride, self.padding = tuple(kernel_size), tuple(stride), tuple(padding)
        self.dilation, self.output_padding = tuple(dilation), tuple(output_padding)
        self._precision = precision
        # (C, P): a linear layer whose Cout = C*P output features are produced in NHWC order (row p*C + c of
        # every pack = master row c*P + p), i.e. already laid out as the (C, h, w) map a following View makes
        self.perm = None if perm is None else (int(perm[0]), int(perm[1]))

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


# ---------------------------------------------------------------------------- packed weights
PARAM_EPOCH = [0]   # bumped whenever parameters change outside torch's version counters


def bump_param_epoch():
    """Invalidate cached weight packs (called after the fused RMSprop kernel and after every
    CUDA-graph replay, which update parameters without touching ``Tensor._version``)."""
    PARAM_EPOCH[0] += 1


class PackedWeights(object):
    """Everything derived from one WN layer's (weight, scale) by ``glis_wn_prepare*``: the
    per-channel norm and the fp32 / split-bf16 GEMM-order packs.  One persistent object (and one
    set of buffers) per weight Parameter: a kind is (re)built in place when it is asked for and
    not ``fresh`` — lazily at first use, or ahead of time by ``refresh_packs`` right after the
    optimizer step.  ``packed_weights`` clears ``fresh`` whenever the parameters changed."""

    def __init__(self, weight, scale, spec):
        self.weight, self.scale, self.spec = weight, scale, spec
        self.out_axis = 1 if spec.transposed else 0
        self.cout, self.cin = weight.shape[self.out_axis], weight.shape[1 - self.out_axis]
        self.t = weight.numel() // (self.cout * self.cin)
        self.norm = None
        self.io = self.oi = None            # fp32 [T][Cin][Cout], [T][Cout][Cin]
        self.fwd = self.bwd = None          # bf16 (hi, lo): [T][Cout][Cin], [T][Cin][Cout]
        self.mat = self.mat_t = None        # bf16 (hi, lo): E [A][J] and E^T [J][A] (image-side layers)
        self.fresh = set()                  # kinds valid for the current parameter values
        self.pending = None                 # (event, kinds) of a rebuild in flight on the side stream

    def invalidate(self):
        self.fresh.clear()

    def _w(self):
        w = self.weight.detach().contiguous()
        sc = None if self.scale is None else self.scale.detach().contiguous()
        return w, sc

    def wanted(self):
        wanted = getattr(self.weight, "_glis_wanted", None)
        if wanted is None:
            wanted = set()
            try:
                self.weight._glis_wanted = wanted
            except AttributeError:
                pass
        return wanted

    def _wait(self, kinds):
        """Kinds being rebuilt on the side stream (refresh_packs): the first use waits for that stream."""
        if self.pending is not None and (self.pending[1] & set(kinds)):
            torch.cuda.current_stream().wait_event(self.pending[0])
            self.pending = None

    def need_fp32(self, io, oi):
        # a layer that needs one fp32 pack this step needs the other one in its backward: build both
        # on the first request (one launch pair instead of two) once backward has been seen to want it
        wanted = self.wanted()
        if io:
            wanted.add("io")
        if oi:
            wanted.add("oi")
        self._wait(("norm",) + (("io",) if io else ()) + (("oi",) if oi else ()))
        io = (io or "io" in wanted) and "io" not in self.fresh
        oi = (oi or "oi" in wanted) and "oi" not in self.fresh
        if not (io or oi) and "norm" in self.fresh:
            return
        self._build_fp32(io, oi)

    def _build_fp32(self, io, oi):
        w, sc = self._w()
        dev = w.device
        if self.norm is None:
            self.norm = torch.empty(self.cout, device=dev, dtype=torch.float32)
        if io and self.io is None:
            self.io = torch.empty(self.t, self.cin, self.cout, device=dev, dtype=torch.float32)
        if oi and self.oi is None:
            self.oi = torch.empty(self.t, self.cout, self.cin, device=dev, dtype=torch.float32)
        pc, pp = self.spec.perm or (0, 0)
        L.call("glis_wn_prepare_perm", L.ptr(w), L.ptr(sc), self.out_axis, self.cout, self.cin, self.t,
               self.spec.norm_factor, L.ptr(self.norm), L.ptr(self.io if io else None),
               L.ptr(self.oi if oi else None), pc, pp, L.stream(), kernels=2 if (io or oi) else 1)
        self.fresh.add("norm")
        if io:
            self.fresh.add("io")
        if oi:
            self.fresh.add("oi")

    def need_bf16(self, fwd, bwd, lo):
        wanted = self.wanted()
        if fwd:
            wanted.add("fwd")
        if bwd:
            wanted.add("bwd")
        self._wait((("fwd",) if fwd else ()) + (("bwd",) if bwd else ()))
        fwd = (fwd or "fwd" in wanted) and "fwd" not in self.fresh
        bwd = (bwd or "bwd" in wanted) and "bwd" not in self.fresh
        if not (fwd or bwd):
            return
        self._build_bf16(fwd, bwd, lo)

    def _build_bf16(self, fwd, bwd, lo):
        w, sc = self._w()
        dev = w.device
        if self.norm is None:
            self.norm = torch.empty(self.cout, device=dev, dtype=torch.float32)

        def planes(old, shape):
            hi = old[0] if old is not None else torch.empty(shape, device=dev, dtype=torch.bfloat16)
            lo_t = old[1] if old is not None else None
            if lo and lo_t is None:
                lo_t = torch.empty(shape, device=dev, dtype=torch.bfloat16)
            return hi, lo_t

        if fwd:
            self.fwd = planes(self.fwd, (self.t, self.cout, self.cin))
        if bwd:
            self.bwd = planes(self.bwd, (self.t, self.cin, self.cout))
        fh, fl = self.fwd if fwd else (None, None)
        bh, bl = self.bwd if bwd else (None, None)
        pc, pp = self.spec.perm or (0, 0)
        L.call("glis_wn_prepare_bf16_perm", L.ptr(w), L.ptr(sc), self.out_axis, self.cout, self.cin, self.t,
               self.spec.norm_factor, L.ptr(self.norm), L.ptr16(fh), L.ptr16(fl if lo else None), L.ptr16(bh),
               L.ptr16(bl if lo else None), pc, pp, L.stream(), kernels=2)
        self.fresh.add("norm")
        if fwd:
            self.fresh.add("fwd")
        if bwd:
            self.fresh.add("bwd")

    FORWARD_KINDS, BACKWARD_KINDS = ("fwd", "io", "mat", "norm"), ("bwd", "oi")

    def refresh(self, part="all"):
        """Rebuild, in place, the kinds this layer has ever been asked for and that are stale:
        ``part`` = "forward" (what forward launches read), "backward" (what data-gradient launches
        read) or "all"."""
        wanted = self.wanted()
        lo = self.spec.precision == L.PREC_BF16X3
        f, b = part in ("all", "forward"), part in ("all", "backward")
        fwd = f and "fwd" in wanted and "fwd" not in self.fresh
        bwd = b and "bwd" in wanted and "bwd" not in self.fresh
        if fwd or bwd:
            self._build_bf16(fwd, bwd, lo)
        io = f and "io" in wanted and "io" not in self.fresh
        oi = b and "oi" in wanted and "oi" not in self.fresh
        if io or oi or (f and "norm" not in self.fresh):
            self._build_fp32(io, oi)
        if f and "mat" in wanted:
            _need_matrix(self, lo)


def _dp(t):
    return None if t is None else t.data_ptr()


def _describe(pw, part):
    """glis_wn_layer_t for the kinds ``PackedWeights.refresh(part)`` would rebuild (buffers allocated, kinds marked
    fresh), or None when there is nothing to do."""
    wanted = pw.wanted()
    lo = pw.spec.precision == L.PREC_BF16X3
    f, b = part in ("all", "forward"), part in ("all", "backward")
    fwd = f and "fwd" in wanted and "fwd" not in pw.fresh
    bwd = b and "bwd" in wanted and "bwd" not in pw.fresh
    io = f and "io" in wanted and "io" not in pw.fresh
    oi = b and "oi" in wanted and "oi" not in pw.fresh
    mat = f and "mat" in wanted and "mat" not in pw.fresh
    norm = "norm" not in pw.fresh and (f or fwd or bwd or io or oi or mat)
    if not (fwd or bwd or io or oi or mat or norm):
        return None
    w, sc = pw._w()
    dev = w.device
    mk16 = lambda shape: torch.empty(shape, device=dev, dtype=torch.bfloat16)
    if pw.norm is None:
        pw.norm = torch.empty(pw.cout, device=dev, dtype=torch.float32)
    if io and pw.io is None:
        pw.io = torch.empty(pw.t, pw.cin, pw.cout, device=dev, dtype=torch.float32)
    if oi and pw.oi is None:
        pw.oi = torch.empty(pw.t, pw.cout, pw.cin, device=dev, dtype=torch.float32)
    for kind, want, shape in (("fwd", fwd, (pw.t, pw.cout, pw.cin)), ("bwd", bwd, (pw.t, pw.cin, pw.cout))):
        if not want:
            continue
        cur = getattr(pw, kind)
        hi = cur[0] if cur is not None else mk16(shape)
        lo_t = cur[1] if cur is not None else None
        if lo and lo_t is None:
            lo_t = mk16(shape)
        setattr(pw, kind, (hi, lo_t))
    a = w.shape[0]
    j = w.numel() // a
    if mat:
        if pw.mat is None:
            pw.mat, pw.mat_t = (mk16((a, j)), mk16((a, j)) if lo else None), (mk16((j, a)), mk16((j, a)) if lo else None)
        elif lo and pw.mat[1] is None:
            pw.mat, pw.mat_t = (pw.mat[0], torch.empty_like(pw.mat[0])), (pw.mat_t[0], torch.empty_like(pw.mat_t[0]))
    d = L.WnLayer()
    d.w, d.scale, d.norm = w.data_ptr(), _dp(sc), pw.norm.data_ptr()
    d.pack_io, d.pack_oi = _dp(pw.io) if io else None, _dp(pw.oi) if oi else None
    if fwd:
        d.fwd_hi, d.fwd_lo = pw.fwd[0].data_ptr(), _dp(pw.fwd[1]) if lo else None
    if bwd:
        d.bwd_hi, d.bwd_lo = pw.bwd[0].data_ptr(), _dp(pw.bwd[1]) if lo else None
    if mat:
        d.mat_hi, d.mat_lo = pw.mat[0].data_ptr(), _dp(pw.mat[1]) if lo else None
        d.matt_hi, d.matt_lo = pw.mat_t[0].data_ptr(), _dp(pw.mat_t[1]) if lo else None
        d.mat_rows = a
    d.out_axis, d.Cout, d.Cin, d.T = pw.out_axis, pw.cout, pw.cin, pw.t
    d.perm_c, d.perm_p = pw.spec.perm or (0, 0)
    d.need_norm, d.c = int(norm), pw.spec.norm_factor
    pw.fresh.add("norm")
    for kind, want in (("fwd", fwd), ("bwd", bwd), ("io", io), ("oi", oi), ("mat", mat)):
        if want:
            pw.fresh.add(kind)
    d._keep = (w, sc)               # the contiguous copies the pointers refer to stay alive until the call
    return d


MULTI_PREPARE = os.environ.get("GLIS_MULTI_PREPARE", "1") != "0"


def refresh_many(pws, part):
    """``pw.refresh(part)`` for every pack set in ``pws`` as ONE glis_wn_prepare_multi call (two launches: all
    norms, then all packs) on the current stream."""
    if not MULTI_PREPARE:
        for pw in pws:
            pw.refresh(part)
        return
    descs = [d for d in (_describe(pw, part) for pw in pws) if d is not None]
    if not descs:
        return
    arr = (L.WnLayer * len(descs))(*descs)
    L.call("glis_wn_prepare_multi", arr, len(descs), L.stream(), kernels=2 * ((len(descs) + 23) // 24))


def _need_matrix(pw, lo):
    """E / E^T packs of an image-side layer (csrc/image_side.cu): one norm + one pack launch."""
    pw.wanted().add("mat")
    pw._wait(("mat", "norm"))
    if "mat" in pw.fresh:
        return
    w, sc = pw._w()
    dev = w.device
    if "norm" not in pw.fresh:
        pw.need_fp32(False, False)          # the norm alone
    a = w.shape[0]
    j = w.numel() // a
    if pw.mat is None:
        mk = lambda shape, want: torch.empty(shape, device=dev, dtype=torch.bfloat16) if want else None
        pw.mat, pw.mat_t = (mk((a, j), True), mk((a, j), lo)), (mk((j, a), True), mk((j, a), lo))
    elif lo and pw.mat[1] is None:
        pw.mat = (pw.mat[0], torch.empty_like(pw.mat[0]))
        pw.mat_t = (pw.mat_t[0], torch.empty_like(pw.mat_t[0]))
    (eh, el), (th, tl) = pw.mat, pw.mat_t
    L.call("glis_wn_pack_matrix_bf16", L.ptr(w), L.ptr(sc), L.ptr(pw.norm), pw.out_axis, a, j, pw.t,
           L.ptr16(eh), L.ptr16(el if lo else None), L.ptr16(th), L.ptr16(tl if lo else None), L.stream())
    pw.fresh.add("mat")


def _pack_key(weight, scale, spec):
    return (PARAM_EPOCH[0], getattr(weight, "_glis_epoch", 0), weight.data_ptr(), weight._version,
            None if scale is None else scale._version, spec.precision, spec.transposed, spec.stride, spec.perm)


def packed_weights(weight, scale, spec):
    key = _pack_key(weight, scale, spec)
    cached = getattr(weight, "_glis_packed", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    if cached is not None and cached[0][2] == key[2] and cached[0][5:] == key[5:] and cached[1].scale is scale:
        pw = cached[1]          # same storage and layer structure, new parameter values: rebuild in place
        pw.invalidate()
    else:
        pw = PackedWeights(weight, scale, spec)
    try:
        weight._glis_packed = (key, pw)
    except AttributeError:
        pass
    return pw


def refresh_packs(flat, part="all", side=True):
    """Rebuild the stale weight packs of every layer of ``flat`` (a trainer.FlatParams) for its CURRENT
    parameter values.  ``part``: "forward" / "backward" / "all" kinds (PackedWeights.refresh).  With
    ``side`` the rebuild runs on the side stream and the first later use of one of those kinds waits
    for it, so it overlaps whatever the main stream does next: the discriminator's packs rebuild
    under the generator's forward, the generator's data-gradient packs under the whole D update."""
    todo = []
    for p in flat.params:
        cached = getattr(p, "_glis_packed", None)
        if cached is None:
            continue
        pw = cached[1]
        key = _pack_key(p, pw.scale, pw.spec)
        if cached[0] != key:
            if cached[0][5] != key[5]:
                continue        # precision switched since: rebuilt lazily as a new object
            pw.invalidate()
            p._glis_packed = (key, pw)
        todo.append(pw)
    if not todo:
        return
    use_side = side and Overlap.enabled and torch.cuda.is_available()
    if not use_side:
        all_kinds = PackedWeights.FORWARD_KINDS + PackedWeights.BACKWARD_KINDS
        if MULTI_PREPARE or not (Overlap.enabled and torch.cuda.is_available() and len(todo) > 2):
            for pw in todo:
                pw._wait(all_kinds)
            refresh_many(todo, part)
            return
        # Needed by the very next kernel (G's forward packs at the end of an iteration): the layers are
        # independent and each is a pair of short launches, so build them on three streams at once — the
        # heaviest layer (G's initial linear, 3.3 M weights) bounds the wait instead of the sum of all
        main = torch.cuda.current_stream()
        lanes = [main] + Overlap.aux_streams(2)
        load = [0] * len(lanes)
        plan = [[] for _ in lanes]
        for pw in sorted(todo, key=lambda q: -q.weight.numel()):
            k = load.index(min(load))
            plan[k].append(pw)
            load[k] += pw.weight.numel() + 200000      # + a launch pair's fixed cost
        ready = torch.cuda.Event()
        ready.record(main)
        for s, pws in zip(lanes[1:], plan[1:]):
            s.wait_event(ready)
            with torch.cuda.stream(s):
                for pw in pws:
                    pw._wait(all_kinds)
                    pw.refresh(part)
        for pw in plan[0]:
            pw._wait(all_kinds)
            pw.refresh(part)
        for s in lanes[1:]:
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)
        return
    kinds = set(PackedWeights.FORWARD_KINDS if part in ("all", "forward") else ()) | \
        set(PackedWeights.BACKWARD_KINDS if part in ("all", "backward") else ()) | {"norm"}
    s = Overlap.stream()
    main = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(main)
    s.wait_event(ready)
    launches = L.launch_count
    with torch.cuda.stream(s):
        refresh_many(todo, part)
        if L.launch_count == launches:
            return              # everything was fresh: nothing enqueued, nothing to wait for
        done = torch.cuda.Event()
        done.record(s)
    for pw in todo:
        pw.pending = (done, kinds)


def forget_pending(flat):
    """Drop the 'rebuild in flight on the side stream' marks of ``flat``'s packs — after the caller has joined the
    side stream itself (their events were recorded inside a CUDA-graph capture, or the join made them moot)."""
    for p in flat.params:
        cached = getattr(p, "_glis_packed", None)
        if cached is not None:
            cached[1].pending = None


def wn_prepare(weight, scale, spec, want_io=True, want_oi=True):
    """(norm [Cout], pack_io [T][Cin][Cout], pack_oi [T][Cout][Cin]) of the effective weights (uncached)."""
    pw = PackedWeights(weight, scale, spec)
    pw.need_fp32(want_io, want_oi)
    return pw.norm, pw.io, pw.oi


# ---------------------------------------------------------------------------- split-bf16 planes
def split_bf16(x, lo=True):
    """(hi, lo) bf16 planes of a dense fp32 tensor, same memory order."""
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo_t = torch.empty_like(x, dtype=torch.bfloat16) if lo else None
    L.call("glis_split_bf16", L.ptr(x), L.ptr16(hi), L.ptr16(lo_t), x.numel(), L.stream())
    return hi, lo_t


def attach_planes(t, planes):
    """Remember the hi/lo planes a producer kernel already wrote for ``t`` (consumed by the next
    tensor-core layer instead of re-splitting)."""
    if planes is not None and planes[0] is not None:
        t._glis_planes = (planes[0], planes[1], t.data_ptr(), t._version)
    return t


def planes_of(t, lo=True):
    """Planes of ``t``: the attached ones if still valid, else a fresh split."""
    if isinstance(t, PlanesOnly):
        return t._glis_planes_only
    tag = getattr(t, "_glis_planes", None)
    if tag is not None and tag[2] == t.data_ptr() and tag[3] == t._version and (tag[1] is not None or not lo):
        return tag[0], tag[1]
    _require_f32(t, "a bf16 split")
    return split_bf16(t, lo)


def tc_supported(g):
    return bool(L.load().glis_conv_tc_supported(C.byref(g)))


def _usable_out(buf, shape, like):
    """``buf`` if the caller's output buffer can take an NHWC result of NCHW ``shape`` on ``like``'s device."""
    if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != torch.float32 or buf.device != like.device:
        return None
    if not buf.is_contiguous(memory_format=torch.channels_last if buf.dim() == 4 else torch.contiguous_format):
        return None
    return buf


def _launch_geom(spec, relation, in_shape, out_shape_nchw, like, out=None):
    if len(in_shape) == 4:
        n, ci, hi, wi = in_shape
        _, co, ho, wo = out_shape_nchw
        out = _usable_out(out, (n, co, ho, wo), like)
        if out is None:
            out = _empty_nhwc(n, co, ho, wo, like)
    else:
        n, ci = in_shape
        hi = wi = ho = wo = 1
        co = out_shape_nchw[1]
        out = torch.empty((n, co), device=like.device, dtype=torch.float32)
    return spec.geom(relation, n, hi, wi, ci, ho, wo, co), out


def _tag(relation, g):
    if relation == L.CONV:
        return "conv_forward M=%d N=%d K=%d" % (g.N * g.Ho * g.Wo, g.Co, g.Ci * g.KH * g.KW)
    return "tconv_forward M=%d N=%d K=%d" % (g.N * g.Hi * g.Wi, g.Co * g.KH * g.KW, g.Ci)


def _use_tc(spec, relation, in_shape, out_shape):
    """Whether the launch reading NCHW-shaped ``in_shape`` and writing ``out_shape`` runs on tcgen05."""
    if spec.precision == L.PREC_FP32:
        return False
    if len(in_shape) != 4:
        # batch-sized linears stay on the FFMA kernel — except the wide linear that feeds a feature map
        # (spec.perm).  Forward: 100+ channel tiles of 128 x batch, the fused TPReLU epilogue writes planes;
        # data gradient: a 12800-deep contraction into 256 columns, split 32 ways over K
        if spec.perm is None or len(in_shape) != 2:
            return False
        return tc_supported(spec.geom(relation, in_shape[0], 1, 1, in_shape[1], 1, 1, out_shape[1]))
    n, ci, hi, wi = in_shape
    _, co, ho, wo = out_shape
    return tc_supported(spec.geom(relation, n, hi, wi, ci, ho, wo, co))


SPLIT_K_FORWARD = os.environ.get("GLIS_SPLIT_K_FORWARD", "1") != "0"


def _split_k_forward(g, relation):
    """Run a TPReLU-epilogue launch as split-K sums + one pointwise pass?  Yes when the kernel's own plan
    would split a plain-output launch of this geometry at least 4 ways, or 2 ways over a deep contraction
    (>= 64 k-steps): few pixels and a deep K — D's last level — whose tiles alone leave most SMs idle.
    Measured (tools/tc_microbench.py k1 k8): 256->512 5x5 at batch 64 48.9 -> 25.3 us, at 128 49.6 -> 39.3 us."""
    if not SPLIT_K_FORWARD:
        return False
    ks = int(L.load().glis_conv_tc_ksplit(C.byref(g)))
    if ks >= 4:
        return True
    taps = g.KH * g.KW if relation == L.CONV else -(-g.KH // g.stride_h) * -(-g.KW // g.stride_w)
    return ks >= 2 and taps * ((g.Ci + 63) // 64) >= 64


_SPEC_1X1 = None


def _spec_1x1():
    global _SPEC_1X1
    if _SPEC_1X1 is None:
        _SPEC_1X1 = ContractionSpec(False, (1, 1), (1, 1), (0, 0), (1, 1))
    return _SPEC_1X1


def _is_4x4s2p1(spec):
    return (spec.kernel_size == (4, 4) and spec.stride == (2, 2) and spec.padding == (1, 1)
            and spec.dilation == (1, 1) and tuple(spec.output_padding) == (0, 0))


def image_side_mode(spec, relation, in_shape, out_shape):
    """"unfold" / "fold" when this launch is an image-side 4x4-s2-p1 contraction that runs as a 1x1
    product on the tensor cores (csrc/image_side.cu), else None."""
    if spec.precision == L.PREC_FP32 or len(in_shape) != 4 or not _is_4x4s2p1(spec):
        return None
    n, ci, hi, wi = in_shape
    _, co, ho, wo = out_shape
    if relation == L.CONV and ci <= 4 and co % 8 == 0 and co >= 32 and hi == 2 * ho and wi == 2 * wo and wo <= 256:
        return "unfold"
    if relation == L.TCONV and co <= 4 and ci % 8 == 0 and ci >= 32 and ho == 2 * hi and wo == 2 * wi and wi <= 256:
        return "fold"
    return None


def unfolded_planes(x, lo=True):
    """bf16 hi/lo planes [N, H/2, W/2, 16*C] of the 4x4-s2-p1 patches of ``x`` (fp32 NHWC-dense, C <= 4);
    cached on the tensor so that a layer's forward / data gradient and its weight gradient share them."""
    tag = getattr(x, "_glis_unfolded", None)
    if tag is not None and tag[2] == x.data_ptr() and tag[3] == x._version and (tag[1] is not None or not lo):
        return tag[0], tag[1]
    _require_f32(x, "an unfold")
    n, c, h, w = x.shape
    hi = torch.empty((n, h // 2, w // 2, 16 * c), device=x.device, dtype=torch.bfloat16)
    lo_t = torch.empty_like(hi) if lo else None
    L.call("glis_unfold4x4s2_bf16", L.ptr(x), n, h, w, c, L.ptr16(hi), L.ptr16(lo_t), L.stream())
    try:
        x._glis_unfolded = (hi, lo_t, x.data_ptr(), x._version)
    except AttributeError:
        pass
    return hi, lo_t


def _launch_image_side(mode, spec, relation, x, out_shape, pw, forward_pack, bias, act, act_a, act_b,
                       want_preact, want_planes, want_f32=True, out_buf=None):
    """The two image-side launches as 1x1 tensor-core products (see image_side_mode)."""
    prec = spec.precision
    lo = prec == L.PREC_BF16X3
    _need_matrix(pw, lo)
    n, ci, hi, wi = x.shape
    _, co, ho, wo = out_shape
    s1 = _spec_1x1()
    if mode == "unfold":
        # out[pix][co] = sum_j unfold(x)[pix][j] * M[co][j]:  M = E (conv forward) or E (transposed dgrad)
        xp = unfolded_planes(x, lo)
        j = 16 * ci
        g = s1.geom(L.CONV, n, ho, wo, j, ho, wo, co)
        out = _empty_nhwc(n, co, ho, wo, x)
        preact = torch.empty_like(out) if want_preact else None
        planes = None
        if want_planes:
            planes = (torch.empty_like(out, dtype=torch.bfloat16),
                      torch.empty_like(out, dtype=torch.bfloat16) if lo else None)
        wp = pw.mat
        skip_f32 = planes is not None and not want_f32       # the consumer reads the planes only
        ep = L.Epilogue(L.ptr(bias), act, L.ptr(act_a), L.ptr(act_b), L.ptr(preact), None, None)
        with L.timed("image_side unfold M=%d N=%d K=%d tc" % (n * ho * wo, co, j)):
            L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xp[0]), L.ptr16(xp[1]), L.ptr16(wp[0]), L.ptr16(wp[1]),
                   C.byref(ep), None if skip_f32 else L.ptr(out), L.ptr16(planes[0]) if planes else None,
                   L.ptr16(planes[1]) if planes else None, prec, L.stream())
        if skip_f32:
            out._glis_f32_invalid = True
        return out, preact, planes
    # fold: cols[pix][j] = sum_ci x[pix][ci] * E^T[j][ci];  out = fold(cols) + bias
    xp = planes_of(x, lo)
    j = 16 * co
    g = s1.geom(L.CONV, n, hi, wi, ci, hi, wi, j)
    cols = torch.empty((n, hi, wi, j), device=xp[0].device, dtype=torch.float32)
    wp = pw.mat_t
    ep = L.Epilogue(None, L.ACT_NONE, None, None, None, None, None)
    with L.timed("image_side fold M=%d N=%d K=%d tc" % (n * hi * wi, j, ci)):
        L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xp[0]), L.ptr16(xp[1]), L.ptr16(wp[0]), L.ptr16(wp[1]),
               C.byref(ep), L.ptr(cols), None, None, prec, L.stream())
    out = _usable_out(out_buf, (n, co, ho, wo), xp[0])
    if out is None:
        out = _empty_nhwc(n, co, ho, wo, xp[0])
    L.call("glis_fold4x4s2", L.ptr(cols), n, hi, wi, co, L.ptr(bias), act, L.ptr(out), L.stream())
    return out, None, None


def launch(spec, relation, x, out_shape, pw, forward_pack, bias=None, act=L.ACT_NONE, act_a=None, act_b=None,
           want_preact=False, want_planes=False, want_f32=True, out_buf=None):
    """One gather-GEMM launch (tensor cores when the geometry tiles, FFMA otherwise).

    ``x``: fp32 NHWC-dense (or 2-D) input; ``forward_pack`` selects the layer's forward
    ([T][Cout][Cin] K-major / [T][Cin][Cout] fp32) or data-gradient packs.
    Returns (out_fp32, preact or None, (hi, lo) planes of out or None).  ``want_f32=False`` (the caller
    knows that every consumer of the output reads its planes): a tensor-core launch that writes planes
    leaves the fp32 buffer unwritten and marks it ``_glis_f32_invalid``.  ``out_buf``: write the fp32 result into
    this caller-owned NHWC buffer (ignored when its shape / layout does not fit).
    """
    in_shape = tuple(x.shape)
    prec = spec.precision
    lo = prec == L.PREC_BF16X3
    mode = image_side_mode(spec, relation, in_shape, tuple(out_shape))
    if mode == "fold" and (act not in (L.ACT_NONE, L.ACT_SIGMOID) or want_preact or want_planes):
        mode = None
    if mode == "unfold" and isinstance(x, PlanesOnly):
        mode = None
    if mode is not None:
        return _launch_image_side(mode, spec, relation, x, out_shape, pw, forward_pack, bias, act, act_a, act_b,
                                  want_preact, want_planes, want_f32, out_buf)
    like = x._glis_planes_only[0] if isinstance(x, PlanesOnly) else x
    g, out = _launch_geom(spec, relation, in_shape, out_shape, like, out_buf)
    preact = torch.empty_like(out) if want_preact else None
    use_tc = _use_tc(spec, relation, in_shape, out_shape)
    planes = None
    if want_planes and prec != L.PREC_FP32 and (out.dim() == 4 or spec.perm is not None) and g.Co > 4:
        planes = (torch.empty_like(out, dtype=torch.bfloat16),
                  torch.empty_like(out, dtype=torch.bfloat16) if lo else None)
    if use_tc:
        pw.need_bf16(forward_pack, not forward_pack, lo)
        wp = pw.fwd if forward_pack else pw.bwd
        xp = planes_of(x, lo)
        skip_f32 = planes is not None and not want_f32
        act_ch = spec.perm[0] if (spec.perm and act == L.ACT_TPRELU) else 0
        if act == L.ACT_TPRELU and planes is not None and g.Co % 4 == 0 and _split_k_forward(g, relation):
            # split-K partial sums (+ bias) into one slab per share — stores, not atomics: the forward pass stays
            # bit-reproducible — then sum + TPReLU + planes as one pointwise pass
            ks = int(L.load().glis_conv_tc_ksplit(C.byref(g)))
            slabs = torch.empty((ks,) + tuple(out.shape), device=out.device, dtype=torch.float32)
            ep = L.Epilogue(L.ptr(bias), L.ACT_NONE, None, None, None, None, None, 0, ks)
            with L.timed(_tag(relation, g) + " tc"):
                L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xp[0]), L.ptr16(xp[1]), L.ptr16(wp[0]),
                       L.ptr16(wp[1]), C.byref(ep), L.ptr(slabs), None, None, prec, L.stream())
            L.call("glis_tprelu_forward_planes_sum", L.ptr(slabs), ks, out.numel(), L.ptr(act_a), L.ptr(act_b),
                   L.ptr(preact), None if skip_f32 else L.ptr(out), L.ptr16(planes[0]), L.ptr16(planes[1]),
                   out.numel(), g.Co, act_ch, L.stream())
        else:
            ep = L.Epilogue(L.ptr(bias), act, L.ptr(act_a), L.ptr(act_b), L.ptr(preact), None, None, act_ch)
            with L.timed(_tag(relation, g) + " tc"):
                L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xp[0]), L.ptr16(xp[1]), L.ptr16(wp[0]),
                       L.ptr16(wp[1]), C.byref(ep), None if skip_f32 else L.ptr(out),
                       L.ptr16(planes[0]) if planes else None, L.ptr16(planes[1]) if planes else None, prec,
                       L.stream())
        if skip_f32:
            out._glis_f32_invalid = True
    else:
        if isinstance(x, PlanesOnly):
            raise RuntimeError("glis_b200: planes-only gradient reached an fp32 kernel (internal planning error)")
        _require_f32(x, "an fp32 contraction")
        pw.need_fp32(forward_pack, not forward_pack)
        wp = pw.io if forward_pack else pw.oi
        ep = L.Epilogue(L.ptr(bias), act, L.ptr(act_a), L.ptr(act_b), L.ptr(preact),
                        L.ptr16(planes[0]) if planes else None, L.ptr16(planes[1]) if planes else None,
                        spec.perm[0] if (spec.perm and act == L.ACT_TPRELU) else 0)
        with L.timed(_tag(relation, g) + " fp32"):
            L.call("glis_conv_forward", C.byref(g), L.ptr(x), L.ptr(wp), C.byref(ep), L.ptr(out), L.PREC_FP32,
                   L.stream())
    return out, preact, planes


def conv_forward(spec, relation, x_nhwc, wpack, out_shape_nchw, bias=None, act=L.ACT_NONE,
                 act_a=None, act_b=None, preact=None):
    """fp32 FFMA gather-GEMM on an explicit fp32 pack (kept for direct kernel tests)."""
    g, out = _launch_geom(spec, relation, tuple(x_nhwc.shape), out_shape_nchw, x_nhwc)
    ep = L.Epilogue(L.ptr(bias), act, L.ptr(act_a), L.ptr(act_b), L.ptr(preact), None, None)
    with L.timed(_tag(relation, g) + " fp32"):
        L.call("glis_conv_forward", C.byref(g), L.ptr(x_nhwc), L.ptr(wpack), C.byref(ep), L.ptr(out),
               L.PREC_FP32, L.stream())
    return out


# ---------------------------------------------------------------------------- the WN layer operator
def _layer_shapes(xc, weight, spec):
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
    return shape


def layer_out_shape(x_shape, weight, spec):
    """NCHW shape of a 4-D WN layer's output for an input of shape ``x_shape``."""
    n, _, h, w = x_shape
    ho, wo = spec.out_hw(h, w)
    return (n, weight.shape[1 if spec.transposed else 0], ho, wo)


def _dense_grad(p):
    """``p.grad`` if gradients may be added into it in place (flat-buffer parameters), else None."""
    g = getattr(p, "grad", None)
    if g is None or not getattr(p, "_glis_direct_grad", False) or not g.is_contiguous() or g.dtype != torch.float32:
        return None
    return g


def _take_scratch(weight):
    """Zero-filled buffer for the raw filter gradient: a slice of the owner's per-step scratch
    (zeroed once per backward with the gradients) or a fresh zeros tensor."""
    sc = getattr(weight, "_glis_scratch", None)
    if sc is not None:
        # the weight-gradient kernels ADD into this buffer; the owner zero-fills it once per zero_grad.  A second
        # backward through the same layer before the next zero_grad (the reference's own D pattern of two
        # `.backward()` calls, gradient accumulation) must not see the first one's sums: clear it here
        # (the caller clears it on the stream the weight gradient runs on, behind the previous projection)
        dirty = getattr(weight, "_glis_scratch_dirty", False)
        weight._glis_scratch_dirty = True
        return sc, dirty
    return torch.zeros_like(weight, memory_format=torch.contiguous_format), False


def _touch_hooks(*params):
    """Gradients written in place bypass autograd's accumulation, hence its post-accumulate hooks
    (the overlapped gradient exchange counts on them): fire them by hand."""
    for p in params:
        if p is None:
            continue
        for hook in getattr(p, "_glis_grad_hooks", ()):
            hook(p)


def _touch_hooks_part(p, part):
    """Announce that rows chunk ``part`` of ``p``'s gradient is complete (see dp.OverlappedGradSync)."""
    for hook in getattr(p, "_glis_grad_hooks", ()):
        hook(p, part)


class Overlap(object):
    """Weight-gradient work (wgrad kernel + weight-norm projection) on a SIDE stream.

    Nothing reads a parameter gradient before the optimizer, while the data-gradient chain is the
    critical path of backward and is made of many short kernels that leave most SMs idle: between
    ``Overlap.begin()`` and ``Overlap.join()`` (the trainer brackets its backward passes with them)
    every layer forks its weight gradient onto one side stream behind an event, so it fills the
    machine under the data-gradient / TPReLU-backward kernels.  Stream-ordered only, hence
    capturable into the step's CUDA graph (fork/join inside the capture).  Tensors the side stream
    reads are kept alive until the join."""

    enabled = os.environ.get("GLIS_OVERLAP_WGRAD", "1") != "0"
    # one glis_wn_project_multi launch per backward pass instead of one launch per layer: measured SLOWER inside the
    # step (1.95 vs 1.90 ms at config 2) — the per-layer projections are short kernels that run beside the main
    # stream's data-gradient kernels for free, the batched one sits at the end of backward, in front of the optimizer
    defer_projections = os.environ.get("GLIS_MULTI_PROJECT", "0") != "0"
    _projections = []
    _on = False
    _side = None
    _main = None
    _dirty = False
    _keep = []

    @classmethod
    def active(cls):
        return cls._on

    @classmethod
    def side_stream(cls):
        return cls._side if cls._on else None

    @classmethod
    def main_stream(cls):
        return cls._main if cls._on else None

    @classmethod
    def stream(cls):
        if cls._side is None:
            cls._side = torch.cuda.Stream()
        return cls._side

    # A second side stream for LIGHT weight gradients (the LIS linears: a few dozen blocks, ~7 us): at the end of
    # G's backward they would otherwise queue behind the generator head's 13 MB weight gradient although they only
    # need a fraction of the machine — and that queue is the tail of the iteration.
    light_enabled = os.environ.get("GLIS_LIGHT_STREAM", "1") != "0"
    _light = None
    _dirty_light = False

    @classmethod
    def light_stream(cls):
        if cls._light is None:
            cls._light = torch.cuda.Stream()
        return cls._light

    @classmethod
    def side_streams(cls):
        """Every stream weight gradients may have been enqueued on since ``begin()``."""
        if not cls._on:
            return []
        return [cls._side] + ([cls._light] if cls._dirty_light else [])

    _aux = []

    @classmethod
    def aux_streams(cls, n):
        """``n`` more streams for short independent bursts (parallel pack builds)."""
        while len(cls._aux) < n:
            cls._aux.append(torch.cuda.Stream())
        return cls._aux[:n]

    @classmethod
    def begin(cls):
        if not cls.enabled or not torch.cuda.is_available():
            return
        cls.stream()
        cls._main = torch.cuda.current_stream()
        cls._on, cls._dirty, cls._dirty_light = True, False, False

    @classmethod
    def run(cls, fn, keep=(), light=False):
        light = light and cls.light_enabled
        side = cls.light_stream() if light else cls._side
        ready = torch.cuda.Event()
        ready.record(cls._main)               # everything the side work reads has been enqueued
        side.wait_event(ready)
        if light:
            cls._dirty_light = True           # (before fn: gradient hooks fired inside it ask side_streams())
        with torch.cuda.stream(side):
            fn()
        cls._keep.append(keep)
        if not light:
            cls._dirty = True

    @classmethod
    def queue_projection(cls, graw, w, scale, norm, out_axis, cout, cin, t, c, dw, dscale, acc):
        d = L.WnProj()
        d.G, d.w, d.scale, d.norm = graw.data_ptr(), w.data_ptr(), _dp(scale), norm.data_ptr()
        d.dw, d.dscale = dw.data_ptr(), _dp(dscale)
        d.out_axis, d.Cout, d.Cin, d.T, d.accumulate, d.c = out_axis, cout, cin, t, int(acc), c
        cls._projections.append((d, (graw, w, scale, norm, dw, dscale)))

    @classmethod
    def flush_projections(cls):
        """The deferred weight-norm projections of this backward pass as one launch (on the side stream, behind
        the weight gradients they read)."""
        if not cls._projections:
            return
        items, cls._projections = cls._projections, []
        arr = (L.WnProj * len(items))(*[d for d, _ in items])
        with torch.cuda.stream(cls._side):
            L.call("glis_wn_project_multi", arr, len(items), L.stream(), kernels=(len(items) + 23) // 24)
        cls._keep.append(items)

    @classmethod
    def join(cls):
        """The main stream waits for the side stream; call before anything reads a parameter gradient."""
        if not cls._on:
            return
        cls.flush_projections()
        if cls._dirty:
            cls._main.wait_stream(cls._side)
        if cls._dirty_light:
            cls._main.wait_stream(cls._light)
        cls._keep = []
        cls._on, cls._dirty, cls._dirty_light = False, False, False


class PlanesOnly(object):
    """Stand-in for a gradient that exists ONLY as bf16 hi/lo planes (every consumer of the layer's
    output gradient runs on tensor cores, so the fp32 copy is never written)."""

    def __init__(self, shape, planes):
        self.shape, self._glis_planes_only = tuple(shape), planes

    def dim(self):
        return len(self.shape)


def _backward_plan(spec, pw, x_shape, dy_shape, need_dx, need_dw):
    """(dgrad on tensor cores?, wgrad on tensor cores?) for a layer with input ``x_shape`` (NCHW-shaped)."""
    if spec.precision == L.PREC_FP32 or len(x_shape) != 4:
        return False, False
    n, cin, h, w = x_shape
    _, cout, ho, wo = dy_shape
    tc_dx = tc_dw = True
    if need_dx:
        rel = L.CONV if spec.transposed else L.TCONV
        # (a "fold" data gradient reads dy as planes; an "unfold" one needs the fp32 dy)
        tc_dx = (_use_tc(spec, rel, tuple(dy_shape), tuple(x_shape))
                 or image_side_mode(spec, rel, tuple(dy_shape), tuple(x_shape)) == "fold")
    if need_dw:
        if _image_side_wgrad(spec, x_shape, dy_shape) is not None:
            tc_dw = not spec.transposed     # conv: small = dy as planes; transposed: unfold(dy) needs fp32 dy
        else:
            g = spec.geom(L.CONV, n, ho, wo, cout, h, w, cin) if spec.transposed else \
                spec.geom(L.CONV, n, h, w, cin, ho, wo, cout)
            tc_dw = bool(L.load().glis_wgrad_tc_supported(C.byref(g)))
    return tc_dx, tc_dw


def _image_side_wgrad(spec, x_shape, dy_shape):
    """(n, hs, ws, ca, c_img) when the layer's weight gradient runs as the 1x1 product
    G[a][j] = sum_pix small[pix][a] * unfold(big)[pix][j] on tcgen05 (csrc/image_side.cu), else None."""
    if spec.precision == L.PREC_FP32 or len(x_shape) != 4 or not _is_4x4s2p1(spec):
        return None
    n, cin, h, w = x_shape
    _, cout, ho, wo = dy_shape
    if spec.transposed:       # small = x (coarse, Cin channels), big = dy (fine, Cout <= 4)
        ok = cout <= 4 and cin % 8 == 0 and cin >= 64 and ho == 2 * h and wo == 2 * w and w <= 64
        return (n, h, w, cin, cout) if ok else None
    ok = cin <= 4 and cout % 8 == 0 and cout >= 64 and h == 2 * ho and w == 2 * wo and wo <= 64
    return (n, ho, wo, cout, cin) if ok else None


def _image_side_wgrad_geom(n, hs, ws, c_img, ca):
    """Geometry of the image-side weight gradient as a 1x1 product over pixels.  A 1x1 product has no spatial
    structure, so the pixel list is presented as rows of 64: the kernel's K tiles are then 64 full rows (a 40-pixel
    image row would fill 40 of 48)."""
    pix = n * hs * ws
    if pix % 64 == 0:
        return _spec_1x1().geom(L.CONV, pix // 64, 1, 64, 16 * c_img, 1, 64, ca)
    return _spec_1x1().geom(L.CONV, n, hs, ws, 16 * c_img, hs, ws, ca)


def consumes_planes_only(spec, weight, x_shape, followed_by_tprelu):
    """True when a WN layer with this ``spec`` / ``weight`` reads an input of (NCHW) shape ``x_shape`` through
    its bf16 planes ONLY — forward on tcgen05 (or as the image-side "fold" product) and weight gradient on
    tcgen05 too — so that whoever produces that input need not write its fp32 copy."""
    if spec.precision == L.PREC_FP32 or len(x_shape) != 4 or spec.perm is not None:
        return False
    out_axis = 1 if spec.transposed else 0
    n, cin, h, w = x_shape
    if cin != weight.shape[1 - out_axis]:
        return False
    cout = weight.shape[out_axis]
    ho, wo = spec.out_hw(h, w)
    out_shape = (n, cout, ho, wo)
    rel = L.TCONV if spec.transposed else L.CONV
    mode = image_side_mode(spec, rel, tuple(x_shape), out_shape)
    if mode == "fold" and followed_by_tprelu:
        mode = None                      # launch() sends a TPReLU epilogue down the generic path
    if mode == "unfold" or (mode is None and not _use_tc(spec, rel, tuple(x_shape), out_shape)):
        return False
    if _image_side_wgrad(spec, tuple(x_shape), out_shape) is not None:
        return spec.transposed           # transposed: x is the small operand (planes); conv: unfold(x) reads fp32
    g = spec.geom(L.CONV, n, ho, wo, cout, h, w, cin) if spec.transposed else \
        spec.geom(L.CONV, n, h, w, cin, ho, wo, cout)
    return bool(L.load().glis_wgrad_tc_supported(C.byref(g)))


def _layer_backward(spec, pw, xc, dyc, dy_planes, need_dx, need_dw, need_dscale, need_dbias, bias_shape,
                    pw_bias=None):
    """dgrad + wgrad + weight-norm projection + bias gradient of one WN layer.
    ``dyc``: fp32 gradient w.r.t. the layer's affine output (NHWC-dense), or ``PlanesOnly``."""
    weight, scale = pw.weight, pw.scale
    cout, cin, t = pw.cout, pw.cin, pw.t
    prec = spec.precision
    lo = prec == L.PREC_BF16X3
    if dy_planes is not None and not isinstance(dyc, PlanesOnly):
        attach_planes(dyc, dy_planes)
    if xc.dim() == 4:
        n, _, h, w = xc.shape
        ho, wo = dyc.shape[2], dyc.shape[3]
    else:
        n, h, w, ho, wo = xc.shape[0], 1, 1, 1, 1

    dw = dscale = dbias = None
    # batch-sized linear layers: weight gradient + weight-norm projection as ONE kernel (no raw gradient in memory)
    fused_lin = ((need_dw or need_dscale) and xc.dim() == 2 and t == 1 and not spec.transposed
                 and not isinstance(dyc, PlanesOnly) and FUSED_LINEAR_WGRAD
                 and bool(L.load().glis_linear_wgrad_project_supported(n, cout, cin)))
    if need_dw or need_dscale:
        graw, graw_stale = (None, False) if fused_lin else _take_scratch(weight)
        if spec.transposed:   # small = x (Cin), big = dy (Cout)
            g = spec.geom(L.CONV, n, ho, wo, cout, h, w, cin)
            small, big = xc, dyc
        else:                 # small = dy (Cout), big = x (Cin)
            g = spec.geom(L.CONV, n, h, w, cin, ho, wo, cout)
            small, big = dyc, xc
        tag = "conv_wgrad M=%d N=%d K=%d" % (g.Co, g.Ci * t, n * g.Ho * g.Wo)
        isw = _image_side_wgrad(spec, tuple(xc.shape), tuple(dyc.shape)) if xc.dim() == 4 else None
        use_tc_w = isw is None and prec != L.PREC_FP32 and xc.dim() == 4 and \
            bool(L.load().glis_wgrad_tc_supported(C.byref(g)))
        # operands of the weight gradient are produced on the main stream (planes may need a split /
        # unfold kernel) — before the fork, so that the data-gradient launch shares them
        sp = bp = None
        if isw is not None:
            sp, bp = planes_of(small, lo), unfolded_planes(big, lo)
        elif use_tc_w:
            sp, bp = planes_of(small, lo), planes_of(big, lo)
        direct_w = _dense_grad(weight)
        direct_s = _dense_grad(scale) if scale is not None else None
        pw.need_fp32(False, False)  # the norm
        wc = weight.detach().contiguous()
        sc = None if scale is None else scale.detach().contiguous()
        # Parameters owned by a FlatParams buffer carry a dense `.grad`: add into it in place and
        # hand autograd nothing to accumulate (one kernel less per parameter); otherwise return it.
        in_place = direct_w is not None and (scale is None or direct_s is not None)
        fork = Overlap.active() and in_place     # results nobody reads before the optimizer: side stream
        if in_place:
            dw_buf, ds_buf, acc = direct_w, direct_s, 1
        else:
            dw_buf = torch.empty_like(weight, memory_format=torch.contiguous_format)
            ds_buf = torch.empty(cout, device=dw_buf.device, dtype=torch.float32) if scale is not None else None
            acc = 0

        # tensor-core weight gradients in their deterministic form: one slab per K split, summed in order by the
        # projection (no atomics, no zero-filled scratch)
        slabs = None
        if DETERMINISTIC_WGRAD and (isw is not None or use_tc_w):
            g_w = g
            if isw is not None:
                n_, hs, ws, ca, c_img = isw
                g_w = _image_side_wgrad_geom(n_, hs, ws, c_img, ca)
            n_slabs = int(L.load().glis_wgrad_tc_splits(C.byref(g_w)))
            if n_slabs > 0:
                slabs = torch.empty((n_slabs, weight.numel()), device=weight.device, dtype=torch.float32)

        def weight_gradient():
            if slabs is not None:
                with L.timed(tag + (" tc (image side)" if isw is not None else " tc")):
                    L.call("glis_conv_wgrad_bf16_slabs", C.byref(g_w), L.ptr16(sp[0]), L.ptr16(sp[1]), L.ptr16(bp[0]),
                           L.ptr16(bp[1]), L.ptr(slabs), slabs.shape[0], prec, L.stream())
                if slabs.shape[0] > 2 and weight.numel() % 4 == 0:
                    # many slabs: one streaming pass adds them (in slab order) into the raw-gradient buffer, then the
                    # usual projection; reading 36 slabs per element inside the projection was measured 3x slower
                    L.call("glis_slab_reduce", L.ptr(slabs), slabs.shape[0], slabs.shape[1], L.ptr(graw), weight.numel(),
                           L.stream())
                    L.call("glis_wn_project", L.ptr(graw), L.ptr(wc), L.ptr(sc), L.ptr(pw.norm), pw.out_axis, cout, cin,
                           t, spec.norm_factor, L.ptr(dw_buf), L.ptr(ds_buf), acc, L.stream())
                else:
                    L.call("glis_wn_project_slabs", L.ptr(slabs), slabs.shape[0], slabs.shape[1], L.ptr(wc), L.ptr(sc),
                           L.ptr(pw.norm), pw.out_axis, cout, cin, t, spec.norm_factor, L.ptr(dw_buf), L.ptr(ds_buf),
                           acc, L.stream())
                if acc:
                    _touch_hooks(weight, scale)
                return
            if fused_lin:
                _require_f32(small, "an fp32 weight gradient")
                _require_f32(big, "an fp32 weight gradient")
                pc, pp = spec.perm or (0, 0)
                # a large gradient that a data-parallel exchange wants in pieces (dp.OverlappedGradSync marks the
                # parameter): one launch per row chunk, each announced as soon as it is enqueued
                parts = getattr(weight, "_glis_grad_parts", None) if acc else None
                for k, (r0, rc) in enumerate(parts or [(0, cout)]):
                    with L.timed(tag + " fp32 (fused projection)"):
                        L.call("glis_linear_wgrad_project", L.ptr(small), L.ptr(big), L.ptr(wc), L.ptr(sc),
                               L.ptr(pw.norm), L.ptr(dw_buf), L.ptr(ds_buf), n, cout, cin, pc, pp, acc, r0, rc,
                               L.stream())
                    if parts:
                        _touch_hooks_part(weight, k)
                if acc:
                    _touch_hooks(weight, scale)
                return
            if graw_stale:
                graw.zero_()
            if isw is not None:
                n_, hs, ws, ca, c_img = isw
                g1 = _image_side_wgrad_geom(n_, hs, ws, c_img, ca)
                with L.timed(tag + " tc (image side)"):
                    L.call("glis_conv_wgrad_bf16", C.byref(g1), L.ptr16(sp[0]), L.ptr16(sp[1]), L.ptr16(bp[0]),
                           L.ptr16(bp[1]), L.ptr(graw), prec, L.stream())
            elif use_tc_w:
                with L.timed(tag + " tc"):
                    L.call("glis_conv_wgrad_bf16", C.byref(g), L.ptr16(sp[0]), L.ptr16(sp[1]), L.ptr16(bp[0]),
                           L.ptr16(bp[1]), L.ptr(graw), prec, L.stream())
            elif spec.perm is not None:
                _require_f32(big, "an fp32 weight gradient")
                with L.timed(tag + " fp32 (permuted rows)"):
                    L.call("glis_linear_wgrad", L.ptr(small), L.ptr(big), L.ptr(graw), n, cout, cin, spec.perm[0],
                           spec.perm[1], L.stream())
            else:
                _require_f32(small, "an fp32 weight gradient")
                _require_f32(big, "an fp32 weight gradient")
                with L.timed(tag + " fp32"):
                    L.call("glis_conv_wgrad", C.byref(g), L.ptr(small), L.ptr(big), L.ptr(graw), L.PREC_FP32,
                           L.stream())
            hooked = getattr(weight, "_glis_grad_hooks", None) or (scale is not None and
                                                                   getattr(scale, "_glis_grad_hooks", None))
            if fork and acc and not hooked and Overlap.defer_projections:
                # nobody waits for this gradient before the join: its projection joins the network's other ones in
                # ONE glis_wn_project_multi launch at Overlap.join() (a gradient exchange that fires per parameter
                # — data parallelism — keeps the per-layer launch so that buckets leave as early as they can)
                Overlap.queue_projection(graw, wc, sc, pw.norm, pw.out_axis, cout, cin, t, spec.norm_factor, dw_buf,
                                         ds_buf, acc)
                return
            L.call("glis_wn_project", L.ptr(graw), L.ptr(wc), L.ptr(sc), L.ptr(pw.norm), pw.out_axis, cout, cin, t,
                   spec.norm_factor, L.ptr(dw_buf), L.ptr(ds_buf), acc, L.stream())
            if acc:
                _touch_hooks(weight, scale)

        if fork:
            Overlap.run(weight_gradient, keep=(small, big, sp, bp, wc, sc, graw, slabs),
                        light=fused_lin and cout * cin <= (1 << 18))
        else:
            weight_gradient()
        if not acc:
            dw = dw_buf
            dscale = None if ds_buf is None else ds_buf.view_as(scale)

    dx = None
    if need_dx:
        # conv layer: dx gathers dy through the transposed relation; transposed layer: the direct one
        rel = L.CONV if spec.transposed else L.TCONV
        dx, _, _ = launch(spec, rel, dyc, tuple(xc.shape), pw, forward_pack=False)

    if need_dbias:
        direct_b = _dense_grad(pw_bias) if pw_bias is not None else None
        if direct_b is not None:
            L.call("glis_channel_sum", L.ptr(dyc), L.ptr(direct_b), dyc.numel(), cout, 1, 1, L.stream())
            _touch_hooks(pw_bias)
        else:
            dbias = torch.empty(cout, device=dyc.device, dtype=torch.float32)
            L.call("glis_channel_sum", L.ptr(dyc), L.ptr(dbias), dyc.numel(), cout, 1, 0, L.stream())
            dbias = dbias.view(bias_shape)
    return dx, dw, dscale, dbias


class WNContraction(torch.autograd.Function):
    """``norm_scale_bias(F.conv2d / F.conv_transpose2d / F.linear (x, w))``.

    Reference: common/modules/WeightNormalizedConv.py:79-81, :96-99, :29-49 and
    common/modules/WeightNormalizedLinear.py:30-42.  The per-channel ``scale/norm`` is
    folded into the packed weights (one pass over the parameters per optimizer step), the
    bias into the GEMM epilogue; backward applies the closed-form projection of SURVEY.md App. E.
    """

    @staticmethod
    def forward(ctx, x, weight, scale, bias, spec):
        xc = _nhwc(x)
        if getattr(x, "_glis_planes", None) is not None and xc is x:
            pass  # planes travel with the tensor object
        shape = _layer_shapes(xc, weight, spec)
        pw = packed_weights(weight, scale, spec)
        b = None if bias is None else bias.detach().reshape(-1).contiguous()
        rel_f = L.TCONV if spec.transposed else L.CONV
        out, _, _ = launch(spec, rel_f, xc, shape, pw, forward_pack=True, bias=b)
        ctx.spec, ctx.pw, ctx.bias_param = spec, pw, bias
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        ctx.x_planes = getattr(xc, "_glis_planes", None)
        ctx.x_unfolded = getattr(xc, "_glis_unfolded", None)
        ctx.x_f32_invalid = getattr(xc, "_glis_f32_invalid", False)
        ctx.save_for_backward(xc)
        return out

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        if ctx.x_planes is not None:
            xc._glis_planes = ctx.x_planes
        if ctx.x_unfolded is not None:
            xc._glis_unfolded = ctx.x_unfolded
        if ctx.x_f32_invalid:
            xc._glis_f32_invalid = True
        dyc = _nhwc(dy)
        ni = ctx.needs_input_grad
        dx, dw, dscale, dbias = _layer_backward(ctx.spec, ctx.pw, xc, dyc, None, ni[0], ni[1],
                                                ctx.pw.scale is not None and ni[2],
                                                ctx.bias_shape is not None and ni[3], ctx.bias_shape, ctx.bias_param)
        return dx, dw, dscale, dbias, None


def wn_contraction(x, weight, scale, bias, spec):
    return WNContraction.apply(x, weight, scale, bias, spec)


class WNContractionSigmoid(torch.autograd.Function):
    """A WN layer followed by ``nn.Sigmoid`` — the last two modules of the generators (common/model.py:131-136,
    :256-259) — with the sigmoid in the contraction's epilogue (for the 3-channel image layer: in the fold
    kernel).  Backward: ``dy = dout * s * (1 - s)`` from the saved output, then the layer's backward."""

    @staticmethod
    def forward(ctx, x, weight, scale, bias, spec, out_buf=None):
        xc = _nhwc(x)
        shape = _layer_shapes(xc, weight, spec)
        pw = packed_weights(weight, scale, spec)
        b = None if bias is None else bias.detach().reshape(-1).contiguous()
        rel_f = L.TCONV if spec.transposed else L.CONV
        out, _, _ = launch(spec, rel_f, xc, shape, pw, forward_pack=True, bias=b, act=L.ACT_SIGMOID, out_buf=out_buf)
        ctx.spec, ctx.pw, ctx.bias_param = spec, pw, bias
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        ctx.x_planes = getattr(xc, "_glis_planes", None)
        ctx.x_unfolded = getattr(xc, "_glis_unfolded", None)
        ctx.x_f32_invalid = getattr(xc, "_glis_f32_invalid", False)
        ctx.save_for_backward(xc, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, s = ctx.saved_tensors
        if ctx.x_planes is not None:
            xc._glis_planes = ctx.x_planes
        if ctx.x_unfolded is not None:
            xc._glis_unfolded = ctx.x_unfolded
        if ctx.x_f32_invalid:
            xc._glis_f32_invalid = True
        dout_c = _nhwc(dout)
        dy = torch.empty_like(s)
        L.call("glis_sigmoid_backward", L.ptr(s), L.ptr(dout_c), L.ptr(dy), s.numel(), L.stream())
        ni = ctx.needs_input_grad
        dx, dw, dscale, dbias = _layer_backward(ctx.spec, ctx.pw, xc, dy, None, ni[0], ni[1],
                                                ctx.pw.scale is not None and ni[2],
                                                ctx.bias_shape is not None and ni[3], ctx.bias_shape, ctx.bias_param)
        return dx, dw, dscale, dbias, None, None


def wn_contraction_sigmoid(x, weight, scale, bias, spec, out=None):
    """``out``: (under no_grad) a caller-owned NHWC buffer the images are written into."""
    if out is not None and not torch.is_grad_enabled():
        return WNContractionSigmoid.apply(x, weight, scale, bias, spec, out)
    return WNContractionSigmoid.apply(x, weight, scale, bias, spec)


class WNContractionTPReLU(torch.autograd.Function):
    """A WN layer followed by TPReLU as ONE forward kernel: the GEMM epilogue applies the bias and
    the translated PReLU, stores the pre-activation for backward and (tensor-core mode) the bf16
    hi/lo planes of the activated output for the next layer.  Backward runs the TPReLU gradient as
    one pass that writes the layer's output gradient directly as planes.
    Reference: the (conv, tprelu) pairs of common/model.py:31-45, :112-127, :222-251."""

    @staticmethod
    def forward(ctx, x, weight, scale, bias, a_raw, b_t, spec, flags=(True, True)):
        # flags = (a backward may follow: keep the pre-activation, some consumer reads the fp32 output)
        need_bwd, want_f32 = flags
        xc = _nhwc(x)
        shape = _layer_shapes(xc, weight, spec)
        pw = packed_weights(weight, scale, spec)
        b = None if bias is None else bias.detach().reshape(-1).contiguous()
        rel_f = L.TCONV if spec.transposed else L.CONV
        out, preact, planes = launch(spec, rel_f, xc, shape, pw, forward_pack=True, bias=b, act=L.ACT_TPRELU,
                                     act_a=a_raw.detach().contiguous(), act_b=b_t.detach().contiguous(),
                                     want_preact=need_bwd, want_planes=True, want_f32=want_f32)
        _ELIDED[0] = getattr(out, "_glis_f32_invalid", False)
        PreactTap.report(a_raw, b_t, preact)
        ctx.spec, ctx.pw, ctx.bias_param = spec, pw, bias
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        ctx.x_planes = getattr(xc, "_glis_planes", None)
        ctx.x_unfolded = getattr(xc, "_glis_unfolded", None)
        ctx.x_f32_invalid = getattr(xc, "_glis_f32_invalid", False)
        if need_bwd:
            ctx.save_for_backward(xc, preact, a_raw, b_t)
        ctx.set_materialize_grads(False)   # no zero tensors for the (non-differentiable) plane outputs
        if planes is None:
            return out, None, None
        ctx.mark_non_differentiable(planes[0])
        if planes[1] is not None:
            ctx.mark_non_differentiable(planes[1])
        return out, planes[0], planes[1]

    @staticmethod
    def backward(ctx, dout, _dhi, _dlo):
        xc, preact, a_raw, b_t = ctx.saved_tensors
        if ctx.x_planes is not None:
            xc._glis_planes = ctx.x_planes
        if ctx.x_unfolded is not None:
            xc._glis_unfolded = ctx.x_unfolded
        if ctx.x_f32_invalid:
            xc._glis_f32_invalid = True
        spec = ctx.spec
        c = a_raw.numel()
        if dout is None:
            dout = torch.zeros_like(preact)
        doc = _nhwc(dout)
        lo = spec.precision == L.PREC_BF16X3
        want_planes = spec.precision != L.PREC_FP32 and doc.dim() == 4
        ni = ctx.needs_input_grad
        need_dw = ni[1] or (ctx.pw.scale is not None and ni[2])
        need_db = ctx.bias_shape is not None and ni[3]
        tc_dx, tc_dw = _backward_plan(spec, ctx.pw, tuple(xc.shape), tuple(preact.shape), ni[0], need_dw)
        # every consumer of dy on tensor cores (and no bias gradient): the fp32 copy is never read
        planes_only = want_planes and tc_dx and tc_dw and not need_db and (ni[0] or need_dw)
        dy = None if planes_only else torch.empty_like(preact)
        dy_hi = torch.empty_like(preact, dtype=torch.bfloat16) if want_planes else None
        dy_lo = torch.empty_like(preact, dtype=torch.bfloat16) if (want_planes and lo) else None
        ga, gb = _dense_grad(a_raw), _dense_grad(b_t)
        want_ab = ctx.needs_input_grad[4] or ctx.needs_input_grad[5]
        direct = ga is not None and gb is not None and ctx.needs_input_grad[4] and ctx.needs_input_grad[5]
        if not want_ab:   # frozen TPReLU parameters (D during the G update): no sums, no buffers
            da = db = None
        elif direct:      # the kernel adds atomically: straight into the (zero-filled) flat gradients
            da, db = ga, gb
        else:
            da = torch.zeros(c, device=doc.device, dtype=torch.float32)
            db = torch.zeros(c, device=doc.device, dtype=torch.float32)
        L.call("glis_tprelu_backward_planes", L.ptr(preact), L.ptr(a_raw.detach()), L.ptr(b_t.detach()), L.ptr(doc),
               L.ptr(dy), L.ptr16(dy_hi), L.ptr16(dy_lo), L.ptr(da), L.ptr(db), preact.numel(), c, 1, L.stream())
        if direct:
            _touch_hooks(a_raw, b_t)
            da = db = None
        if planes_only:
            dy = PlanesOnly(preact.shape, (dy_hi, dy_lo))
        dx, dw, dscale, dbias = _layer_backward(spec, ctx.pw, xc, dy, (dy_hi, dy_lo) if want_planes else None,
                                                ni[0], ni[1], ctx.pw.scale is not None and ni[2],
                                                ctx.bias_shape is not None and ni[3], ctx.bias_shape, ctx.bias_param)
        return dx, dw, dscale, dbias, da, db, None, None


_ELIDED = [False]    # did the last fused forward leave its fp32 output unwritten (see launch)


class PreactTap(object):
    """Test hook: ``with PreactTap() as tap:`` collects ``(TPReLU slope Parameter, branch mask)`` of every
    TPReLU forward that keeps its pre-activation for a backward pass (logical ``(N, C, ...)`` shape; the mask is
    taken from the very tensor backward reads it from).  The parity tests hand those masks to the oracle so that both
    sides differentiate the same piecewise-linear function (oracle/flipaware.py)."""

    active = None

    def __init__(self):
        self.records = []

    def __enter__(self):
        self._prev, PreactTap.active = PreactTap.active, self
        return self

    def __exit__(self, *exc):
        PreactTap.active = self._prev
        return False

    @classmethod
    def report(cls, a_raw, b_t, preact):
        if cls.active is not None and preact is not None:
            # the branch mask as the backward kernels will compute it, !(t > 0) with t = y - b in fp32 — taken NOW:
            # by the time a whole training iteration has run the optimizer has moved b
            shape = (1, -1) + (1,) * (preact.dim() - 2)
            neg = ~((preact.detach() - b_t.detach().view(shape)) > 0)
            cls.active.records.append((a_raw, neg))


def _may_need_backward(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def wn_contraction_tprelu(x, weight, scale, bias, a_raw, b_t, spec, planes_suffice=False):
    """``planes_suffice``: the caller guarantees that the result is consumed through its bf16 planes only
    (see ``consumes_planes_only``); its fp32 buffer may then stay unwritten."""
    need_bwd = _may_need_backward(x, weight, scale, bias, a_raw, b_t)
    out, hi, lo = WNContractionTPReLU.apply(x, weight, scale, bias, a_raw, b_t, spec,
                                            (need_bwd, not planes_suffice))
    if hi is not None:
        attach_planes(out, (hi, lo))
    if _ELIDED[0]:
        out._glis_f32_invalid = True
    return out


class WNLinearViewTPReLU(torch.autograd.Function):
    """``WeightNormalizedLinear -> View(C, h, w) -> TPReLU(C)`` — the head of the generators
    (common/model.py:196-212, :94-104) — as ONE kernel.  The weight packs are row-permuted
    (ContractionSpec.perm) so that the GEMM writes its features in NHWC order: the View is free, the
    per-channel TPReLU sits in the epilogue (parameter index = feature % C) together with the
    pre-activation and the bf16 planes for the first transposed convolution; on the tensor cores when
    the precision allows.  Backward: TPReLU backward in NHWC, data gradient through the permuted pack,
    weight gradient written back to master rows."""

    @staticmethod
    def forward(ctx, x, weight, scale, a_raw, b_t, spec, view, flags=(True, True)):
        need_bwd, want_f32 = flags
        xc = _nhwc(x)
        c, h, w = view
        n, n_out = xc.shape[0], weight.shape[0]
        pw = packed_weights(weight, scale, spec)
        out, preact, planes = launch(spec, L.CONV, xc, (n, n_out), pw, forward_pack=True, act=L.ACT_TPRELU,
                                     act_a=a_raw.detach().contiguous(), act_b=b_t.detach().contiguous(),
                                     want_preact=need_bwd, want_planes=True, want_f32=want_f32)
        _ELIDED[0] = getattr(out, "_glis_f32_invalid", False)
        ctx.spec, ctx.pw, ctx.view = spec, pw, view
        if need_bwd:
            ctx.save_for_backward(xc, preact, a_raw, b_t)
        ctx.set_materialize_grads(False)
        as_map = lambda t: t.view(n, h, w, c).permute(0, 3, 1, 2)     # NHWC memory, logical (N, C, h, w)
        if preact is not None:
            PreactTap.report(a_raw, b_t, as_map(preact))
        if planes is None:
            return as_map(out), None, None
        hi, lo = as_map(planes[0]), None if planes[1] is None else as_map(planes[1])
        ctx.mark_non_differentiable(hi)
        if lo is not None:
            ctx.mark_non_differentiable(lo)
        return as_map(out), hi, lo

    @staticmethod
    def backward(ctx, dout, _dhi, _dlo):
        xc, preact, a_raw, b_t = ctx.saved_tensors
        spec, c = ctx.spec, a_raw.numel()
        n = xc.shape[0]
        if dout is None:
            dout = torch.zeros((n,) + tuple(ctx.view), device=xc.device)
        doc = dout.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(n, -1)
        ni = ctx.needs_input_grad
        dy = torch.empty_like(preact)
        # the data gradient runs on tensor cores (split-K over the 12800 features): dy also as planes
        tc_dx = ni[0] and _use_tc(spec, L.TCONV, tuple(preact.shape), tuple(xc.shape))
        lo = spec.precision == L.PREC_BF16X3
        dy_hi = torch.empty_like(preact, dtype=torch.bfloat16) if tc_dx else None
        dy_lo = torch.empty_like(preact, dtype=torch.bfloat16) if (tc_dx and lo) else None
        ga, gb = _dense_grad(a_raw), _dense_grad(b_t)
        want_ab = ni[3] or ni[4]
        direct = ga is not None and gb is not None and ni[3] and ni[4]
        if not want_ab:
            da = db = None
        elif direct:
            da, db = ga, gb
        else:
            da = torch.zeros(c, device=doc.device, dtype=torch.float32)
            db = torch.zeros(c, device=doc.device, dtype=torch.float32)
        L.call("glis_tprelu_backward_planes", L.ptr(preact), L.ptr(a_raw.detach()), L.ptr(b_t.detach()), L.ptr(doc),
               L.ptr(dy), L.ptr16(dy_hi), L.ptr16(dy_lo), L.ptr(da), L.ptr(db), preact.numel(), c, 1, L.stream())
        if direct:
            _touch_hooks(a_raw, b_t)
            da = db = None
        dx, dw, dscale, _ = _layer_backward(spec, ctx.pw, xc, dy, (dy_hi, dy_lo) if tc_dx else None, ni[0], ni[1],
                                            ctx.pw.scale is not None and ni[2], False, None)
        return dx, dw, dscale, da, db, None, None, None


def wn_linear_view_tprelu(x, weight, scale, a_raw, b_t, view, planes_suffice=False):
    """See WNLinearViewTPReLU.  ``view`` = (C, h, w) of the View module between the linear and the TPReLU;
    ``planes_suffice`` as in ``wn_contraction_tprelu``."""
    c, h, w = (int(v) for v in view)
    spec = ContractionSpec(False, (1, 1), (1, 1), (0, 0), (1, 1), linear=True, perm=(c, h * w))
    need_bwd = _may_need_backward(x, weight, scale, a_raw, b_t)
    out, hi, lo = WNLinearViewTPReLU.apply(x, weight, scale, a_raw, b_t, spec, (c, h, w),
                                           (need_bwd, not planes_suffice))
    if hi is not None:
        attach_planes(out, (hi, lo))
    if _ELIDED[0]:
        out._glis_f32_invalid = True
    return out


class LISModuleFunction(torch.autograd.Function):
    """``u + lis(u)`` for one residual LIS block ``linear -> TPReLU -> linear`` (common/model.py:176-192,
    :281-297) as ONE cluster kernel per direction (csrc/lis.cu).  The two weight gradients stay the layer
    operators' (``_layer_backward``: batch-sized fp32 kernel + weight-norm projection, side stream)."""

    @staticmethod
    def forward(ctx, u, w1, w2, a_raw, b_t, spec1, spec2, need_bwd):
        uc = _nhwc(u)
        n, code = uc.shape
        pw1, pw2 = packed_weights(w1, None, spec1), packed_weights(w2, None, spec2)
        pw1.need_fp32(True, False)
        pw2.need_fp32(True, False)
        h = torch.empty_like(uc) if need_bwd else None
        act = torch.empty_like(uc) if need_bwd else None
        out = torch.empty_like(uc)
        L.call("glis_lis_forward", L.ptr(uc), L.ptr(pw1.io), None, L.ptr(a_raw.detach().contiguous()),
               L.ptr(b_t.detach().contiguous()), L.ptr(pw2.io), None, n, code, L.ptr(h), L.ptr(act), L.ptr(out),
               L.stream())
        ctx.pw1, ctx.pw2, ctx.spec1, ctx.spec2 = pw1, pw2, spec1, spec2
        PreactTap.report(a_raw, b_t, h)
        if need_bwd:
            ctx.save_for_backward(uc, h, act, a_raw, b_t)
        return out

    @staticmethod
    def backward(ctx, dout):
        uc, h, act, a_raw, b_t = ctx.saved_tensors
        n, code = uc.shape
        doc = _nhwc(dout)
        ni = ctx.needs_input_grad
        ctx.pw1.need_fp32(False, True)
        ctx.pw2.need_fp32(False, True)
        dh, du = torch.empty_like(uc), torch.empty_like(uc)
        ga, gb = _dense_grad(a_raw), _dense_grad(b_t)
        want_ab = ni[3] or ni[4]
        direct = ga is not None and gb is not None and ni[3] and ni[4]
        if not want_ab:
            da = db = None
        elif direct:      # the kernel adds atomically: straight into the (zero-filled) flat gradients
            da, db = ga, gb
        else:
            da = torch.zeros(code, device=uc.device, dtype=torch.float32)
            db = torch.zeros(code, device=uc.device, dtype=torch.float32)
        # the second linear's weight gradient needs nothing the chain kernel computes (the saved activation and the
        # incoming gradient): forked FIRST, it runs beside that kernel instead of behind it — this is the tail of
        # the generator's backward pass, where every microsecond is exposed
        _, dw2, _, _ = _layer_backward(ctx.spec2, ctx.pw2, act, doc, None, False, ni[2], False, False, None)
        L.call("glis_lis_backward", L.ptr(doc), L.ptr(ctx.pw2.oi), L.ptr(h), L.ptr(a_raw.detach().contiguous()),
               L.ptr(b_t.detach().contiguous()), L.ptr(ctx.pw1.oi), n, code, L.ptr(dh), L.ptr(du), L.ptr(da),
               L.ptr(db), L.stream())
        if direct:
            _touch_hooks(a_raw, b_t)
            da = db = None
        _, dw1, _, _ = _layer_backward(ctx.spec1, ctx.pw1, uc, dh, None, False, ni[1], False, False, None)
        return (du if ni[0] else None), dw1, dw2, da, db, None, None, None


LIS_FUSED = os.environ.get("GLIS_LIS_FUSED", "1") != "0"
FUSED_LINEAR_WGRAD = os.environ.get("GLIS_FUSED_LINEAR_WGRAD", "1") != "0"
# Tensor-core weight gradients with one slab per K split, added in a fixed order (bit-reproducible) instead of atomic
# adds into one buffer.  Opt-in: the slabs are ~230 MB of extra traffic per config-2 iteration, 1.94 instead of 1.88 ms.
DETERMINISTIC_WGRAD = os.environ.get("GLIS_DETERMINISTIC_WGRAD", "0") != "0"


def lis_supported(code):
    return LIS_FUSED and bool(L.load().glis_lis_supported(int(code)))


def lis_module(u, w1, w2, a_raw, b_t, spec1, spec2):
    """``u + linear2(TPReLU(linear1(u)))`` for weight-normalised linears without scale / bias (norm='weight')."""
    return LISModuleFunction.apply(u, w1, w2, a_raw, b_t, spec1, spec2, _may_need_backward(u, w1, w2, a_raw, b_t))


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
        if any(ctx.needs_input_grad):
            PreactTap.report(a_raw, b, xc)
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


def rmsprop_(p_flat, g_flat, v_flat, lr, alpha=0.9, eps=1e-6, gscale=1.0, params=None):
    """Fused RMSprop over flat parameter / gradient / square-average buffers (g_lis/main.py:313-314).
    ``params``: the Parameters living in ``p_flat`` — their cached weight packs are invalidated;
    without it every cached pack is."""
    L.call("glis_rmsprop", L.ptr(p_flat), L.ptr(g_flat), L.ptr(v_flat), p_flat.numel(), lr, alpha, eps, gscale,
           L.stream())
    if params is None:
        bump_param_epoch()
    else:
        for p in params:
            p._glis_epoch = getattr(p, "_glis_epoch", 0) + 1


def randn_(out, seed, offset=0):
    L.call("glis_randn", L.ptr(out), out.numel(), seed, offset, L.stream())
    return out


def uniform_(out, seed, offset=0):
    L.call("glis_uniform", L.ptr(out), out.numel(), seed, offset, L.stream())
    return out


class BCEWithLogitsConst(torch.autograd.Function):
    """mean BCE(sigmoid(logit), t) against a constant target t — nn.Sigmoid + nn.BCELoss of the
    reference's D head (common/model.py:61, g_lis/main.py:311,555,564,578) as one kernel that also
    leaves d(loss)/d(logit); backward only scales it."""

    @staticmethod
    def forward(ctx, logit, target):
        lg = logit.detach().reshape(-1).contiguous()
        loss = torch.empty(1, device=lg.device, dtype=torch.float32)
        dl = torch.empty_like(lg)
        L.call("glis_bce_logits", L.ptr(lg), float(target), lg.numel(), 1.0, L.ptr(loss), L.ptr(dl), None, L.stream())
        ctx.save_for_backward(dl)
        ctx.shape = tuple(logit.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gloss):
        (dl,) = ctx.saved_tensors
        return (dl * gloss).view(ctx.shape), None


def bce_with_logits_const(logit, target):
    return BCEWithLogitsConst.apply(logit, target)


def bce_logits(logit, target, gscale=1.0, want_grad=True, want_prob=False):
    """(loss[1], dlogit or None, prob or None) for mean BCE of sigmoid(logit) against a constant target."""
    lg = logit.detach().reshape(-1).contiguous()
    loss = torch.empty(1, device=lg.device, dtype=torch.float32)
    dl = torch.empty_like(lg) if want_grad else None
    pr = torch.empty_like(lg) if want_prob else None
    L.call("glis_bce_logits", L.ptr(lg), float(target), lg.numel(), float(gscale), L.ptr(loss), L.ptr(dl),
           L.ptr(pr), L.stream())
    return loss, dl, pr


def mse_scaled(u, z, lam, du=None, accumulate=False, unscaled_loss=False):
    """loss[1] = lam * mean((u - z)^2) (or the plain mean with ``unscaled_loss``); du (+)= 2 lam (u - z) / numel."""
    uc, zc = u.detach().contiguous(), z.detach().contiguous()
    loss = torch.empty(1, device=uc.device, dtype=torch.float32)
    L.call("glis_mse_scaled", L.ptr(uc), L.ptr(zc), uc.numel(), float(lam), L.ptr(loss), L.ptr(du),
           (1 if accumulate else 0) | (2 if unscaled_loss else 0), L.stream())
    return loss


def lsq_logits(logit, target, gscale=1.0, want_grad=True, want_prob=False):
    """(loss[1], dlogit or None, prob or None) for --ls: mean((sigmoid(logit) - target)^2) (g_lis/main.py:308-311)."""
    lg = logit.detach().reshape(-1).contiguous()
    loss = torch.empty(1, device=lg.device, dtype=torch.float32)
    dl = torch.empty_like(lg) if want_grad else None
    pr = torch.empty_like(lg) if want_prob else None
    L.call("glis_lsq_logits", L.ptr(lg), float(target), lg.numel(), float(gscale), L.ptr(loss), L.ptr(dl),
           L.ptr(pr), L.stream())
    return loss, dl, pr


# ---------------------------------------------------------------------------- dropout
class DropoutClock(object):
    """Where the dropout masks come from: ``keep = Philox(seed, stream = counter + call, element) >= p``.

    ``counter`` is a device uint64 advanced by ``tick()`` — once per training iteration, inside the captured
    graph — and ``call`` numbers the dropout calls since the last tick (a Python count, frozen per call site by a
    capture).  So eager calls never repeat a mask, a CUDA-graph replay draws fresh masks without any host work,
    and backward regenerates the mask of its forward from (counter, call) instead of storing it.  ``seed`` differs
    per data-parallel rank (dp.seed_everything)."""

    TICK = 1 << 20           # > dropout calls per iteration
    seed = None
    _counter = {}
    calls = 0

    @classmethod
    def counter(cls, device):
        c = cls._counter.get(device)
        if c is None:
            c = cls._counter[device] = torch.zeros(1, device=device, dtype=torch.int64)
        return c

    @classmethod
    def get_seed(cls):
        if cls.seed is None:
            cls.seed = torch.initial_seed() & 0x7fffffffffffffff
        return cls.seed

    @classmethod
    def tick(cls, device):
        L.call("glis_counter_add", C.c_void_p(cls.counter(device).data_ptr()), cls.TICK, L.stream())
        cls.calls = 0

    @classmethod
    def next_call(cls):
        if cls.calls >= cls.TICK:
            raise RuntimeError("glis_b200: %d dropout calls without a DropoutClock.tick()" % cls.calls)
        cls.calls += 1
        return cls.calls - 1


class DropoutFunction(torch.autograd.Function):
    """``nn.Dropout(p)`` (common/model.py:52-53) / ``nn.Dropout2d(p)`` (:344-346) in training mode on the
    counter-based generator of ``glis_dropout``; backward = the same call on the gradient."""

    @staticmethod
    def forward(ctx, x, p, channel_mode):
        xc = _nhwc(x)
        out = torch.empty_like(xc)
        ctx.args = DropoutFunction._args(xc, p, channel_mode, DropoutClock.next_call())
        DropoutFunction._run(xc, out, ctx.args)
        return out

    @staticmethod
    def _args(xc, p, channel_mode, call):
        if channel_mode and xc.dim() != 4:
            raise RuntimeError("glis_b200: Dropout2d expects a 4-D tensor")
        c = xc.shape[1]
        per_image = xc.numel() // max(1, xc.shape[0])
        return (c, per_image, 1 if channel_mode else 0, float(p), DropoutClock.get_seed(),
                DropoutClock.counter(xc.device), call)

    @staticmethod
    def _run(src, dst, args):
        c, per_image, mode, p, seed, counter, call = args
        L.call("glis_dropout", L.ptr(src), L.ptr(dst), src.numel(), c, 1, per_image, mode, p, seed,
               C.c_void_p(counter.data_ptr()), call, L.stream())

    @staticmethod
    def backward(ctx, dout):
        dc = _nhwc(dout)
        dx = torch.empty_like(dc)
        DropoutFunction._run(dc, dx, ctx.args)
        return dx, None, None


def dropout(x, p, channel_mode=False):
    """Training-mode dropout of a CUDA fp32 tensor (2-D, or 4-D handled in NHWC storage order)."""
    if p <= 0:
        return x
    if p >= 1:
        raise RuntimeError("glis_b200: dropout probability must be < 1")
    return DropoutFunction.apply(x, p, channel_mode)
