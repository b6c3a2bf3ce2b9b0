"""CPU-side checks: the C ABI loads and exports everything the header declares, the host
mirror of the reference API has the reference's names / shapes / keys, and the product
refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from glis_b200 import _lib
    header = open(os.path.join(ROOT, "include", "glis_b200.h")).read()
    declared = set(re.findall(r"\b(glis_[a-z0-9_]+)\s*\(", header))
    declared -= {"glis_geom", "glis_epilogue"}
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python __graft_entry__.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.SIGNATURES) | {"glis_last_error", "glis_version", "glis_set_pdl", "glis_set_reserved_sms"}
    assert _lib.load().glis_version() >= 100


def test_struct_layout_matches_header():
    from glis_b200 import _lib
    assert ctypes.sizeof(_lib.Geom) == 16 * 4
    assert ctypes.sizeof(_lib.Epilogue) == 8 * 8


def test_error_convention_bad_args_no_gpu_needed():
    from glis_b200 import _lib
    lib = _lib.load()
    rc = lib.glis_rmsprop(None, None, None, 10, 0.1, 0.9, 1e-6, 1.0, None)
    assert rc == -1 and b"NULL" in lib.glis_last_error()
    g = _lib.Geom()
    rc = lib.glis_conv_forward(ctypes.byref(g), None, None, None, None, 0, None)
    assert rc == -1 and b"relation" in lib.glis_last_error() or b"extent" in lib.glis_last_error()


def test_product_keys_match_reference_catalogue(golden_dir):
    from common.model import (GeneratorLearnedInputSpace, build_discriminator, build_generator,
                              build_reverser)
    z = np.load(os.path.join(golden_dir, "models.npz"))
    builders = {
        "D_16": lambda: build_discriminator(16, 16, 4, 2, "weight", 0),
        "D_20x12_pad": lambda: build_discriminator(20, 12, 4, 3, "weight", 0),
        "R_16": lambda: build_reverser(16, 16, 4, 2, 8, "weight", 0),
        "G_16": lambda: build_generator(16, 16, 4, 2, 8, "weight"),
        "G_20x12_pad": lambda: build_generator(20, 12, 4, 3, 8, "weight"),
        "GLIS_16_k2of3": lambda: GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional"),
        "GLIS_16_nearest": lambda: GeneratorLearnedInputSpace(16, 16, 4, 3, 8, "weight", 1, "nearest"),
        "D_16_affine": lambda: build_discriminator(16, 16, 4, 2, "weight-affine", 0),
    }
    for case, build in builders.items():
        want = {k[len(case) + 3:]: z[k].shape for k in z.files if k.startswith(case + "/p.")}
        sd = build().state_dict()
        assert sorted(sd) == sorted(want), case
        for k in sd:
            assert tuple(sd[k].shape) == want[k], (case, k)


def test_state_dict_roundtrip_with_dotted_keys():
    from common.model import build_discriminator
    a, b = build_discriminator(16, 16, 4, 2, "weight", 0), build_discriminator(16, 16, 4, 2, "weight", 0)
    b.load_state_dict(a.state_dict())
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert "level.0.conv.weight" in a.state_dict()


def test_module_api_surface():
    from common.modules import (TPReLU, View, WeightNormalizedConv2d, WeightNormalizedConvTranspose2d,
                                WeightNormalizedLinear)
    c = WeightNormalizedConv2d(3, 8, 4, 2, 1)
    assert c.weight.shape == (8, 3, 4, 4) and c.scale.shape == (1, 8, 1, 1) and c.bias.shape == (1, 8, 1, 1)
    assert c.weight_norm_factor == 1.0 and not c.transposed
    t = WeightNormalizedConvTranspose2d(8, 3, 4, 2, 1, scale=False, bias=False)
    assert t.weight.shape == (8, 3, 4, 4) and t.scale is None and t.bias is None
    assert t.weight_norm_factor == 0.25 and t.transposed
    l = WeightNormalizedLinear(5, 7, init_factor=0.01)
    assert l.weight.shape == (7, 5) and l.scale.shape == (1, 7) and l.bias.shape == (1, 7)
    assert float(l.weight.abs().max()) <= 0.01 / 5 ** 0.5 + 1e-9
    p = TPReLU(6)
    assert p.num_parameters == 6 and float(p.weight[0]) == 0.25 and float(p.bias.abs().sum()) == 0
    v = View(2, 3)
    assert v(torch.arange(12.).view(2, 6)).shape == (2, 2, 3)
    assert "WeightNormalizedLinear (5 -> 7)" == repr(l)


def test_builder_errors_match_reference():
    from common.model import GeneratorLearnedInputSpace, build_discriminator
    with pytest.raises(ValueError):
        build_discriminator(15, 16, 4, 2, "weight", 0)
    with pytest.raises(Exception):
        GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 1, "cubic")
    with pytest.raises(NotImplementedError):
        build_discriminator(16, 16, 4, 2, "batch", 0)


def test_no_cpu_fallback():
    from common.modules import TPReLU, WeightNormalizedLinear
    with pytest.raises(RuntimeError):
        WeightNormalizedLinear(4, 4)(torch.zeros(2, 4))
    with pytest.raises(RuntimeError):
        TPReLU(4)(torch.zeros(2, 4))


def test_lis_depth_rule_product():
    from common.model import GeneratorLearnedInputSpace

    class Seq(object):
        def __init__(self, vals):
            self.vals, self.n = list(vals), 0

        def random(self):
            self.n += 1
            return self.vals.pop(0)

    g = GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional")
    g.rng = Seq([0.2, 0.2, 0.6]); assert g.lis_depth() == 1 and g.rng.n == 2
    g.rng = Seq([0.9, 0.9, 0.9]); assert g.lis_depth() == 3
    g.eval(); g.rng = Seq([0.0, 0.0, 0.0]); assert g.lis_depth() == 3
    g.train(); g.rng = Seq([0.5] * 3); assert g.lis_depth(2) == 2 and g.rng.n == 3


def test_tensor_core_planning_queries():
    """Host-side planning (no GPU needed): which launches split K, which fused forwards run as split-K sums +
    a pointwise pass, which activations may stay planes-only, which code sizes the LIS kernel takes."""
    import ctypes as C
    import torch
    from glis_b200 import _lib as L, ops
    lib = L.load()
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    geom = lambda n, hi, ci, ho, co: spec.geom(L.CONV, n, hi, hi, ci, ho, ho, co)
    d3_b, d3_2b = geom(64, 10, 256, 5, 512), geom(128, 10, 256, 5, 512)
    d1_2b, d2_b = geom(128, 40, 64, 20, 128), geom(64, 20, 128, 10, 256)
    assert lib.glis_conv_tc_ksplit(C.byref(d3_b)) >= 2 and ops._split_k_forward(d3_b, L.CONV)       # 64 k-steps
    assert lib.glis_conv_tc_ksplit(C.byref(d3_2b)) >= 2 and ops._split_k_forward(d3_2b, L.CONV)
    assert lib.glis_conv_tc_ksplit(C.byref(d1_2b)) == 1 and not ops._split_k_forward(d1_2b, L.CONV)
    assert not ops._split_k_forward(d2_b, L.CONV)                     # a 2-way split over 32 k-steps does not pay
    # the generator head's data gradient: 200 blocks of 64 features into 256 columns -> a deep split
    head = ops.ContractionSpec(False, (1, 1), (1, 1), (0, 0), (1, 1), linear=True, perm=(512, 25))
    g = head.geom(L.TCONV, 64, 1, 1, 12800, 1, 1, 256)
    assert lib.glis_conv_tc_supported(C.byref(g)) == 1 and lib.glis_conv_tc_ksplit(C.byref(g)) >= 16
    assert lib.glis_lis_supported(256) == 1 and lib.glis_lis_supported(128) == 1
    assert lib.glis_lis_supported(512) == 0 and lib.glis_lis_supported(48) == 0
    # planes-only activations: a tensor-core conv consumer (forward + weight gradient on tcgen05) suffices,
    # D's 5x5 head (one output channel: fp32 kernel) does not, nothing does in fp32 mode
    w_conv, w_head = torch.zeros(256, 128, 4, 4), torch.zeros(1, 512, 5, 5)
    head_spec = ops.ContractionSpec(False, (5, 5), (1, 1), (0, 0), (1, 1))
    assert ops.consumes_planes_only(spec, w_conv, (64, 128, 20, 20), True)
    assert not ops.consumes_planes_only(head_spec, w_head, (64, 512, 5, 5), False)
    assert not ops.consumes_planes_only(ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1), precision=L.PREC_FP32),
                                        w_conv, (64, 128, 20, 20), True)
    # G's last layer (64 -> 3, transposed, followed by the sigmoid): the "fold" product reads planes
    tspec = ops.ContractionSpec(True, (4, 4), (2, 2), (1, 1), (1, 1))
    assert ops.consumes_planes_only(tspec, torch.zeros(64, 3, 4, 4), (64, 64, 40, 40), False)
    assert not ops.consumes_planes_only(tspec, torch.zeros(64, 3, 4, 4), (64, 64, 40, 40), True)


def test_tc_conv_plan_invariants():
    """Property test of the tcgen05 launch planner (host code, no GPU): for every geometry the kernel claims to
    support, the plan tiles all pixels and channels, fits TMEM and shared memory, keeps at least two pipeline
    stages, and only splits K where the epilogue allows it."""
    import ctypes as C
    from hypothesis import given, settings, strategies as st
    from glis_b200 import _lib as L, ops
    lib = L.load()
    cdiv = lambda a, b: -(-a // b)

    def check(relation, kernel, stride, pad, n, h, w, ci, co, plain):
        spec = ops.ContractionSpec(relation == L.TCONV, (kernel, kernel), (stride, stride), (pad, pad), (1, 1))
        if relation == L.CONV:
            ho, wo = spec.__class__(False, (kernel, kernel), (stride, stride), (pad, pad), (1, 1)).out_hw(h, w)
            if ho < 1 or wo < 1:
                return
            g = spec.geom(L.CONV, n, h, w, ci, ho, wo, co)
        else:
            ho, wo = spec.out_hw(h, w)
            g = spec.geom(L.TCONV, n, h, w, ci, ho, wo, co)
        if not lib.glis_conv_tc_supported(C.byref(g)):
            return
        out = (C.c_int * 15)()
        assert lib.glis_conv_tc_plan(C.byref(g), int(plain), out) == 0, lib.glis_last_error()
        (tw, th, tn, n_mma, tmem, kblocks, ksplit, a_rows, stages, tiles_h, tiles_x, tiles_co, total, groups,
         smem) = list(out)
        nphase = stride * stride if relation == L.TCONV else 1
        hq, wq = (cdiv(ho, stride), cdiv(wo, stride)) if relation == L.TCONV else (ho, wo)
        assert tw == wq and 1 <= th <= hq and 1 <= tn <= n and (tn == 1 or th == hq)
        assert tw * th * tn <= n_mma <= 256 and n_mma % 16 == 0 and n_mma - tw * th * tn < 16
        assert tiles_h == cdiv(hq, th) and tiles_x >= tiles_h * cdiv(n, tn) and tiles_co == cdiv(co, 128)
        assert total == tiles_x * tiles_co * nphase and groups == total * ksplit
        assert tmem in (64, 128, 256, 512) and tmem >= 2 * n_mma
        assert kblocks == cdiv(ci, 64) and 1 <= ksplit <= min(32, kblocks) and (plain or ksplit == 1)
        assert a_rows in (64, 128) and (a_rows == 128 or (co <= 64 and n_mma >= 64))
        stage = 2 * a_rows * 128 + 2 * n_mma * 128
        assert 2 <= stages <= 4 and smem == stages * stage + 2304 and smem <= 227 * 1024
        assert a_rows * 128 + 128 * 128 <= stage          # the 128-row MMA read that starts in the lo weight tile

    geometry = st.tuples(st.sampled_from([(L.CONV, 4, 2, 1), (L.TCONV, 4, 2, 1), (L.CONV, 3, 1, 1), (L.CONV, 1, 1, 0),
                                          (L.CONV, 5, 1, 0)]),
                         st.integers(1, 130), st.integers(1, 100).map(lambda v: 2 * v), st.integers(1, 100).map(lambda v: 2 * v),
                         st.sampled_from([32, 40, 48, 64, 72, 96, 128, 200, 256, 512, 1024]),
                         st.sampled_from([32, 33, 40, 48, 64, 100, 128, 129, 192, 256, 512]), st.booleans())

    @settings(max_examples=600, deadline=None)
    @given(geometry)
    def run(case):
        (relation, kernel, stride, pad), n, h, w, ci, co, plain = case
        check(relation, kernel, stride, pad, n, h, w, ci, co, plain)

    run()
    # the layers of configs 2 and 4 explicitly
    for (rel, n, h, ci, co) in ((L.CONV, 128, 40, 64, 128), (L.CONV, 64, 10, 256, 512), (L.TCONV, 64, 5, 512, 256),
                                (L.TCONV, 64, 20, 128, 64), (L.CONV, 32, 80, 64, 128), (L.TCONV, 32, 40, 128, 64)):
        check(rel, 4, 2, 1, n, h, h, ci, co, True)
        check(rel, 4, 2, 1, n, h, h, ci, co, False)


def test_pair_plan_invariants(monkeypatch):
    """The cta_group::2 planner (host code, no GPU): off unless GLIS_TC_PAIR=1; when it takes a launch the pair's tile
    splits into two TMA-expressible halves (rows or images), each padded to 8 B rows, UMMA N = 2 x that <= 256, the
    accumulators fit TMEM twice and at least two stages fit shared memory."""
    import ctypes as C
    from glis_b200 import _lib as L, ops
    lib = L.load()
    cdiv = lambda a, b: -(-a // b)
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    cases = [(L.CONV, 128, 20, 128, 256), (L.CONV, 128, 10, 256, 512), (L.CONV, 64, 10, 256, 512), (L.TCONV, 64, 5, 512, 256),
             (L.CONV, 7, 10, 64, 256), (L.CONV, 3, 20, 128, 512)]
    out = (C.c_int * 16)()
    g0 = spec.geom(L.CONV, 128, 20, 20, 128, 10, 10, 256)
    monkeypatch.delenv("GLIS_TC_PAIR", raising=False)
    assert lib.glis_conv_tc_pair_plan(C.byref(g0), 0, out) == -2          # opt-in
    monkeypatch.setenv("GLIS_TC_PAIR", "1")
    for rel, n, h, ci, co in cases:
        for plain in (0, 1):
            if rel == L.CONV:
                g = spec.geom(L.CONV, n, h, h, ci, h // 2, h // 2, co)
                hq, wq, nphase = h // 2, h // 2, 1
            else:
                g = spec.geom(L.TCONV, n, h, h, ci, 2 * h, 2 * h, co)
                hq, wq, nphase = h, h, 4
            assert lib.glis_conv_tc_pair_plan(C.byref(g), plain, out) == 0, lib.glis_last_error()
            tw, th, tn, hh, hn, n_half, n_mma, tmem, kblocks, ksplit, stages, tiles_h, tiles_x, pairs, total, groups = list(out)
            assert tw == wq and ((hn == tn == 1 and 2 * hh == th) or (hh == th == hq and 2 * hn == tn))
            assert n_half == tw * hh * hn and n_mma % 16 == 0 and n_half <= n_mma // 2 < n_half + 8 and n_mma <= 256
            assert tmem in (64, 128, 256, 512) and tmem >= 2 * n_mma
            assert kblocks == ci // 64 and 1 <= ksplit <= min(32, kblocks) and (plain or ksplit == 1)
            assert 2 <= stages <= 4 and stages * (2 * 128 * 128 + n_mma * 128) <= 219 * 1024
            assert pairs == co // 256 and tiles_h == cdiv(hq, th) and tiles_x == tiles_h * cdiv(n, tn)
            assert total == tiles_x * pairs * nphase and groups == total * ksplit
    # not for < 256 output channels, nor where the halo kernel runs (maps >= 16 wide)
    assert lib.glis_conv_tc_pair_plan(C.byref(spec.geom(L.CONV, 64, 20, 20, 128, 10, 10, 128)), 0, out) == -2
    assert lib.glis_conv_tc_pair_plan(C.byref(spec.geom(L.CONV, 64, 80, 80, 64, 40, 40, 256)), 0, out) == -2


def test_process_wide_switches_and_reserved_sms():
    """glis_set_pdl / glis_set_reserved_sms return the previous setting; reserving SMs changes launch plans only."""
    import ctypes as C
    from glis_b200 import _lib as L, ops
    lib = L.load()
    prev = lib.glis_set_pdl(2)
    assert lib.glis_set_pdl(1) == 2 and lib.glis_set_pdl(7) == 1 and lib.glis_set_pdl(prev) == 2
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    g = spec.geom(L.TCONV, 64, 5, 5, 512, 10, 10, 256)            # a split-K data gradient: the split follows the SM count
    out = (C.c_int * 15)()
    r0 = lib.glis_set_reserved_sms(0)
    try:
        assert lib.glis_conv_tc_plan(C.byref(g), 1, out) == 0
        full = list(out)
        assert lib.glis_set_reserved_sms(100) == 0
        assert lib.glis_conv_tc_plan(C.byref(g), 1, out) == 0
        small = list(out)
        assert small[13] <= full[13] and small[:3] != full[:3] or small[6] <= full[6]     # fewer work items / a smaller split
        assert lib.glis_set_reserved_sms(-5) == 100 and lib.glis_set_reserved_sms(0) == 0
    finally:
        lib.glis_set_reserved_sms(r0)


def test_peer_allreduce_argument_checks():
    """glis_peer_allreduce validates before it launches (no GPU needed): rank / world, slice alignment, block count."""
    import ctypes as C
    from glis_b200 import _lib as L
    lib = L.load()
    two = (C.c_void_p * 2)(C.c_void_p(4096), C.c_void_p(8192))
    e = C.c_void_p(4096)
    call = lambda *a: lib.glis_peer_allreduce(*a)
    assert call(None, two, 0, 2, 0, 64, e, 8, None) == -1
    assert call(two, two, 2, 2, 0, 64, e, 8, None) == -1 and b"rank" in lib.glis_last_error()
    assert call(two, two, 0, 9, 0, 72, e, 8, None) == -1
    assert call(two, two, 0, 2, 2, 64, e, 8, None) == -1 and b"multiple of 4" in lib.glis_last_error()
    assert call(two, two, 0, 2, 0, 60, e, 8, None) == -1
    assert call(two, two, 0, 2, 0, 64, e, 0, None) == -1 and call(two, two, 0, 2, 0, 64, e, 1000, None) == -1
    assert call(two, two, 0, 2, 0, 0, e, 8, None) == 0              # an empty slice is a no-op
    three = (C.c_void_p * 3)(C.c_void_p(4096), C.c_void_p(8192), C.c_void_p(12288))
    assert call(three, three, 0, 3, 0, 96, e, 8, None) == -2        # 2, 4 or 8 ranks
