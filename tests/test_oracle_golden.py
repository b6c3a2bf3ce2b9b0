"""The oracle against golden vectors produced by the reference's own sources
(tests/golden/make_golden.py).  fp64, so agreement is to rounding (1e-11)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.step import GLISOracleTrainer

TOL = dict(rtol=1e-10, atol=1e-12)


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: z[k] for k in z.files}


def _group(store, prefix):
    pre = prefix + "/"
    return {k[len(pre):]: v for k, v in store.items() if k.startswith(pre)}


def _t(a):
    return torch.from_numpy(np.array(a)).double()


MODULE_CASES = {
    "conv_s2": lambda: oracle.WeightNormalizedConv2d(3, 5, 4, 2, (1, 1), scale=False, bias=False),
    "conv_s2_pad2_affine": lambda: oracle.WeightNormalizedConv2d(4, 6, 4, 2, (2, 1)),
    "conv_head": lambda: oracle.WeightNormalizedConv2d(6, 1, (3, 5)),
    "conv_3x3_s1": lambda: oracle.WeightNormalizedConv2d(4, 3, 3, 1, (1, 1), scale=False, bias=False),
    "deconv_s2": lambda: oracle.WeightNormalizedConvTranspose2d(6, 4, 4, 2, (1, 1), scale=False, bias=False),
    "deconv_s2_pad2_affine": lambda: oracle.WeightNormalizedConvTranspose2d(4, 3, 4, 2, (2, 1)),
    "linear_plain": lambda: oracle.WeightNormalizedLinear(7, 10, scale=False, bias=False, init_factor=0.01),
    "linear_affine": lambda: oracle.WeightNormalizedLinear(7, 5),
    "tprelu_2d": lambda: oracle.TPReLU(6),
    "tprelu_4d": lambda: oracle.TPReLU(3),
}


@pytest.mark.parametrize("case", sorted(MODULE_CASES))
def test_modules_match_reference(golden_dir, case):
    g = _group(_load(golden_dir, "modules.npz"), case)
    m = MODULE_CASES[case]().double()
    names = [n for n, _ in m.named_parameters()]
    assert sorted(names) == sorted(k[2:] for k in g if k.startswith("p."))
    with torch.no_grad():
        for n, p in m.named_parameters():
            assert tuple(p.shape) == g["p." + n].shape
            p.copy_(_t(g["p." + n]))
    x = _t(g["x"]).requires_grad_(True)
    y = m(x)
    np.testing.assert_allclose(y.detach().numpy(), g["y"], **TOL)
    (y * _t(g["r"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g["dx"], **TOL)
    for n, p in m.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), g["g." + n], err_msg=n, **TOL)


MODEL_CASES = {
    "D_16": (lambda: oracle.build_discriminator(16, 16, 4, 2, "weight", 0), None),
    "D_20x12_pad": (lambda: oracle.build_discriminator(20, 12, 4, 3, "weight", 0), None),
    "R_16": (lambda: oracle.build_reverser(16, 16, 4, 2, 8, "weight", 0), None),
    "G_16": (lambda: oracle.build_generator(16, 16, 4, 2, 8, "weight"), None),
    "G_20x12_pad": (lambda: oracle.build_generator(20, 12, 4, 3, 8, "weight"), None),
    "GLIS_16_k2of3": (lambda: oracle.GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional"), 2),
    "GLIS_16_nearest": (lambda: oracle.GeneratorLearnedInputSpace(16, 16, 4, 3, 8, "weight", 1, "nearest"), "all"),
    "D_16_affine": (lambda: oracle.build_discriminator(16, 16, 4, 2, "weight-affine", 0), None),
}


@pytest.mark.parametrize("case", sorted(MODEL_CASES))
def test_models_match_reference(golden_dir, case):
    g = _group(_load(golden_dir, "models.npz"), case)
    build, depth = MODEL_CASES[case]
    net = build().double().eval()
    want = {k[2:]: v for k, v in g.items() if k.startswith("p.")}
    sd = net.state_dict()
    # key catalogue and shapes are exactly the reference's
    assert sorted(sd) == sorted(want)
    for k in sd:
        assert tuple(sd[k].shape) == want[k].shape, k
    net.load_state_dict({k: _t(v) for k, v in want.items()})
    x = _t(g["x"]).requires_grad_(True)
    out = net(x, n_execute_lis_layers=depth) if depth is not None else net(x)
    flat = [out[0]] + list(out[1]) if isinstance(out, tuple) else [out]
    assert len(flat) == int(g["n_out"])
    loss = 0
    for i, o in enumerate(flat):
        np.testing.assert_allclose(o.detach().numpy(), g["y%d" % i], **TOL)
        loss = loss + (o * _t(g["r%d" % i])).sum()
    loss.backward()
    np.testing.assert_allclose(x.grad.numpy(), g["dx"], **TOL)
    from oracle.model import _DOT
    for n, p in net.named_parameters():
        ref = g["g." + n.replace(_DOT, ".")]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(ref)
        np.testing.assert_allclose(got, ref, err_msg=n, **TOL)


@pytest.mark.parametrize("fixture,ls", [("glis_steps.npz", False), ("glis_steps_ls.npz", True)])
def test_glis_iterations_match_reference(golden_dir, fixture, ls):
    s = _load(golden_dir, fixture)
    cfg = _group(s, "cfg")
    W, H, B, code, nf, nl, n_lis = (int(cfg[k]) for k in ("W", "H", "B", "code", "nf", "nl", "n_lis"))
    gen = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, "fractional").double()
    dis = oracle.build_discriminator(W, H, nf, nl, "weight", 0).double()
    gen.load_state_dict({k: _t(v) for k, v in _group(s, "init/g").items()})
    dis.load_state_dict({k: _t(v) for k, v in _group(s, "init/d").items()})
    tr = GLISOracleTrainer(gen, dis, lr=float(cfg["lr"]), lambda_r=float(cfg["lam"]), ls=ls)
    for it, (kd, kg) in enumerate(cfg["depths"]):
        g = _group(s, "it%d" % it)
        out = tr.step(_t(g["real"]), _t(g["zd"]), _t(g["zg"]), depth_d=int(kd), depth_g=int(kg))
        assert out["depth_d"] == kd and out["depth_g"] == kg
        np.testing.assert_allclose(out["d_real"], g["d_real"], **TOL)
        np.testing.assert_allclose(out["d_fake"], g["d_fake"], **TOL)
        np.testing.assert_allclose(out["g"], g["g"], **TOL)
        np.testing.assert_allclose(np.array(out["r"]), g["r"], **TOL)
        for k, v in gen.state_dict().items():
            np.testing.assert_allclose(v.numpy(), g["g/" + k], err_msg="it%d gen %s" % (it, k), **TOL)
        for k, v in dis.state_dict().items():
            np.testing.assert_allclose(v.numpy(), g["d/" + k], err_msg="it%d dis %s" % (it, k), **TOL)


def test_lis_depth_rule():
    """Break probabilities (1/2)**(n-i) in training, 0 in eval, forced by n_execute (model.py:281-297)."""
    class Seq:
        def __init__(self, vals): self.vals, self.n = list(vals), 0
        def random(self):
            self.n += 1
            return self.vals.pop(0)
    g = oracle.GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional")
    g.rng = Seq([0.2, 0.2, 0.6]); assert g.lis_depth() == 1 and g.rng.n == 2   # 0.2>=1/8, 0.2<1/4 -> break at i=1
    g.rng = Seq([0.9, 0.9, 0.9]); assert g.lis_depth() == 3 and g.rng.n == 3
    g.rng = Seq([0.1]); assert g.lis_depth() == 0 and g.rng.n == 1
    g.eval(); g.rng = Seq([0.0, 0.0, 0.0]); assert g.lis_depth() == 3
    g.train(); g.rng = Seq([0.5, 0.5, 0.5]); assert g.lis_depth(2) == 2 and g.rng.n == 3
    g.rng = Seq([0.0, 0.0, 0.0]); assert g.lis_depth("all") == 3


def test_builder_errors():
    with pytest.raises(ValueError):
        oracle.build_discriminator(15, 16, 4, 2, "weight", 0)
    with pytest.raises(ValueError):
        oracle.GeneratorLearnedInputSpace(16, 15, 4, 2, 8, "weight", 1, "fractional")
    with pytest.raises(Exception):
        oracle.GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 1, "cubic")


def test_r_iterative_iterations_match_reference(golden_dir):
    """oracle.riter_iteration against two outer iterations of r_iterative/main.py:428-535 run on the reference's
    own builders with three stock RMSprops (all hops trained; then [skip, train, train])."""
    from oracle.step import riter_iteration
    s = _load(golden_dir, "riter_steps.npz")
    cfg = _group(s, "cfg")
    W, H, B, code, nf, nl, R = (int(cfg[k]) for k in ("W", "H", "B", "code", "nf", "nl", "R"))
    gen = oracle.build_generator(W, H, nf, nl, code, "weight").double()
    rev = oracle.build_reverser(W, H, nf // 2, nl, code, "weight", 0).double()
    dis = oracle.build_discriminator(W, H, nf, nl, "weight", 0).double()
    for tag, net in (("g", gen), ("r", rev), ("d", dis)):
        net.load_state_dict({k: _t(v) for k, v in _group(s, "init/" + tag).items()})
    gs, rs, ds = {}, {}, {}
    for it, flags in enumerate(cfg["schedules"]):
        flags = [bool(f) for f in flags]
        g = _group(s, "it%d" % it)
        reals = [_t(g["real%d" % i]) for i in range(sum(flags))]
        out = riter_iteration(gen, rev, dis, gs, rs, ds, _t(g["z"]), reals, float(cfg["lr"]), float(cfg["lam"]), R, flags)
        for hop, rec in enumerate(out):
            assert (rec is None) == (not flags[hop])
            if rec is None:
                continue
            want = _group(s, "it%d/hop%d" % (it, hop))
            assert sorted(rec) == sorted(want)
            for k in rec:
                np.testing.assert_allclose(rec[k], want[k], err_msg="it%d hop%d %s" % (it, hop, k), **TOL)
        for tag, net in (("g", gen), ("r", rev), ("d", dis)):
            for k, v in net.state_dict().items():
                np.testing.assert_allclose(v.numpy(), g[tag + "/" + k], err_msg="it%d %s %s" % (it, tag, k), **TOL)


def test_r_separate_iterations_match_reference(golden_dir):
    """oracle.rsep_iteration against two iterations of g_lis/train_r.py:406-436 run on the reference's builders."""
    from oracle.step import rsep_iteration
    s = _load(golden_dir, "rsep_steps.npz")
    cfg = _group(s, "cfg")
    W, H, B, code, nf, nl, n_lis = (int(cfg[k]) for k in ("W", "H", "B", "code", "nf", "nl", "n_lis"))
    gen = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, "fractional").double()
    rev = oracle.build_reverser(W, H, nf // 2, nl, code, "weight", 0).double()
    dis = oracle.build_discriminator(W, H, nf, nl, "weight", 0).double()
    for tag, net in (("g", gen), ("r", rev), ("d", dis)):
        net.load_state_dict({k: _t(v) for k, v in _group(s, "init/" + tag).items()})
    state = {}
    for it in range(int(cfg["iters"])):
        g = _group(s, "it%d" % it)
        out = rsep_iteration(gen, rev, dis, state, _t(g["z"]), float(cfg["lr"]), n_lis)
        for k in ("stage1", "r", "stage2"):
            np.testing.assert_allclose(out[k], g[k], err_msg="it%d %s" % (it, k), **TOL)
        for k, v in rev.state_dict().items():
            np.testing.assert_allclose(v.numpy(), g["r/" + k], err_msg="it%d %s" % (it, k), **TOL)
