"""GPU parity: the sm_100a path (through the ctypes C ABI) against the oracle on identical
weights and inputs, and against the golden vectors produced by the reference's sources.

Tolerances (north star): relative error = max|got-want| / max|want|.
  fp32 contraction mode :  forward / losses <= 1e-4,  gradients <= 1e-3
  (the fp32 kernels land around 1e-6 / 1e-5; the bounds are the contract, not the typical value)
"""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.step import GLISOracleTrainer
from util import FlipAwarePair, copy_params, randomize_params_, rel_err, rel_l2

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL = 1e-4, 1e-3
# A TPReLU branch mask may differ from the fp64 oracle's only where the oracle's pre-activation lies within the
# forward tolerance of the kink (oracle/flipaware.py); everywhere else a differing bit is an error.
FLIP_TOL = FWD_TOL
DEV = "cuda"


@pytest.fixture(params=["fp32", "bf16x3"], autouse=True)
def precision(request):
    """Every parity test runs twice: fp32 FFMA contractions everywhere, and the tcgen05 split-bf16
    (fp32-faithful) mode on the layers that tile for it.  Same tolerances in both."""
    from glis_b200 import _lib
    _lib.set_precision(request.param)
    yield request.param
    _lib.set_precision("bf16x3")


def _product():
    import common.model as pm
    import common.modules as pmod
    return pm, pmod


def _check_module(ref, prod, x, fwd_tol=FWD_TOL, grad_tol=GRAD_TOL, seed=0):
    """ref: oracle module (CPU, fp64), prod: product module (CUDA, fp32) with equal params."""
    gen = torch.Generator().manual_seed(seed)
    xr = x.double().clone().requires_grad_(True)
    xp = x.float().to(DEV).requires_grad_(True)
    yr, yp = ref(xr), prod(xp)
    assert tuple(yr.shape) == tuple(yp.shape)
    assert rel_err(yp, yr) <= fwd_tol, "forward %g" % rel_err(yp, yr)
    r = torch.rand(yr.shape, generator=gen, dtype=torch.float64) * 2 - 1
    (yr * r).sum().backward()
    (yp * r.float().to(DEV)).sum().backward()
    assert rel_err(xp.grad, xr.grad) <= grad_tol, "dx %g" % rel_err(xp.grad, xr.grad)
    for (n, pr), (_, pp) in zip(ref.named_parameters(), prod.named_parameters()):
        e = rel_err(pp.grad, pr.grad)
        assert e <= grad_tol, "grad %s %g" % (n, e)


def _pair(make_ref, make_prod, seed=1):
    gen = torch.Generator().manual_seed(seed)
    ref = make_ref()
    randomize_params_(ref, gen)
    prod = make_prod()
    copy_params(prod, ref)
    return ref.double(), prod.to(DEV)


CONV_CASES = [
    # (Cin, Cout, k, s, p, scale/bias, N, H, W)
    (3, 5, 4, 2, (1, 1), False, 2, 8, 12),
    (4, 6, 4, 2, (2, 1), True, 2, 10, 8),
    (6, 1, (3, 5), 1, 0, True, 3, 3, 5),
    (4, 3, 3, 1, (1, 1), False, 2, 6, 6),
    (3, 64, 4, 2, 1, False, 3, 32, 32),       # D level 0 shape family (Cin = 3)
    (64, 128, 4, 2, 1, False, 2, 20, 20),     # D level 1 family
    (70, 33, 4, 2, (2, 2), True, 2, 10, 14),  # ragged channel counts, pad 2
    (128, 1, (5, 5), 1, 0, True, 5, 5, 5),    # D head
    (512, 1, (5, 5), 1, 0, True, 128, 5, 5),  # D head at config 2, 2B-image batch (K = 12800)
    (32, 16, (4, 4), 1, 0, True, 4, 4, 4),    # R head family (Cout = code)
    (64, 128, 4, 2, 1, False, 3, 40, 40),     # tcgen05: multi-row pixel tiles (3 rows of 40 -> ragged last tile)
    (128, 256, 4, 2, 1, True, 5, 20, 20),     # tcgen05: two channel tiles, whole 10x10 image per tile
    (256, 192, 4, 2, 1, False, 7, 10, 10),    # tcgen05: 5x5 images, 5 per tile, ragged batch, Cout % 128 != 0
    (64, 64, 4, 2, (2, 2), False, 2, 12, 20), # tcgen05: pad 2, non-square, Cout = 64
    (128, 64, 3, 1, 1, False, 2, 12, 12),     # tcgen05: 3x3 stride 1 (nearest-upsampling generator variant)
    (128, 256, 4, 2, 1, False, 32, 8, 8),     # config-1 D level 2: 4x4 output, four images per K tile
    (64, 128, 4, 2, 1, False, 32, 16, 16),    # config-1 D level 1
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_wn_conv2d(case):
    _, pmod = _product()
    ci, co, k, s, p, aff, n, h, w = case
    ref, prod = _pair(lambda: oracle.WeightNormalizedConv2d(ci, co, k, s, p, scale=aff, bias=aff),
                      lambda: pmod.WeightNormalizedConv2d(ci, co, k, s, p, scale=aff, bias=aff))
    x = torch.rand(n, ci, h, w, generator=torch.Generator().manual_seed(3)) * 2 - 1
    _check_module(ref, prod, x)


DECONV_CASES = [
    (6, 4, 4, 2, (1, 1), False, 2, 5, 3),
    (4, 3, 4, 2, (2, 1), True, 2, 5, 4),
    (64, 3, 4, 2, 1, True, 2, 16, 16),        # G level 0 family (Cout = 3)
    (128, 64, 4, 2, 1, False, 2, 10, 10),
    (40, 24, 4, 2, (2, 2), False, 3, 6, 9),
    (16, 8, 3, 1, 1, True, 2, 7, 7),          # stride 1 transposed
    (512, 256, 4, 2, 1, False, 7, 5, 5),      # tcgen05: G level 3 family, 5 images per tile
    (256, 128, 4, 2, 1, True, 3, 10, 10),     # tcgen05: G level 2 family
    (128, 64, 4, 2, 1, False, 2, 20, 20),     # tcgen05: G level 1 family (Cout = 64)
    (64, 128, 4, 2, (2, 2), False, 2, 6, 10), # tcgen05: pad 2, non-square
    (256, 128, 4, 2, 1, False, 32, 4, 4),     # config-1 G level 2
    (128, 64, 4, 2, 1, False, 32, 8, 8),      # config-1 G level 1
]


@pytest.mark.parametrize("case", DECONV_CASES)
def test_wn_conv_transpose2d(case):
    _, pmod = _product()
    ci, co, k, s, p, aff, n, h, w = case
    ref, prod = _pair(lambda: oracle.WeightNormalizedConvTranspose2d(ci, co, k, s, p, scale=aff, bias=aff),
                      lambda: pmod.WeightNormalizedConvTranspose2d(ci, co, k, s, p, scale=aff, bias=aff))
    x = torch.rand(n, ci, h, w, generator=torch.Generator().manual_seed(4)) * 2 - 1
    _check_module(ref, prod, x)


def test_conv_transpose_output_size():
    _, pmod = _product()
    ref, prod = _pair(lambda: oracle.WeightNormalizedConvTranspose2d(4, 3, 3, 2, 1),
                      lambda: pmod.WeightNormalizedConvTranspose2d(4, 3, 3, 2, 1))
    x = torch.rand(2, 4, 5, 5) - 0.5
    yr = ref(x.double(), output_size=(10, 10))
    yp = prod(x.to(DEV), output_size=(10, 10))
    assert yp.shape == (2, 3, 10, 10) and rel_err(yp, yr) <= FWD_TOL
    with pytest.raises(ValueError):
        prod(x.to(DEV), output_size=(12, 12))


@pytest.mark.parametrize("case", [(7, 10, False, 4), (7, 5, True, 4), (256, 256, False, 64), (64, 800, False, 32),
                                  (256, 12800, False, 64),    # G's initial linear at config 2 (K cluster of 4)
                                  (300, 70, True, 130),       # ragged: 3 row tiles, K = 300 -> two clusters (atomics + bias)
                                  (1100, 33, True, 9)])       # K deeper than one cluster of 8 chunks
def test_wn_linear(case):
    _, pmod = _product()
    i, o, aff, b = case
    ref, prod = _pair(lambda: oracle.WeightNormalizedLinear(i, o, scale=aff, bias=aff, init_factor=0.01),
                      lambda: pmod.WeightNormalizedLinear(i, o, scale=aff, bias=aff, init_factor=0.01))
    _check_module(ref, prod, torch.randn(b, i, generator=torch.Generator().manual_seed(5)))


@pytest.mark.parametrize("shape", [(5, 6), (2, 3, 4, 5), (64, 256), (4, 32, 10, 10)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_tprelu(shape, channels_last):
    _, pmod = _product()
    if channels_last and len(shape) != 4:
        pytest.skip("2-D input has one layout")
    ref, prod = _pair(lambda: oracle.TPReLU(shape[1]), lambda: pmod.TPReLU(shape[1]))
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(6))
    gen = torch.Generator().manual_seed(7)
    xr = x.double().requires_grad_(True)
    xp = x.to(DEV)
    if channels_last:
        xp = xp.contiguous(memory_format=torch.channels_last)
    xp.requires_grad_(True)
    yr, yp = ref(xr), prod(xp)
    assert rel_err(yp, yr) <= 1e-6
    r = torch.rand(yr.shape, generator=gen, dtype=torch.float64) - 0.5
    (yr * r).sum().backward()
    (yp * r.float().to(DEV)).sum().backward()
    assert rel_err(xp.grad, xr.grad) <= 1e-6
    assert rel_err(prod.weight.grad, ref.weight.grad) <= 1e-4
    assert rel_err(prod.bias.grad, ref.bias.grad) <= 1e-4


def test_weight_norm_and_norm_scale_bias_helpers():
    _, pmod = _product()
    for ref, prod in [
        _pair(lambda: oracle.WeightNormalizedConv2d(5, 7, 4, 2, 1), lambda: pmod.WeightNormalizedConv2d(5, 7, 4, 2, 1)),
        _pair(lambda: oracle.WeightNormalizedConvTranspose2d(5, 7, 4, 2, 1),
              lambda: pmod.WeightNormalizedConvTranspose2d(5, 7, 4, 2, 1)),
    ]:
        nr, np_ = ref.weight_norm(), prod.weight_norm()
        assert sorted(np_.shape) == sorted(nr.shape) and rel_err(np_.reshape(-1), nr.reshape(-1)) <= 1e-6
        z = torch.randn(2, 7, 3, 3)
        assert rel_err(prod.norm_scale_bias(z.to(DEV)), ref.norm_scale_bias(z.double())) <= 1e-6
    ref, prod = _pair(lambda: oracle.WeightNormalizedLinear(6, 9), lambda: pmod.WeightNormalizedLinear(6, 9))
    assert prod.weight_norm().shape == (9, 1)
    assert rel_err(prod.weight_norm(), ref.weight_norm()) <= 1e-6


# ---------------------------------------------------------------- golden vectors (reference sources)
def _golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: z[k] for k in z.files}


def _grp(store, prefix):
    pre = prefix + "/"
    return {k[len(pre):]: v for k, v in store.items() if k.startswith(pre)}


def test_golden_modules(golden_dir):
    _, pmod = _product()
    cases = {
        "conv_s2": lambda: pmod.WeightNormalizedConv2d(3, 5, 4, 2, (1, 1), scale=False, bias=False),
        "conv_s2_pad2_affine": lambda: pmod.WeightNormalizedConv2d(4, 6, 4, 2, (2, 1)),
        "conv_head": lambda: pmod.WeightNormalizedConv2d(6, 1, (3, 5)),
        "conv_3x3_s1": lambda: pmod.WeightNormalizedConv2d(4, 3, 3, 1, (1, 1), scale=False, bias=False),
        "deconv_s2": lambda: pmod.WeightNormalizedConvTranspose2d(6, 4, 4, 2, (1, 1), scale=False, bias=False),
        "deconv_s2_pad2_affine": lambda: pmod.WeightNormalizedConvTranspose2d(4, 3, 4, 2, (2, 1)),
        "linear_plain": lambda: pmod.WeightNormalizedLinear(7, 10, scale=False, bias=False, init_factor=0.01),
        "linear_affine": lambda: pmod.WeightNormalizedLinear(7, 5),
        "tprelu_2d": lambda: pmod.TPReLU(6),
        "tprelu_4d": lambda: pmod.TPReLU(3),
    }
    store = _golden(golden_dir, "modules.npz")
    for case, build in cases.items():
        g = _grp(store, case)
        m = build().to(DEV)
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.copy_(torch.from_numpy(g["p." + n]).float())
        x = torch.from_numpy(g["x"]).float().to(DEV).requires_grad_(True)
        y = m(x)
        assert rel_err(y, torch.from_numpy(g["y"])) <= FWD_TOL, case
        (y * torch.from_numpy(g["r"]).float().to(DEV)).sum().backward()
        assert rel_err(x.grad, torch.from_numpy(g["dx"])) <= GRAD_TOL, case
        for n, p in m.named_parameters():
            assert rel_err(p.grad, torch.from_numpy(g["g." + n])) <= GRAD_TOL, (case, n)


def test_golden_models(golden_dir):
    pm, _ = _product()
    cases = {
        "D_16": (lambda: pm.build_discriminator(16, 16, 4, 2, "weight", 0), None),
        "D_20x12_pad": (lambda: pm.build_discriminator(20, 12, 4, 3, "weight", 0), None),
        "R_16": (lambda: pm.build_reverser(16, 16, 4, 2, 8, "weight", 0), None),
        "G_16": (lambda: pm.build_generator(16, 16, 4, 2, 8, "weight"), None),
        "G_20x12_pad": (lambda: pm.build_generator(20, 12, 4, 3, 8, "weight"), None),
        "GLIS_16_k2of3": (lambda: pm.GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional"), 2),
        "GLIS_16_nearest": (lambda: pm.GeneratorLearnedInputSpace(16, 16, 4, 3, 8, "weight", 1, "nearest"), "all"),
        # norm='weight-affine' (common/model.py:31-34): scale and bias on every layer, nn.PReLU activations
        "D_16_affine": (lambda: pm.build_discriminator(16, 16, 4, 2, "weight-affine", 0), None),
    }
    store = _golden(golden_dir, "models.npz")
    for case, (build, depth) in cases.items():
        g = _grp(store, case)
        net = build()
        net.load_state_dict({k[2:]: torch.from_numpy(v).float() for k, v in g.items() if k.startswith("p.")})
        net = net.to(DEV).eval()
        x = torch.from_numpy(g["x"]).float().to(DEV).requires_grad_(True)
        out = net(x, n_execute_lis_layers=depth) if depth is not None else net(x)
        flat = [out[0]] + list(out[1]) if isinstance(out, tuple) else [out]
        loss = 0
        for i, o in enumerate(flat):
            assert rel_err(o, torch.from_numpy(g["y%d" % i])) <= FWD_TOL, (case, i)
            loss = loss + (o * torch.from_numpy(g["r%d" % i]).float().to(DEV)).sum()
        loss.backward()
        assert rel_err(x.grad, torch.from_numpy(g["dx"])) <= GRAD_TOL, case
        grads = {k: p.grad for k, p in zip(net.state_dict().keys(), net.parameters())}
        for k, p in net.named_parameters():
            key = k.replace("·", ".")
            want = torch.from_numpy(g["g." + key])
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            assert rel_err(got, want) <= GRAD_TOL, (case, key)


@pytest.mark.parametrize("fixture,ls", [("glis_steps.npz", False), ("glis_steps_ls.npz", True)])
def test_golden_training_iterations(golden_dir, precision, fixture, ls):
    """Reference-semantics iterations (stock torch optimizers and losses on the reference's own modules): three with
    nn.BCELoss, two with --ls (nn.MSELoss on D's sigmoid output, g_lis/main.py:308-311)."""
    pm, _ = _product()
    from glis_b200.trainer import GLISTrainer
    s = _golden(golden_dir, fixture)
    cfg = _grp(s, "cfg")
    W, H, B, code, nf, nl, n_lis = (int(cfg[k]) for k in ("W", "H", "B", "code", "nf", "nl", "n_lis"))
    gen = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, "fractional")
    dis = pm.build_discriminator(W, H, nf, nl, "weight", 0)
    gen.load_state_dict({k: torch.from_numpy(v).float() for k, v in _grp(s, "init/g").items()})
    dis.load_state_dict({k: torch.from_numpy(v).float() for k, v in _grp(s, "init/d").items()})
    tr = GLISTrainer(gen.to(DEV), dis.to(DEV), lr=float(cfg["lr"]), lambda_r=float(cfg["lam"]), ls=ls)
    for it, (kd, kg) in enumerate(cfg["depths"]):
        g = _grp(s, "it%d" % it)
        f = lambda a: torch.from_numpy(a).float().to(DEV)
        out = tr.step(f(g["real"]), f(g["zd"]), f(g["zg"]), depth_d=int(kd), depth_g=int(kg))
        # iteration 0 starts from identical weights; later ones inherit the (bounded) parameter drift below
        ltol = FWD_TOL if it == 0 else (5e-4 if precision == "fp32" else 3e-3)
        for name in ("d_real", "d_fake", "g"):
            assert abs(out[name].item() - float(g[name])) <= ltol * abs(float(g[name])), (it, name)
        for i, l in enumerate(out["r"]):
            assert abs(l.item() - float(g["r"][i])) <= ltol * abs(float(g["r"][i])), (it, "r", i)
        # lr = 1e-2 makes every RMSprop update O(lr): parameters must track to ~1e-3 of their scale
        # (sign-like early RMSprop steps amplify gradient error where |g| ~ eps; see _run_steps)
        ptol = 1e-2 if precision == "fp32" else 5e-2
        for k, v in gen.state_dict().items():
            assert rel_err(v, torch.from_numpy(g["g/" + k])) <= ptol, (it, "gen", k, rel_err(v, torch.from_numpy(g["g/" + k])))
        for k, v in dis.state_dict().items():
            assert rel_err(v, torch.from_numpy(g["d/" + k])) <= ptol, (it, "dis", k, rel_err(v, torch.from_numpy(g["d/" + k])))


# ---------------------------------------------------------------- whole-step parity vs the oracle
def _make_pair(W, H, nf, nl, code, n_lis, seed=11):
    """fp64 oracle on the CPU and fp32 product on the GPU, identical initial weights."""
    pm, _ = _product()
    torch.manual_seed(seed)
    og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, "fractional")
    od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
    pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, "fractional")
    pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
    copy_params(pg, og)
    copy_params(pd, od)
    return og.double(), od.double(), pg.to(DEV), pd.to(DEV)


def _flat_grads(flat):
    return [flat.g[o:o + p.numel()].view(p.shape).clone() for p, o in zip(flat.params, flat.offsets)]


def _lib_precision_is_fp32():
    from glis_b200 import _lib
    return _lib.default_precision == _lib.PREC_FP32


def _run_steps(og, od, pg, pd, B, H, W, code, depths, lr, seed=5, stats=None):
    """Runs both trainers; checks losses (1e-4), EVERY gradient of every iteration at the per-op bound
    GRAD_TOL = 1e-3 (against the fp64 oracle, relative to the tensor's max |grad|) in both contraction modes,
    and the parameters after each update.

    Flip-aware (oracle/flipaware.py): the product runs first and reports the pre-activation of every TPReLU it
    differentiates; the oracle's backward uses those branch masks, and `check` asserts that they differ from the
    oracle's own only within FLIP_TOL of the kink.  Without that, one pre-activation landing on the other side
    of a kink (probability ~ rounding error x element count) moves every gradient below it by percents and says
    nothing about the kernels.

    The first RMSprop steps behave like lr*g/(0.32|g|+eps): where |g| ~ eps = 1e-6 the update
    amplifies gradient error by lr/eps, so the parameter check bounds |dp - dp_ref| by
    lr * (GRAD_TOL * max|g| / eps + 1e-3 * 3.2) per tensor, the worst case the gradient bound allows."""
    from glis_b200.trainer import GLISTrainer
    ot = GLISOracleTrainer(og, od, lr=lr, lambda_r=0.9)
    pt = GLISTrainer(pg, pd, lr=lr, lambda_r=0.9)
    fa = FlipAwarePair([(og, pg), (od, pd)])
    gen = torch.Generator().manual_seed(seed)
    for it, (kd, kg) in enumerate(depths):
        real = torch.rand(B, 3, H, W, generator=gen)
        zd, zg = torch.randn(B, code, generator=gen), torch.randn(B, code, generator=gen)
        before = [[p.detach().clone() for p in net.parameters()] for net in (og, od)]
        with fa.tap():
            lp = pt.step(real.to(DEV), zd.to(DEV), zg.to(DEV), kd, kg)
        fa.feed()
        lo = ot.step(real.double(), zd.double(), zg.double(), kd, kg)
        flips, total, worst_flip = fa.check(FLIP_TOL)
        if stats is not None:
            stats.append((flips, total, worst_flip))
        for name in ("d_real", "d_fake", "g"):
            assert abs(lp[name].item() - lo[name]) <= FWD_TOL * abs(lo[name]), (it, name, lp[name].item(), lo[name])
        assert len(lp["r"]) == len(lo["r"]) and lp["depth_g"] == lo["depth_g"] and lp["depth_d"] == lo["depth_d"]
        for a, b in zip(lp["r"], lo["r"]):
            assert abs(a.item() - b) <= FWD_TOL * abs(b), (it, "r")
        for tag, onet, pnet, flat, prev in (("gen", og, pg, pt.gen_flat, before[0]), ("dis", od, pd, pt.dis_flat, before[1])):
            names = [n for n, _ in onet.named_parameters()]
            for n, po, pp, gp, p0 in zip(names, onet.parameters(), pnet.parameters(), _flat_grads(flat), prev):
                go = po.grad if po.grad is not None else torch.zeros_like(po)
                gmax = go.abs().max().item()
                if gmax > 0:
                    assert rel_err(gp, go) <= GRAD_TOL, (it, tag, n, rel_err(gp, go), "flips", flips)
                else:
                    assert gp.abs().max().item() == 0, (it, tag, n)
                bound = lr * (GRAD_TOL * gmax / 1e-6 + 3.2e-3)
                err = ((pp.detach().cpu().double() - p0) - (po.detach() - p0)).abs().max().item()
                assert err <= bound, (it, tag, n, err, bound)
        # RMSprop's early steps are sign-like, so fp32 and fp64 trajectories drift apart at elements
        # with |g| ~ eps; re-synchronise (parameters and square averages) so that every iteration is an
        # independent check of losses, gradients and update.  Multi-step drift is covered by
        # test_golden_training_iterations.
        with torch.no_grad():
            for onet, flat, state in ((og, pt.gen_flat, ot.gen_state), (od, pt.dis_flat, ot.dis_state)):
                for po, pp, o in zip(onet.parameters(), flat.params, flat.offsets):
                    pp.copy_(po.float())
                    v = state.get(po)
                    flat.v[o:o + po.numel()].copy_((v if v is not None else torch.zeros_like(po)).reshape(-1).float())
                for pp in flat.params:
                    pp._glis_epoch = getattr(pp, "_glis_epoch", 0) + 1      # cached weight packs are stale now
    fa.remove()
    return ot, pt


def test_step_parity_cfg1_shape():
    """BASELINE config 1: 32x32, batch 32, nfeature 64, 3 levels, code 256, 1 LIS module."""
    og, od, pg, pd = _make_pair(32, 32, 64, 3, 256, 1)
    _run_steps(og, od, pg, pd, 32, 32, 32, 256, [(1, 1), (0, 1), (1, 0)], 2e-5)


def test_step_parity_three_lis_modules_and_skips():
    og, od, pg, pd = _make_pair(16, 16, 8, 2, 32, 3, seed=12)
    _run_steps(og, od, pg, pd, 8, 16, 16, 32, [(3, 2), (0, 0), (1, 3), (2, 1)], 1e-3)


def test_step_parity_padded_nonsquare():
    """W=20 (pads at level 1), H=12, three levels — the `w % 4 == 2` rule on one axis only."""
    og, od, pg, pd = _make_pair(20, 12, 8, 3, 16, 1, seed=13)
    _run_steps(og, od, pg, pd, 4, 12, 20, 16, [(1, 1), (1, 1)], 1e-3)


def test_philox_generators():
    from glis_b200 import ops
    a = ops.randn_(torch.empty(1 << 20, device=DEV), seed=1234)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1) < 5e-3
    assert abs((a ** 4).mean().item() - 3) < 0.1
    b = ops.randn_(torch.empty(1 << 20, device=DEV), seed=1234)
    assert torch.equal(a, b)                                      # counter-based: reproducible
    c = ops.randn_(torch.empty(1 << 20, device=DEV), seed=1234, offset=1 << 18)
    assert not torch.equal(a, c)
    u = ops.uniform_(torch.empty(1 << 20, device=DEV), seed=7)
    assert 0 <= u.min().item() and u.max().item() < 1 and abs(u.mean().item() - 0.5) < 2e-3


def test_loss_kernels():
    from glis_b200 import ops
    lg = torch.randn(64, device=DEV) * 3
    for t in (0.0, 1.0):
        loss, dl, pr = ops.bce_logits(lg, t, want_prob=True)
        l64 = lg.double().cpu().requires_grad_(True)
        ref = torch.nn.functional.binary_cross_entropy(torch.sigmoid(l64), torch.full((64,), t, dtype=torch.float64))
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
        assert rel_err(dl, l64.grad) <= 1e-5 and rel_err(pr, torch.sigmoid(l64)) <= 1e-6
    u, z = torch.randn(64, 256, device=DEV), torch.randn(64, 256, device=DEV)
    du = torch.zeros_like(u)
    loss = ops.mse_scaled(u, z, 0.81, du)
    assert abs(loss.item() - 0.81 * ((u - z) ** 2).mean().item()) <= 1e-5 * loss.item()
    assert rel_err(du, 2 * 0.81 * (u - z) / u.numel()) <= 1e-6


def test_rmsprop_kernel_matches_torch():
    from glis_b200 import ops
    n = 100003
    p = torch.randn(n, device=DEV); g = torch.randn(n, device=DEV); v = torch.rand(n, device=DEV)
    pr = p.clone().cpu().double().requires_grad_(True)
    opt = torch.optim.RMSprop([pr], lr=1e-2, eps=1e-6, alpha=0.9)
    opt.state[pr]["square_avg"] = v.cpu().double().clone(); opt.state[pr]["step"] = torch.tensor(0.)
    pr.grad = g.cpu().double()
    opt.step()
    ops.rmsprop_(p, g, v, 1e-2, 0.9, 1e-6)
    assert rel_err(p, pr) <= 1e-6 and rel_err(v, opt.state[pr]["square_avg"]) <= 1e-6


def test_cli_synthetic_training_checkpoint_and_resume(tmp_path, precision):
    """g_lis/main.py --synthetic: trains, writes reference-named checkpoints, resumes from them."""
    import importlib.util
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("glis_main_gpu", os.path.join(PKG, "g_lis", "main.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    save = str(tmp_path / "exp")
    common = ["--synthetic", "--image_size", "32", "--nfeature", "16", "--code_size", "32", "--norm", "weight",
              "--r_iterations", "2", "--batch_size", "8", "--lr", "0.0002", "--vis_interval", "4", "--vis_size", "2",
              "--save_interval", "4", "--test_interval", "1000", "--precision", precision]
    m.main(common + ["--niter", "6", "--save_path", save])
    arch = os.path.join(save, "net_archive")
    for f in ("last_gen.pt", "last_dis.pt", "last_gen_opt.pt", "last_dis_opt.pt", "last_state.pt", "4_gen.pt"):
        assert os.path.exists(os.path.join(arch, f)), f
    assert os.path.exists(os.path.join(save, "samples", "sample_4.jpg"))
    sd = torch.load(os.path.join(arch, "last_gen.pt"))
    assert "lis_layers.1.lis.1-2.linear.weight" in sd and "conv_layers.0.weight" in sd
    state = torch.load(os.path.join(arch, "last_state.pt"), weights_only=False)
    # the reference's state file (g_lis/main.py:350-357): these six keys, the history pickled
    assert sorted(state) == ["best_iter", "current_iter", "current_sample", "history", "index_shuffle", "min_loss"]
    assert state["current_iter"] == 6 and state["current_sample"] == 6 and isinstance(state["history"], bytes)
    from common.plotting import History
    hist = History.from_string(state["history"])
    assert hist.line_groups["loss-d-mix"].lines["train-d-real"].last_index == 5
    opt_sd = torch.load(os.path.join(arch, "last_dis_opt.pt"))
    torch.optim.RMSprop([torch.nn.Parameter(torch.zeros_like(v["square_avg"])) for v in opt_sd["state"].values()],
                        lr=1e-4).load_state_dict(opt_sd)          # loads into the reference's optimizer class
    m.main(common + ["--niter", "8", "--load_path", save, "--no_graph"])
    state = torch.load(os.path.join(arch, "last_state.pt"), weights_only=False)
    assert state["current_iter"] == 8 and state["current_sample"] == 8
    assert History.from_string(state["history"]).line_groups["loss-d-mix"].lines["train-d-real"].last_index == 7


def test_cli_lsgan_and_d_dropout_run_graphed(tmp_path):
    """--ls and --d_dropout are part of the accelerated path: both run under CUDA-graph replay."""
    import importlib.util
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("glis_main_gpu2", os.path.join(PKG, "g_lis", "main.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    save = str(tmp_path / "exp")
    m.main(["--synthetic", "--image_size", "32", "--nfeature", "16", "--code_size", "32", "--norm", "weight",
            "--r_iterations", "1", "--batch_size", "8", "--lr", "0.0002", "--vis_interval", "100", "--save_interval",
            "100", "--test_interval", "1000", "--niter", "5", "--save_path", save, "--ls", "--d_dropout", "0.3",
            "--g_upscaling", "nearest"])
    state = torch.load(os.path.join(save, "net_archive", "last_state.pt"), weights_only=False)
    assert state["current_iter"] == 5


def test_reference_format_checkpoint_loads(tmp_path):
    """A checkpoint in the REFERENCE's own format loads: optimizer state keyed by `id(param)` as mid-2017
    `Optimizer.state_dict()` wrote it (params listed by id in param_groups), state file with `index_shuffle`,
    `current_sample` and the history as a pickled string."""
    import importlib.util
    from conftest import PKG
    pm, _ = _product()
    from glis_b200.trainer import FlatParams
    spec = importlib.util.spec_from_file_location("glis_main_gpu3", os.path.join(PKG, "g_lis", "main.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    save = str(tmp_path / "ref")
    arch = os.path.join(save, "net_archive")
    os.makedirs(arch); os.makedirs(os.path.join(save, "samples"))
    torch.manual_seed(5)
    g = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 1, "fractional")
    d = pm.build_discriminator(32, 32, 16, 3, "weight", 0)
    want_v = {}
    for tag, net in (("gen", g), ("dis", d)):
        torch.save(net.state_dict(), os.path.join(arch, "last_%s.pt" % tag))
        params = list(net.parameters())
        ids = [1000 + 7 * i for i in range(len(params))]          # stand-ins for id(p) of the saving process
        st = {}
        for k, (pid, p) in enumerate(zip(ids, params)):
            if k == 1:
                continue                                            # a parameter that never stepped: no entry
            st[pid] = {"step": 3, "square_avg": torch.rand_like(p) * 1e-3}
        want_v[tag] = (ids, st)
        torch.save({"state": st, "param_groups": [{"lr": 2e-4, "momentum": 0, "alpha": 0.9, "eps": 1e-6, "centered": False,
                                                    "weight_decay": 0, "params": ids}]},
                   os.path.join(arch, "last_%s_opt.pt" % tag))
    hist = m.new_history(1)
    hist.add_value("loss-d-mix", "train-d-real", 1, 0.69)
    torch.save({"index_shuffle": torch.randperm(50), "current_iter": 3, "best_iter": 0, "min_loss": 1e100,
                "current_sample": 96, "history": hist.to_string().decode("latin1")}, os.path.join(arch, "last_state.pt"))
    torch.save(torch.randn(4, 32), os.path.join(save, "samples", "vis_code.pt"))
    # FlatParams reads the id-keyed state positionally through param_groups
    flat = FlatParams(d.to(DEV))
    flat.load_optimizer_state_dict(torch.load(os.path.join(arch, "last_dis_opt.pt")))
    ids, st = want_v["dis"]
    for k, (pid, p, o) in enumerate(zip(ids, flat.params, flat.offsets)):
        seg = flat.v[o:o + p.numel()].view(p.shape).cpu()
        assert torch.equal(seg, st[pid]["square_avg"] if pid in st else torch.zeros_like(seg)), k
    with pytest.raises(ValueError):
        flat.load_optimizer_state_dict({"state": {12345: {"square_avg": torch.zeros(3)}},
                                        "param_groups": [{"params": list(range(len(flat.params)))}]})
    m.main(["--synthetic", "--image_size", "32", "--nfeature", "16", "--code_size", "32", "--norm", "weight",
            "--r_iterations", "1", "--batch_size", "8", "--lr", "0.0002", "--vis_size", "2", "--vis_interval", "100",
            "--save_interval", "100", "--test_interval", "1000", "--niter", "5", "--load_path", save])
    state = torch.load(os.path.join(arch, "last_state.pt"), weights_only=False)
    assert state["current_iter"] == 5 and state["current_sample"] == 98
    from common.plotting import History
    line = History.from_string(state["history"]).line_groups["loss-d-mix"].lines["train-d-real"]
    assert list(line.get_xs()) == [1, 4, 5]                      # the loaded point, then the two new iterations


def test_cli_r_iterative(tmp_path, precision):
    """r_iterative/main.py --synthetic: trains under graph replay (stochastic do_train schedule), writes the
    reference's seven files per prefix, resumes."""
    import importlib.util
    from conftest import PKG
    spec = importlib.util.spec_from_file_location("riter_main_gpu", os.path.join(PKG, "r_iterative", "main.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    save = str(tmp_path / "exp")
    common = ["--synthetic", "--image_size", "32", "--nfeature", "16", "--code_size", "32", "--norm", "weight",
              "--r_iterations", "2", "--batch_size", "4", "--lr", "0.0002", "--vis_interval", "3", "--vis_size", "2",
              "--save_interval", "4", "--precision", precision]
    m.main(common + ["--niter", "5", "--save_path", save])
    arch = os.path.join(save, "net_archive")
    for f in ("last_gen.pt", "last_gen_opt.pt", "last_r.pt", "last_r_opt.pt", "last_dis.pt", "last_dis_opt.pt",
              "last_state.pt", "4_r.pt"):
        assert os.path.exists(os.path.join(arch, f)), f
    assert os.path.exists(os.path.join(save, "samples", "sample_3_r2.jpg"))
    assert "level.0.conv.weight" in torch.load(os.path.join(arch, "last_r.pt"))
    m.main(common + ["--niter", "7", "--load_path", save, "--no_graph", "--always_train_all"])
    assert torch.load(os.path.join(arch, "last_state.pt"), weights_only=False)["current_iter"] == 7


@pytest.mark.parametrize("flags", [None, [False, True, True], [True, False, True]])
def test_r_iterative_iteration_parity(flags, precision):
    """BASELINE config 5a: one outer iteration of the R-iterative chain (r_iterative/main.py:428-535),
    plain G + reverser R (with its own code-size head) + D, two R hops; all-trained and with skipped hops
    (a skipped hop after a trained one exercises the sticky `last_was_trained` only through `flags`)."""
    pm, _ = _product()
    from glis_b200.trainer import RIterTrainer
    from oracle.step import riter_iteration
    W = H = 16; nf, nl, code, B, R = 8, 2, 16, 6, 2
    torch.manual_seed(21)
    og, orv, od = (oracle.build_generator(W, H, nf, nl, code, "weight"),
                   oracle.build_reverser(W, H, nf // 2, nl, code, "weight", 0),
                   oracle.build_discriminator(W, H, nf, nl, "weight"))
    pg, prv, pd = (pm.build_generator(W, H, nf, nl, code, "weight"),
                   pm.build_reverser(W, H, nf // 2, nl, code, "weight", 0),
                   pm.build_discriminator(W, H, nf, nl, "weight"))
    for a, b in ((pg, og), (prv, orv), (pd, od)):
        copy_params(a, b)
    og, orv, od = og.double(), orv.double(), od.double()
    tr = RIterTrainer(pg.to(DEV), prv.to(DEV), pd.to(DEV), lr=1e-3, lambda_r=0.9, r_iterations=R)
    gen = torch.Generator().manual_seed(3)
    z = torch.randn(B, code, generator=gen)
    n_trained = (1 + R) if flags is None else sum(flags)
    reals = [torch.rand(B, 3, H, W, generator=gen) for _ in range(n_trained)]
    gs, rs, ds = {}, {}, {}
    want = riter_iteration(og, orv, od, gs, rs, ds, z.double(), [x.double() for x in reals], 1e-3, 0.9, R, flags)
    got = tr.step(z.to(DEV), [x.to(DEV) for x in reals], flags)
    ltol = FWD_TOL if precision == "fp32" else 2e-3     # later hops inherit the updated (slightly drifted) nets
    for hop, (w, g) in enumerate(zip(want, got)):
        assert (w is None) == (g is None), hop
        if w is None:
            continue
        assert set(w) == set(g)
        for k in w:
            assert abs(g[k].item() - w[k]) <= (FWD_TOL if hop == 0 else ltol * 5) * abs(w[k]) + 1e-7, (hop, k, g[k].item(), w[k])
    ptol = 5e-3 if precision == "fp32" else 5e-2
    for pnet, onet in ((pg, og), (prv, orv), (pd, od)):
        for (k, v), (_, wv) in zip(pnet.state_dict().items(), onet.state_dict().items()):
            assert rel_err(v, wv) <= ptol, (k, rel_err(v, wv))


def test_r_iterative_train_flag_schedule():
    """`do_train` draw: p = (r+1)/(1+R), sticky once true, forced on the last hop (r_iterative/main.py:445-451)."""
    pm, _ = _product()
    from glis_b200.trainer import RIterTrainer

    class Seq(object):
        def __init__(self, vals): self.vals = list(vals)
        def random(self): return self.vals.pop(0)
    W = H = 16
    tr = RIterTrainer(pm.build_generator(W, H, 8, 2, 16, "weight").to(DEV), pm.build_reverser(W, H, 4, 2, 16, "weight", 0).to(DEV),
                      pm.build_discriminator(W, H, 8, 2, "weight").to(DEV), lr=1e-3, r_iterations=3)
    tr.rng = Seq([0.9, 0.9, 0.9, 0.9]); assert tr.draw_train_flags() == [False, False, False, True]
    tr.rng = Seq([0.9, 0.4, 0.99, 0.99]); assert tr.draw_train_flags() == [False, True, True, True]
    tr.rng = Seq([0.2, 0.99, 0.99, 0.99]); assert tr.draw_train_flags() == [True, True, True, True]
    assert tr.draw_train_flags(always_train_all=True) == [True] * 4


def _make_pair_kw(W, H, nf, nl, code, n_lis, upscaling="fractional", seed=31):
    pm, _ = _product()
    torch.manual_seed(seed)
    og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, upscaling)
    od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
    pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, upscaling)
    pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
    copy_params(pg, og)
    copy_params(pd, od)
    return og.double(), od.double(), pg.to(DEV), pd.to(DEV)


def test_step_parity_config4_geometry():
    """BASELINE config 4 geometry: 160x160, five levels (the extra G/D layer), 1 LIS module; reduced
    width (nfeature 32) and batch so that the fp64 oracle stays fast.  Exercises 80- and 40-wide maps."""
    og, od, pg, pd = _make_pair_kw(160, 160, 32, 5, 64, 1)
    _run_steps(og, od, pg, pd, 2, 160, 160, 64, [(1, 1)], 2e-5)


def test_step_parity_config5b_nearest_upsampling():
    """BASELINE config 5b family: 3 levels, `--g_upscaling nearest` (3x3 stride-1 WN convs after a
    nearest-neighbour upsample), 1 LIS module."""
    og, od, pg, pd = _make_pair_kw(32, 32, 32, 3, 32, 1, upscaling="nearest")
    _run_steps(og, od, pg, pd, 4, 32, 32, 32, [(1, 1), (0, 1)], 1e-3)


# ---------------------------------------------------------------- the benchmarked shapes, at size
def _tc_plan(relation, n, hi, wi, ci, ho, wo, co, plain, k=4, s=2, p=1):
    """The launch plan of the tensor-core kernel this geometry runs on — the halo kernel (glis_conv_tc_halo_plan,
    key "halo" = True) when it applies, else tc_conv_kernel (glis_conv_tc_plan) — or None off the tensor cores."""
    import ctypes as C
    from glis_b200 import _lib as L, ops
    g = ops.ContractionSpec(False, (k, k), (s, s), (p, p), (1, 1)).geom(relation, n, hi, wi, ci, ho, wo, co)
    lib = L.load()
    if not lib.glis_conv_tc_supported(C.byref(g)):
        return None
    out = (C.c_int * 20)()
    if lib.glis_conv_tc_halo_plan(C.byref(g), int(plain), out) == 0:
        names = ("tw th tn n_mma tmem kblocks ksplit a_rows stages tiles_h tiles_x tiles_co total groups smem "
                 "hx hy box_rows x_rows classes").split()
        d = dict(zip(names, list(out)), halo=True)
        d["bk"], d["classes"] = d["classes"] // 1000, d["classes"] % 1000      # (channels per stage rides in the last slot)
        return d
    out = (C.c_int * 15)()
    assert lib.glis_conv_tc_plan(C.byref(g), int(plain), out) == 0
    names = "tw th tn n_mma tmem kblocks ksplit a_rows stages tiles_h tiles_x tiles_co total groups smem".split()
    return dict(zip(names, list(out)), halo=False)


def _num_sms():
    return torch.cuda.get_device_properties(0).multi_processor_count


# (kind, Cin, Cout, input H = W, batch): every conv / transposed-conv layer of BASELINE config 2 (80x80, nfeature
# 64, 4 levels, code 256) at the batch the benchmark runs it with — D on the 2B = 128 image batch of the D update
# and on the B = 64 batch of the G update, G at B = 64 — each as the fused (WN layer, TPReLU) pair of the builders.
CFG2_LAYERS = [
    ("conv", 3, 64, 80, 128), ("conv", 64, 128, 40, 128), ("conv", 128, 256, 20, 128), ("conv", 256, 512, 10, 128),
    ("conv", 3, 64, 80, 64), ("conv", 64, 128, 40, 64), ("conv", 128, 256, 20, 64), ("conv", 256, 512, 10, 64),
    ("deconv", 512, 256, 5, 64), ("deconv", 256, 128, 10, 64), ("deconv", 128, 64, 20, 64),
    ("deconv_sigmoid", 64, 3, 40, 64),
    ("head", 256, 512, 5, 64),
]


def _check_chain(ref_layers, prod_layers, x, seed=0):
    """Oracle (fp64, CPU) against the product's fused chain (`run_layers`) on equal parameters: forward <= 1e-4,
    input and every parameter gradient <= 1e-3, flip-aware (see `_run_steps`)."""
    pm, _ = _product()
    ref = torch.nn.Sequential(*ref_layers).double()
    prod = torch.nn.Sequential(*prod_layers)
    fa = FlipAwarePair([(ref, prod)])
    gen = torch.Generator().manual_seed(seed)
    xr = x.double().clone().requires_grad_(True)
    xp = x.float().to(DEV).requires_grad_(True)
    with fa.tap():
        yp = pm.run_layers(prod_layers, xp)
    r = torch.rand(yp.shape, generator=gen, dtype=torch.float64) * 2 - 1
    (yp * r.float().to(DEV)).sum().backward()
    fa.feed()
    yr = ref(xr)
    assert tuple(yr.shape) == tuple(yp.shape)
    (yr * r).sum().backward()
    flips, total, worst = fa.check(FLIP_TOL)
    fa.remove()
    assert rel_err(yp, yr) <= FWD_TOL, "forward %g" % rel_err(yp, yr)
    assert rel_err(xp.grad, xr.grad) <= GRAD_TOL, "dx %g (flips %d of %d)" % (rel_err(xp.grad, xr.grad), flips, total)
    for (n, pr), (_, pp) in zip(ref.named_parameters(), prod.named_parameters()):
        e = rel_err(pp.grad, pr.grad)
        assert e <= GRAD_TOL, "grad %s %g (flips %d of %d)" % (n, e, flips, total)
    return flips, total


@pytest.mark.parametrize("case", CFG2_LAYERS, ids=lambda c: "%s_%dto%d_%d_n%d" % c)
def test_cfg2_layer_at_size(case, precision):
    """Forward + data gradient + weight gradient of every config-2 layer at the benchmark's batch against the fp64
    oracle.  These are the launches bench.py times: persistent CTAs walking SEVERAL tiles (accumulator double
    buffering, the shared-memory ring running across tile boundaries), 208-column tiles with 2 pipeline stages,
    split-K shares, 1600 pixel tiles on the pixel-major kernel — none of which the small shapes above reach."""
    from glis_b200 import _lib as L
    _, pmod = _product()
    kind, ci, co, h, n = case
    gen = torch.Generator().manual_seed(100 + CFG2_LAYERS.index(case))
    if kind == "head":
        f, hh = co, h
        mk = lambda M: [M.WeightNormalizedLinear(ci, f * hh * hh, init_factor=0.01, scale=False, bias=False),
                        M.View(f, hh, hh), M.TPReLU(f)]
        x = torch.randn(n, ci, generator=gen)
    elif kind == "deconv_sigmoid":
        mk = lambda M: [M.WeightNormalizedConvTranspose2d(ci, co, 4, 2, 1), torch.nn.Sigmoid()]
        x = torch.rand(n, ci, h, h, generator=gen) * 2 - 1
    else:
        layer = "WeightNormalizedConv2d" if kind == "conv" else "WeightNormalizedConvTranspose2d"
        mk = lambda M: [getattr(M, layer)(ci, co, 4, 2, 1, scale=False, bias=False), M.TPReLU(co)]
        x = torch.rand(n, ci, h, h, generator=gen) * 2 - 1
    ref_layers = mk(oracle)
    ref_net = torch.nn.Sequential(*ref_layers)
    randomize_params_(ref_net, gen)
    prod_layers = mk(pmod)
    copy_params(torch.nn.Sequential(*prod_layers), ref_net)
    prod_layers = [m.to(DEV) for m in prod_layers]
    _check_chain(ref_layers, prod_layers, x)
    if precision != "bf16x3" or kind in ("head", "deconv_sigmoid") or ci == 3:
        return
    # the launches really are the multi-tile ones (plans are a function of the geometry and the SM count)
    sms = _num_sms()
    if kind == "conv":
        fwd = _tc_plan(L.CONV, n, h, h, ci, h // 2, h // 2, co, False)
        dgrad = _tc_plan(L.TCONV, n, h // 2, h // 2, co, h, h, ci, True)
    else:
        fwd = _tc_plan(L.TCONV, n, h, h, ci, 2 * h, 2 * h, co, False)
        dgrad = _tc_plan(L.CONV, n, 2 * h, 2 * h, co, h, h, ci, True)
    assert fwd is not None and dgrad is not None
    halo = os.environ.get("GLIS_TC_HALO", "1") != "0"
    # the halo kernel takes the launches whose (phase) output grid is at least 16 pixels wide
    fq = h // 2 if kind == "conv" else h          # forward: conv output grid / transposed-conv phase grid
    dq = h // 2 if kind == "conv" else h          # data gradient: phase grid of the conv's input / the deconv's input
    assert fwd["halo"] == (halo and fq >= 16) and dgrad["halo"] == (halo and dq >= 16), (fwd, dgrad)
    if (kind, ci, n) == ("conv", 64, 128):         # D level 1 of the 2B pass: the roofline kernel of bench.py
        assert fwd["groups"] > sms and fwd["n_mma"] >= 208, fwd
        assert dgrad["groups"] > 3 * sms, dgrad
    if (kind, ci, n) == ("deconv", 128, 64):       # G level 1: 64-row weight tiles, several work items per CTA
        assert fwd["groups"] > 3 * sms and fwd["a_rows"] == 64, fwd
    if (kind, ci, n) == ("conv", 128, 128):
        assert dgrad["groups"] > sms and dgrad["n_mma"] >= 208, dgrad


@pytest.mark.parametrize("case", [("conv", 128, 256, 20, 128), ("conv", 256, 512, 10, 128), ("conv", 256, 512, 10, 64),
                                  ("deconv", 512, 256, 5, 64)], ids=lambda c: "%s_%dto%d_%d_n%d" % c)
def test_cta_pair_kernel_at_size(case, precision, monkeypatch):
    """The cta_group::2 form (csrc/tc_conv_pair.cu, GLIS_TC_PAIR=1) on the config-2 layers it takes — >= 256 output
    channels on 5x5 / 10x10 maps, forward with the fused epilogue and the data gradients with 512 / 256 outputs —
    against the fp64 oracle at the benchmark's batch: forward 1e-4, gradients 1e-3."""
    import ctypes as C
    from glis_b200 import _lib as L, ops
    if precision != "bf16x3":
        pytest.skip("tensor-core mode only")
    monkeypatch.setenv("GLIS_TC_PAIR", "1")
    _, pmod = _product()
    kind, ci, co, h, n = case
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    if kind == "conv":
        g = spec.geom(L.CONV, n, h, h, ci, h // 2, h // 2, co)
    else:
        g = spec.geom(L.TCONV, n, h, h, ci, 2 * h, 2 * h, co)
    out = (C.c_int * 16)()
    assert L.load().glis_conv_tc_pair_plan(C.byref(g), 0, out) == 0, "the pair kernel should take this forward launch"
    assert out[6] % 16 == 0 and out[6] <= 256 and out[10] >= 2      # UMMA N, pipeline stages
    gen = torch.Generator().manual_seed(300 + h)
    layer = "WeightNormalizedConv2d" if kind == "conv" else "WeightNormalizedConvTranspose2d"
    mk = lambda M: [getattr(M, layer)(ci, co, 4, 2, 1, scale=False, bias=False), M.TPReLU(co)]
    x = torch.rand(n, ci, h, h, generator=gen) * 2 - 1
    ref_layers = mk(oracle)
    ref_net = torch.nn.Sequential(*ref_layers)
    randomize_params_(ref_net, gen)
    prod_layers = mk(pmod)
    copy_params(torch.nn.Sequential(*prod_layers), ref_net)
    prod_layers = [m.to(DEV) for m in prod_layers]
    _check_chain(ref_layers, prod_layers, x)


def test_persistent_multi_tile_paths_are_covered():
    """The plans of the config-2 launches at the benchmark's batch: at least one fused-epilogue forward and one
    plain-output data gradient run more work items than there are SMs (so test_cfg2_layer_at_size exercises the
    persistent loop), and the pixel-major image-side product runs 1600 pixel tiles."""
    from glis_b200 import _lib as L
    sms = _num_sms()
    fwd = _tc_plan(L.CONV, 128, 40, 40, 64, 20, 20, 128, False)
    assert fwd["groups"] == 256 and fwd["groups"] > sms and fwd["n_mma"] >= 208
    if fwd["halo"]:      # 10 rows of 20 pixels + halo: 11 x 21 box, accumulator columns 9 * 21 + 20 = 209 -> 224
        assert (fwd["th"], fwd["hx"], fwd["hy"], fwd["box_rows"], fwd["n_mma"], fwd["classes"]) == (10, 1, 1, 231, 224, 4)
    assert _tc_plan(L.TCONV, 128, 20, 20, 128, 40, 40, 64, True)["groups"] >= 1024
    assert _tc_plan(L.TCONV, 64, 20, 20, 128, 40, 40, 64, False)["groups"] >= 512
    assert 128 * 40 * 40 // 128 == 1600            # D level 0 at 2B images: M = 128 pixels per tile


def _whole_iteration_at_size(W, H, B, nf, nl, code, n_lis, depths, seed):
    og, od, pg, pd = _make_pair_kw(W, H, nf, nl, code, n_lis, seed=seed)
    stats = []
    _run_steps(og, od, pg, pd, B, H, W, code, depths, 2e-5, stats=stats)
    return stats


def test_whole_iteration_config2_at_size():
    """BASELINE configs[1] exactly as bench.py runs it: 80x80, batch 64, nfeature 64, 4 levels, code 256, 1 LIS
    module, lr 2e-5 — one full iteration (D on 2B images, D update, G + LIS against D, G update) against the fp64
    oracle: losses 1e-4, every gradient 1e-3, the RMSprop update."""
    stats = _whole_iteration_at_size(80, 80, 64, 64, 4, 256, 1, [(1, 1)], seed=71)
    flips, total, _ = stats[0]
    assert total > 30e6 and flips <= total * 1e-4, stats      # ~39 M differentiated pre-activations per iteration


def test_whole_iteration_config3_at_size():
    """BASELINE configs[2]: as config 2 with 3 chained LIS modules (all run in the G update, two in the D update)."""
    _whole_iteration_at_size(80, 80, 64, 64, 4, 256, 3, [(2, 3)], seed=72)


def test_whole_iteration_config4_at_size():
    """BASELINE configs[3]: 160x160, batch 32, nfeature 64, 5 levels (K = 8192 layers, 80- and 40-wide maps)."""
    _whole_iteration_at_size(160, 160, 32, 64, 5, 256, 1, [(1, 1)], seed=73)


def test_discriminator_dropout_runs_and_matches_in_eval():
    """`--d_dropout p` (model.py:52-53): stochastic in training (masks are not comparable across
    implementations), identical to the oracle in eval mode."""
    pm, _ = _product()
    torch.manual_seed(41)
    od = oracle.build_discriminator(32, 32, 16, 3, "weight", 0.3)
    pd = pm.build_discriminator(32, 32, 16, 3, "weight", 0.3)
    copy_params(pd, od)
    x = torch.rand(4, 3, 32, 32)
    pd = pd.to(DEV)
    assert "final·dropout" in pd._modules
    y_train = pd(x.to(DEV))
    assert y_train.shape == (4, 1) and torch.isfinite(y_train).all()
    assert rel_err(pd.eval()(x.to(DEV)), od.double().eval()(x.double())) <= FWD_TOL


def test_cuda_graph_replay_matches_eager(precision):
    """GraphedStep (one captured graph per LIS-depth pair, replayed from static buffers) must follow
    the same trajectory as launching the step eagerly: same losses and parameters after 6 iterations
    that alternate between two depth pairs."""
    pm, _ = _product()
    from glis_b200.trainer import GLISTrainer, GraphedStep
    W = H = 32; nf, nl, code, B = 16, 3, 32, 8
    nets = []
    for _ in range(2):
        torch.manual_seed(51)
        nets.append((pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 2, "fractional").to(DEV),
                     pm.build_discriminator(W, H, nf, nl, "weight", 0).to(DEV)))
    # early RMSprop steps are sign-like (every weight moves ~3.2*lr whatever its gradient), so with a
    # large lr two correct runs diverge through atomics-order noise alone; the reference's lr keeps
    # the comparison meaningful
    eager = GLISTrainer(nets[0][0], nets[0][1], lr=2e-5)
    graphed = GraphedStep(GLISTrainer(nets[1][0], nets[1][1], lr=2e-5), B, H, W, code, DEV, warmup=1)
    gen = torch.Generator().manual_seed(9)
    for it, depth in enumerate([(2, 1), (0, 2), (2, 1), (2, 1), (0, 2), (2, 1)]):
        real = torch.rand(B, 3, H, W, generator=gen).to(DEV)
        zd, zg = torch.randn(B, code, generator=gen).to(DEV), torch.randn(B, code, generator=gen).to(DEV)
        a = eager.step(real, zd, zg, *depth)
        b = graphed.step(real, zd, zg, *depth)
        for k in ("d_real", "d_fake", "g"):
            assert abs(a[k].item() - b[k].item()) <= 2e-3 * abs(a[k].item()), (it, k, a[k].item(), b[k].item())
    # a weight whose gradient is ~eps can step the other way (atomics order): <= 2*3.2*lr per iteration
    for (k, v), (_, w) in zip(nets[0][0].state_dict().items(), nets[1][0].state_dict().items()):
        assert (w - v).abs().max().item() <= 6 * 6.4 * 2e-5 + 1e-7, k
    for (k, v), (_, w) in zip(nets[0][1].state_dict().items(), nets[1][1].state_dict().items()):
        assert (w - v).abs().max().item() <= 6 * 6.4 * 2e-5 + 1e-7, k


def test_host_fed_stepper_matches_direct_steps(precision):
    """HostFedStepper (pinned host batches, H2D on a copy stream overlapped with the previous
    iteration, losses read one iteration late) follows the same trajectory as feeding the same
    batches directly, and returns every iteration's losses exactly once, in order."""
    pm, _ = _product()
    from glis_b200.trainer import GLISTrainer, GraphedStep, HostFedStepper
    W = H = 32; nf, nl, code, B = 16, 3, 32, 8
    runs = []
    gen = torch.Generator().manual_seed(13)
    batches = [(torch.rand(B, 3, H, W, generator=gen).pin_memory(), torch.randn(B, code, generator=gen).pin_memory(),
                torch.randn(B, code, generator=gen).pin_memory()) for _ in range(5)]
    for fed in (False, True):
        torch.manual_seed(61)
        g = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional").to(DEV)
        d = pm.build_discriminator(W, H, nf, nl, "weight", 0).to(DEV)
        gs = GraphedStep(GLISTrainer(g, d, lr=2e-5), B, H, W, code, DEV, warmup=1)
        losses = []
        if fed:
            feeder = HostFedStepper(gs)
            for real, zd, zg in batches:
                r = feeder.submit(real, zd, zg, 1, 1)
                if r is not None:
                    losses.append(r)
            losses.append(feeder.flush())
        else:
            for real, zd, zg in batches:
                o = gs.step(real.to(DEV), zd.to(DEV), zg.to(DEV), 1, 1)
                losses.append({"d_real": o["d_real"].item(), "d_fake": o["d_fake"].item(), "g": o["g"].item(),
                               "r": [v.item() for v in o["r"]]})
        runs.append((losses, torch.cat([gs.tr.gen_flat.p, gs.tr.dis_flat.p]).clone()))
    (la, pa), (lb, pb) = runs
    assert len(la) == len(lb) == 5
    for i, (a, b) in enumerate(zip(la, lb)):
        for k in ("d_real", "d_fake", "g"):
            assert abs(a[k] - b[k]) <= 2e-3 * abs(a[k]), (i, k, a[k], b[k])
        assert len(a["r"]) == len(b["r"]) == 1
    assert (pa - pb).abs().max().item() <= 5 * 6.4 * 2e-5 + 1e-7


# ---------------------------------------------------------------- kernels added for the image side / head
def test_unfold_fold_kernels_match_torch():
    """csrc/image_side.cu against torch: unfold = the 4x4/s2/p1 patches in (c, kh, kw) column order as
    hi+lo bf16 planes; fold = its adjoint (+ bias, sigmoid)."""
    import torch.nn.functional as F
    from glis_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(21)
    for (n, c, h, w) in ((3, 3, 8, 12), (2, 3, 80, 80), (2, 1, 6, 4)):
        x = torch.rand(n, c, h, w, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
        hi, lo = ops.unfolded_planes(x)
        got = (hi.float() + lo.float())                                   # (n, h/2, w/2, 16c)
        want = F.unfold(x.contiguous(), 4, padding=1, stride=2)           # (n, c*16, L), rows (c, kh, kw)
        want = want.view(n, c * 16, h // 2, w // 2).permute(0, 2, 3, 1)
        assert rel_err(got, want) <= 2e-5, (n, c, h, w)
        cols = torch.randn(n, h // 2, w // 2, 16 * c, generator=g).to(DEV)
        bias = torch.randn(c, generator=g).to(DEV)
        for act, fn in ((L.ACT_NONE, lambda t: t), (L.ACT_SIGMOID, torch.sigmoid)):
            out = torch.empty(n, c, h, w, device=DEV).contiguous(memory_format=torch.channels_last)
            L.call("glis_fold4x4s2", L.ptr(cols), n, h // 2, w // 2, c, L.ptr(bias), act, L.ptr(out), L.stream())
            ref = F.fold(cols.permute(0, 3, 1, 2).reshape(n, 16 * c, -1), (h, w), 4, padding=1, stride=2)
            ref = fn(ref + bias.view(1, c, 1, 1))
            assert rel_err(out, ref) <= 1e-6, (n, c, h, w, act)


def test_linear_wgrad_with_row_permutation():
    """glis_linear_wgrad: G[row(a)][b] += sum_m dy[m][a] x[m][b], row(a) = (a % C)*P + a / C."""
    from glis_b200 import _lib as L
    g = torch.Generator().manual_seed(22)
    for (m, c, p, k) in ((64, 8, 25, 40), (130, 16, 4, 256), (5, 3, 7, 9)):
        ca = c * p
        dy, x = torch.randn(m, ca, generator=g).to(DEV), torch.randn(m, k, generator=g).to(DEV)
        G = torch.zeros(ca, k, device=DEV)
        L.call("glis_linear_wgrad", L.ptr(dy), L.ptr(x), L.ptr(G), m, ca, k, c, p, L.stream())
        rows = torch.tensor([(a % c) * p + a // c for a in range(ca)], device=DEV)
        want = torch.zeros(ca, k, dtype=torch.float64, device=DEV)
        want[rows] = dy.double().t() @ x.double()
        assert rel_err(G, want) <= 1e-5, (m, c, p, k)
        G2 = torch.zeros(ca, k, device=DEV)
        L.call("glis_linear_wgrad", L.ptr(dy), L.ptr(x), L.ptr(G2), m, ca, k, 0, 0, L.stream())
        assert rel_err(G2, dy.double().t() @ x.double()) <= 1e-5


def test_generator_head_fused_matches_unfused(precision):
    """linear -> View -> TPReLU as one kernel (row-permuted packs, NHWC output) against the three modules
    run one by one; forward, input gradient and every parameter gradient."""
    pm, pmod = _product()
    from glis_b200 import ops
    torch.manual_seed(23)
    code, f, h, w, b = 32, 16, 5, 3, 6
    lin = pmod.WeightNormalizedLinear(code, f * h * w, init_factor=0.01, scale=False, bias=False).to(DEV)
    view, act = pmod.View(f, h, w), pmod.TPReLU(f).to(DEV)
    with torch.no_grad():
        act.weight.uniform_(-0.2, 1.2)
        act.bias.uniform_(-0.3, 0.3)
    x = torch.randn(b, code, device=DEV, requires_grad=True)
    r = torch.randn(b, f, h, w, device=DEV)
    y1 = pm.run_layers([lin, view, act], x)                      # the fused operator
    (y1 * r).sum().backward()
    g1 = [x.grad.clone()] + [p.grad.clone() for p in (lin.weight, act.weight, act.bias)]
    x.grad = None
    for p in (lin.weight, act.weight, act.bias):
        p.grad = None
    y2 = act(view(lin(x)))                                       # module by module
    (y2 * r).sum().backward()
    g2 = [x.grad] + [p.grad for p in (lin.weight, act.weight, act.bias)]
    tol = 1e-6 if precision == "fp32" else 2e-5
    assert y1.shape == y2.shape and rel_err(y1, y2) <= tol
    for a, bb in zip(g1, g2):
        assert rel_err(a, bb) <= 10 * tol


@pytest.mark.parametrize("knob", ["GLIS_TC_KSPLIT=1", "GLIS_TC_CLUSTER=2", "GLIS_TC_CLUSTER=4", "GLIS_WG_COLS=128"])
def test_tensor_core_tuning_knobs_do_not_change_results(knob, monkeypatch):
    """Split-K, the cluster multicast of the weight tile and the accumulator width are scheduling choices:
    the contraction they compute is the same (summation order aside)."""
    from glis_b200 import _lib
    _lib.set_precision("bf16x3")
    _, pmod = _product()

    def run():
        torch.manual_seed(24)
        conv = pmod.WeightNormalizedConv2d(256, 128, 4, 2, 1, scale=False, bias=False).to(DEV)
        x = torch.randn(6, 256, 10, 10, device=DEV, requires_grad=True)
        y = conv(x)
        (y * torch.linspace(-1, 1, y.numel(), device=DEV).view_as(y)).sum().backward()
        return y.detach(), x.grad, conv.weight.grad

    base = run()
    name, val = knob.split("=")
    monkeypatch.setenv(name, val)
    for a, b in zip(run(), base):
        assert rel_err(a, b) <= 2e-5, knob


def test_side_stream_overlap_matches_single_stream(precision):
    """The trainer's side-stream work (weight gradients, pack rebuilds) only reorders kernels: same losses
    and gradients as the single-stream schedule.  The two trainers run in LOCKSTEP — before every iteration
    the single-stream one takes the other's parameters and optimizer state — because training itself is
    chaotic at rounding level (split-K sums arrive in a different order every run, RMSprop divides
    near-zero gradients by eps, TPReLU masks flip): two runs of the SAME schedule drift apart by 1e-3 on
    G's gradient within two iterations (tools/debug_overlap.py), which says nothing about the schedule."""
    from glis_b200 import ops
    from glis_b200.trainer import GLISTrainer
    pm, pmod = _product()

    def make():
        torch.manual_seed(25)
        g = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 1, "fractional").to(DEV)
        d = pm.build_discriminator(32, 32, 16, 3, "weight", 0).to(DEV)
        # slopes of 1: no TPReLU kink, hence no mask flips — the two schedules add their split-K partial sums in
        # different orders, and ONE pre-activation landing on the other side of a kink would move every
        # gradient by ~1e-3 (seen once in ~10 runs), drowning what this test is after: a stale pack or a
        # missing stream dependency, which do not care about slopes
        for m in list(g.modules()) + list(d.modules()):
            if isinstance(m, pmod.TPReLU):
                m.weight.data.fill_(1.0)
        return GLISTrainer(g, d, lr=1e-4)

    side, single = make(), make()
    gen = torch.Generator().manual_seed(26)
    try:
        for it in range(4):
            for fa, fb in ((side.gen_flat, single.gen_flat), (side.dis_flat, single.dis_flat)):
                fb.p.copy_(fa.p)
                fb.v.copy_(fa.v)
                for p in fb.params:
                    p._glis_epoch = getattr(p, "_glis_epoch", 0) + 1    # its packs are stale now
            args = (torch.rand(8, 3, 32, 32, generator=gen).to(DEV), torch.randn(8, 32, generator=gen).to(DEV),
                    torch.randn(8, 32, generator=gen).to(DEV), 1, 1)
            outs = []
            for tr, overlap in ((side, True), (single, False)):
                ops.Overlap.enabled = overlap
                o = tr.step(*args)
                torch.cuda.synchronize()
                outs.append([o[k].item() for k in ("d_real", "d_fake", "g")] + [r.item() for r in o["r"]])
            assert np.allclose(outs[0], outs[1], rtol=1e-5), (it, outs)
            for fa, fb in ((side.gen_flat, single.gen_flat), (side.dis_flat, single.dis_flat)):
                assert rel_err(fa.g, fb.g) <= 1e-5, it
    finally:
        ops.Overlap.enabled = True


# ---------------------------------------------------------------- split-K forward, planes-only activations
def test_tprelu_forward_planes_kernel_matches_torch():
    """glis_tprelu_forward_planes (the pointwise epilogue of a split-K launch) against torch, incl. the
    generator head's channel = feature % C indexing."""
    import torch.nn.functional as F
    from glis_b200 import _lib as L
    g = torch.Generator().manual_seed(41)
    for (pix, c, ca) in ((50, 512, 0), (7, 8, 0), (6, 240, 16)):
        x = torch.randn(pix, c, generator=g).to(DEV)
        cc = ca or c
        a, b = (torch.rand(cc, generator=g) * 1.4 - 0.2).to(DEV), (torch.randn(cc, generator=g) * 0.3).to(DEV)
        out = torch.empty_like(x)
        hi, lo = torch.empty_like(x, dtype=torch.bfloat16), torch.empty_like(x, dtype=torch.bfloat16)
        L.call("glis_tprelu_forward_planes", L.ptr(x), L.ptr(a), L.ptr(b), L.ptr(out), L.ptr16(hi), L.ptr16(lo),
               x.numel(), c, ca, L.stream())
        idx = torch.arange(c, device=DEV) % cc
        t = x - b[idx]
        want = torch.where(t > 0, t, a.clamp(0, 1)[idx] * t) + b[idx]
        assert rel_err(out, want) <= 1e-6
        assert rel_err(hi.float() + lo.float(), want) <= 2e-5
        out2 = torch.empty_like(x)
        L.call("glis_tprelu_forward_planes", L.ptr(x), L.ptr(a), L.ptr(b), L.ptr(out2), None, None,
               x.numel(), c, ca, L.stream())
        assert torch.equal(out2, out)
        # three partial-sum slabs (a split-K launch's shares), summed in order, pre-activation written too
        parts = torch.randn(3, pix, c, generator=g).to(DEV)
        parts[2] = x - parts[0] - parts[1]
        pre, out3 = torch.empty_like(x), torch.empty_like(x)
        L.call("glis_tprelu_forward_planes_sum", L.ptr(parts), 3, x.numel(), L.ptr(a), L.ptr(b), L.ptr(pre),
               L.ptr(out3), None, None, x.numel(), c, ca, L.stream())
        assert torch.equal(pre, parts[0] + parts[1] + parts[2])
        assert rel_err(out3, want) <= 1e-5


def test_split_k_forward_matches_fused_epilogue():
    """D's last level at batch 64 (256 -> 512 channels, 10x10 -> 5x5: 1600 pixels, 64 k-steps) runs as split-K
    sums + one pointwise TPReLU pass; same forward, input gradient and parameter gradients as the fused launch."""
    from glis_b200 import _lib, ops
    _lib.set_precision("bf16x3")
    pm, pmod = _product()
    spec_g = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1)).geom(_lib.CONV, 64, 10, 10, 256, 5, 5, 512)
    assert ops._split_k_forward(spec_g, _lib.CONV)
    wide = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1)).geom(_lib.CONV, 128, 40, 40, 64, 20, 20, 128)
    assert not ops._split_k_forward(wide, _lib.CONV)      # 256 tiles fill the machine on their own

    def run(split):
        ops.SPLIT_K_FORWARD = split
        torch.manual_seed(42)
        conv = pmod.WeightNormalizedConv2d(256, 512, 4, 2, 1, scale=False, bias=False).to(DEV)
        act = pmod.TPReLU(512).to(DEV)
        with torch.no_grad():
            act.weight.uniform_(-0.2, 1.2)
            act.bias.uniform_(-0.3, 0.3)
        x = torch.randn(64, 256, 10, 10, device=DEV, requires_grad=True)
        y = pm.run_layers([conv, act], x)
        (y * torch.linspace(-1, 1, y.numel(), device=DEV).view_as(y)).sum().backward()
        return y.detach(), x.grad, conv.weight.grad, act.weight.grad, act.bias.grad

    try:
        a, b, a2 = run(True), run(False), run(True)
    finally:
        ops.SPLIT_K_FORWARD = True
    assert rel_err(a[0], b[0]) <= 2e-5
    assert torch.equal(a[0], a2[0])      # per-share slabs summed in a fixed order: the forward is bit-reproducible
    # The two launches add the K blocks in a different order, so 1-2 of the 819 200 pre-activations that sit
    # within rounding distance of their TPReLU kink take the other branch ("TPReLU mask flips");
    # such an element moves its own gradients by (1 - a) * dout and nothing else: bound the flips, hold the rest
    # to rounding.
    for u, v in zip(a[1:], b[1:]):
        d = (u - v).abs() / v.abs().max()
        assert d.max().item() <= 5e-2
        assert (d > 2e-5).float().mean().item() <= 2e-2


def test_planes_only_activations_in_chains():
    """Inside a fused chain a tensor-core layer does not write the fp32 copy of an activation whose consumer
    reads bf16 planes; the chain's result and every gradient are those of the fully materialised chain, the
    elided tensor is marked, and anything that would read its fp32 values fails loudly."""
    from glis_b200 import _lib, ops
    _lib.set_precision("bf16x3")
    pm, pmod = _product()

    def build():
        torch.manual_seed(43)
        layers = [pmod.WeightNormalizedConv2d(32, 64, 4, 2, 1, scale=False, bias=False), pmod.TPReLU(64),
                  pmod.WeightNormalizedConv2d(64, 64, 4, 2, 1, scale=False, bias=False), pmod.TPReLU(64),
                  pmod.WeightNormalizedConv2d(64, 8, (5, 5))]
        return [m.to(DEV) for m in layers]

    def run(elide):
        layers = build()
        x = torch.randn(6, 32, 20, 20, device=DEV, requires_grad=True)
        if elide:
            y = pm.run_layers(layers, x)
        else:          # pair by pair: no pair sees its consumer, every activation is materialised
            y = x
            for i in (0, 2):
                y = pm.run_layers(layers[i:i + 2], y)
            y = layers[4](y)
        (y * torch.linspace(-1, 1, y.numel(), device=DEV).view_as(y)).sum().backward()
        return [y.detach(), x.grad] + [p.grad for m in layers for p in m.parameters()]

    for u, v in zip(run(True), run(False)):
        assert rel_err(u, v) <= 1e-6
    layers = build()
    x = torch.randn(6, 32, 20, 20, device=DEV)
    with torch.no_grad():
        mid = pm.run_layers(layers[:2], x, following=layers[2:])
        assert getattr(mid, "_glis_f32_invalid", False) and getattr(mid, "_glis_planes", None) is not None
        last = pm.run_layers(layers[2:4], mid, following=layers[4:])     # consumer = the 5x5 head: fp32 kernel
        assert not getattr(last, "_glis_f32_invalid", False)
        mid._glis_planes = None                  # without its planes the tensor is unusable, and says so
        with pytest.raises(RuntimeError, match="planes-only"):
            ops.planes_of(mid)
        with pytest.raises(RuntimeError, match="planes-only"):
            layers[4](mid)


def test_generator_tail_sigmoid_fused():
    """Last transposed convolution + nn.Sigmoid of the generators as one operator (sigmoid in the fold kernel /
    the contraction's epilogue) against the two modules run one by one."""
    pm, pmod = _product()
    for (cin, size) in ((64, 10), (32, 6)):
        torch.manual_seed(44)
        conv = pmod.WeightNormalizedConvTranspose2d(cin, 3, 4, 2, 1).to(DEV)
        sig = torch.nn.Sigmoid()
        with torch.no_grad():
            conv.bias.uniform_(-0.5, 0.5)
            conv.scale.uniform_(0.5, 2.0)
        res = []
        for fused in (True, False):
            x = torch.randn(4, cin, size, size, device=DEV, requires_grad=True)
            for p in conv.parameters():
                p.grad = None
            torch.manual_seed(45)
            x.data.normal_()
            y = pm.run_layers([conv, sig], x) if fused else sig(conv(x))
            (y * torch.linspace(-1, 1, y.numel(), device=DEV).view_as(y)).sum().backward()
            res.append([y.detach(), x.grad] + [p.grad.clone() for p in conv.parameters()])
        for u, v in zip(*res):
            assert rel_err(u, v) <= 2e-6


def test_lis_module_fused_matches_unfused(precision):
    """One LIS block (x + linear(TPReLU(linear(x)))) as the cluster kernel pair of csrc/lis.cu against the three
    modules run one by one: output, input gradient, both weight gradients, TPReLU parameter gradients; ragged
    batches (rows beyond the batch in the last 16-row tile) and the three cluster sizes 1 / 4 / 8."""
    from glis_b200 import ops
    pm, pmod = _product()
    for (b, code) in ((64, 256), (5, 32), (20, 128), (33, 256)):
        torch.manual_seed(46)
        block = pm.DottedSequential()
        block.add_module("lis.0-1.linear", pmod.WeightNormalizedLinear(code, code, init_factor=0.01, scale=False, bias=False))
        block.add_module("lis.0-1.act", pmod.TPReLU(code))
        block.add_module("lis.0-2.linear", pmod.WeightNormalizedLinear(code, code, init_factor=0.01, scale=False, bias=False))
        block = block.to(DEV)
        with torch.no_grad():
            block[1].weight.uniform_(-0.2, 1.2)
            block[1].bias.uniform_(-0.3, 0.3)
        res = []
        for fused in (True, False):
            ops.LIS_FUSED = fused
            try:
                for p in block.parameters():
                    p.grad = None
                torch.manual_seed(47)
                x = torch.randn(b, code, device=DEV, requires_grad=True)
                y = pm.lis_residual(block, x)
                (y * torch.linspace(-1, 1, y.numel(), device=DEV).view_as(y)).sum().backward()
                res.append([y.detach(), x.grad] + [p.grad.clone() for p in block.parameters()])
            finally:
                ops.LIS_FUSED = True
        for u, v in zip(*res):
            assert rel_err(u, v) <= 2e-5, (b, code)
        with torch.no_grad():
            ops.LIS_FUSED = True
            y_ng = pm.lis_residual(block, x.detach())
        assert rel_err(y_ng, res[0][0]) <= 1e-6


# ---------------------------------------------------------------- --ls, dropout, R-iterative as first-class paths
def test_lsgan_loss_kernel():
    from glis_b200 import ops
    lg = torch.randn(64, device=DEV) * 3
    for t in (0.0, 1.0):
        loss, dl, pr = ops.lsq_logits(lg, t, gscale=0.7, want_prob=True)
        l64 = lg.double().cpu().requires_grad_(True)
        ref = torch.nn.functional.mse_loss(torch.sigmoid(l64), torch.full((64,), t, dtype=torch.float64))
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
        assert rel_err(dl, 0.7 * l64.grad) <= 1e-5 and rel_err(pr, torch.sigmoid(l64)) <= 1e-6


def test_step_parity_lsgan():
    """One iteration with --ls against the fp64 oracle (losses 1e-4, gradients 1e-3, flip-aware)."""
    from glis_b200.trainer import GLISTrainer
    og, od, pg, pd = _make_pair(32, 32, 16, 3, 32, 2, seed=14)
    ot = GLISOracleTrainer(og, od, lr=2e-5, lambda_r=0.9, ls=True)
    pt = GLISTrainer(pg, pd, lr=2e-5, lambda_r=0.9, ls=True)
    fa = FlipAwarePair([(og, pg), (od, pd)])
    gen = torch.Generator().manual_seed(15)
    real, zd, zg = torch.rand(8, 3, 32, 32, generator=gen), torch.randn(8, 32, generator=gen), torch.randn(8, 32, generator=gen)
    with fa.tap():
        lp = pt.step(real.to(DEV), zd.to(DEV), zg.to(DEV), 1, 2)
    fa.feed()
    lo = ot.step(real.double(), zd.double(), zg.double(), 1, 2)
    fa.check(FLIP_TOL)
    fa.remove()
    for name in ("d_real", "d_fake", "g"):
        assert abs(lp[name].item() - lo[name]) <= FWD_TOL * abs(lo[name]), name
    for onet, flat in ((og, pt.gen_flat), (od, pt.dis_flat)):
        for (n, po), gp in zip(onet.named_parameters(), _flat_grads(flat)):
            go = po.grad if po.grad is not None else torch.zeros_like(po)
            if go.abs().max().item() > 0:
                assert rel_err(gp, go) <= GRAD_TOL, (n, rel_err(gp, go))


def test_dropout_kernel_matches_host_philox():
    """glis_dropout against the host restatement of its generator (tests/util.py::philox_keep_mask): element and
    channel mode, forward and backward share the mask, a tick of the device counter changes it."""
    from glis_b200 import ops
    from util import philox_keep_mask
    ops.DropoutClock.seed = 987654321
    for shape, chan in (((6, 16, 5, 5), False), ((3, 8, 4, 6), True), ((5, 33), False)):
        x = torch.randn(*shape, device=DEV)
        if len(shape) == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        p = 0.3
        ops.DropoutClock.tick(x.device)
        counter = int(ops.DropoutClock.counter(x.device).item())
        y = ops.dropout(x, p, channel_mode=chan)
        call = ops.DropoutClock.calls - 1
        g = torch.randn_like(y)
        y.backward(g)
        if chan:
            keep = philox_keep_mask(ops.DropoutClock.seed, counter + call, shape[0] * shape[1], p).view(shape[0], shape[1], 1, 1)
        elif len(shape) == 4:     # storage order is NHWC
            n, c, h, w = shape
            keep = philox_keep_mask(ops.DropoutClock.seed, counter + call, n * c * h * w, p).view(n, h, w, c).permute(0, 3, 1, 2)
        else:
            keep = philox_keep_mask(ops.DropoutClock.seed, counter + call, shape[0] * shape[1], p).view(shape)
        keep = keep.to(DEV).float()
        assert torch.equal(y.detach(), x.detach() * keep / (1 - p))
        assert torch.equal(x.grad, g * keep / (1 - p))
        assert 0.5 < keep.mean().item() < 0.9
        y2 = ops.dropout(x.detach(), p, channel_mode=chan)          # next call: another mask
        assert not torch.equal(y2, y.detach())


def test_discriminator_dropout_training_parity():
    """--d_dropout (common/model.py:52-53) in TRAINING mode: one whole iteration (BASELINE config 5b family:
    64x64-style 3-level nearest-upsampling G, dropout before D's last layer) against the fp64 oracle fed the masks
    the device generator drew — reproduced on the host from (seed, counter, call)."""
    pm, _ = _product()
    from glis_b200 import ops
    from glis_b200.trainer import GLISTrainer
    from util import philox_keep_mask
    W = H = 32; nf, nl, code, B, p = 16, 3, 32, 6, 0.3
    torch.manual_seed(81)
    og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "nearest")
    od = oracle.build_discriminator(W, H, nf, nl, "weight", p)
    pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "nearest")
    pd = pm.build_discriminator(W, H, nf, nl, "weight", p)
    copy_params(pg, og); copy_params(pd, od)
    og, od, pg, pd = og.double(), od.double(), pg.to(DEV), pd.to(DEV)
    ops.DropoutClock.seed = 13579
    ot = GLISOracleTrainer(og, od, lr=2e-5, lambda_r=0.9)
    pt = GLISTrainer(pg, pd, lr=2e-5, lambda_r=0.9)
    assert pt.dropout
    fa = FlipAwarePair([(og, pg), (od, pd)])
    gen = torch.Generator().manual_seed(82)
    real, zd, zg = torch.rand(B, 3, H, W, generator=gen), torch.randn(B, code, generator=gen), torch.randn(B, code, generator=gen)
    with fa.tap():
        lp = pt.step(real.to(DEV), zd.to(DEV), zg.to(DEV), 1, 1)
    fa.feed()
    counter = int(ops.DropoutClock.counter(torch.device(DEV, 0)).item())
    f, hh = nf * 4, H // 8                       # the tensor the dropout sees: (N, 4 nf, H/8, W/8), NHWC storage
    def mask(call, n):
        k = philox_keep_mask(ops.DropoutClock.seed, counter + call, n * f * hh * hh, p)
        return (k.view(n, hh, hh, f).permute(0, 3, 1, 2).double() / (1 - p))
    m_d, m_g = mask(0, 2 * B), mask(1, B)        # call 0: the 2B batch of the D update; call 1: the G update
    queue = [m_d[:B], m_d[B:], m_g]
    drop = od._modules["final\u00b7dropout"]
    drop.forward = lambda x: x * queue.pop(0)
    lo = ot.step(real.double(), zd.double(), zg.double(), 1, 1)
    del drop.__dict__["forward"]
    assert not queue
    fa.check(FLIP_TOL)
    fa.remove()
    for name in ("d_real", "d_fake", "g"):
        assert abs(lp[name].item() - lo[name]) <= FWD_TOL * abs(lo[name]), (name, lp[name].item(), lo[name])
    for onet, flat in ((og, pt.gen_flat), (od, pt.dis_flat)):
        for (n, po), gp in zip(onet.named_parameters(), _flat_grads(flat)):
            go = po.grad if po.grad is not None else torch.zeros_like(po)
            if go.abs().max().item() > 0:
                assert rel_err(gp, go) <= GRAD_TOL, (n, rel_err(gp, go))


def test_dropout_graph_replay_draws_fresh_masks():
    """Config 5b under CUDA-graph replay: the step with D dropout is captured once and every replay draws another
    mask (the device counter is advanced inside the graph)."""
    pm, _ = _product()
    from glis_b200 import ops
    from glis_b200.trainer import GLISTrainer, GraphedStep
    W = H = 32; nf, nl, code, B = 16, 3, 32, 8
    runs = []
    gen = torch.Generator().manual_seed(83)
    batches = [(torch.rand(B, 3, H, W, generator=gen).to(DEV), torch.randn(B, code, generator=gen).to(DEV),
                torch.randn(B, code, generator=gen).to(DEV)) for _ in range(3)]
    for graphed in (False, True):
        torch.manual_seed(84)
        g = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "nearest").to(DEV)
        d = pm.build_discriminator(W, H, nf, nl, "weight", 0.4).to(DEV)
        tr = GLISTrainer(g, d, lr=2e-5)
        ops.DropoutClock.seed = 2468
        ops.DropoutClock.counter(torch.device(DEV, 0)).zero_()
        ops.DropoutClock.calls = 0
        stepper = GraphedStep(tr, B, H, W, code, DEV, warmup=1) if graphed else tr
        losses = []
        for real, zd, zg in batches:
            o = stepper.step(real, zd, zg, 1, 1)
            losses.append([o[k].item() for k in ("d_real", "d_fake", "g")])
        runs.append(losses)
    eager, replay = runs
    assert len({tuple(l) for l in replay}) == 3                     # three different batches AND masks
    # the graph's counter starts one tick later (its warm-up iteration), so masks differ from the eager run's:
    # the losses agree statistically, not bitwise — what must hold is that replays are not frozen to one mask
    real, zd, zg = batches[0]
    a = [stepper.step(real, zd, zg, 1, 1)["d_real"].item() for _ in range(2)]
    assert a[0] != a[1]


@pytest.mark.parametrize("ls", [False, True])
def test_r_iterative_golden(golden_dir, precision, ls):
    """RIterTrainer against two outer iterations run on the reference's own builders with stock optimizers
    (tests/golden/riter_steps.npz); with ``ls`` only the machinery is exercised against the oracle instead."""
    pm, _ = _product()
    from glis_b200.trainer import RIterTrainer
    from oracle.step import riter_iteration
    s = _golden(golden_dir, "riter_steps.npz")
    cfg = _grp(s, "cfg")
    W, H, B, code, nf, nl, R = (int(cfg[k]) for k in ("W", "H", "B", "code", "nf", "nl", "R"))
    lr, lam = float(cfg["lr"]), float(cfg["lam"])
    nets, onets = {}, {}
    for tag, build_p, build_o in (
            ("g", lambda: pm.build_generator(W, H, nf, nl, code, "weight"), lambda: oracle.build_generator(W, H, nf, nl, code, "weight")),
            ("r", lambda: pm.build_reverser(W, H, nf // 2, nl, code, "weight", 0), lambda: oracle.build_reverser(W, H, nf // 2, nl, code, "weight", 0)),
            ("d", lambda: pm.build_discriminator(W, H, nf, nl, "weight", 0), lambda: oracle.build_discriminator(W, H, nf, nl, "weight", 0))):
        sd = {k: torch.from_numpy(v) for k, v in _grp(s, "init/" + tag).items()}
        nets[tag] = build_p()
        nets[tag].load_state_dict({k: v.float() for k, v in sd.items()})
        nets[tag] = nets[tag].to(DEV)
        onets[tag] = build_o().double()
        onets[tag].load_state_dict(sd)
    tr = RIterTrainer(nets["g"], nets["r"], nets["d"], lr=lr, lambda_r=lam, r_iterations=R, ls=ls)
    gs, rs, ds = {}, {}, {}
    f = lambda a: torch.from_numpy(a).float().to(DEV)
    for it, flags in enumerate(cfg["schedules"]):
        flags = [bool(v) for v in flags]
        g = _grp(s, "it%d" % it)
        reals = [g["real%d" % i] for i in range(sum(flags))]
        got = tr.step(f(g["z"]), [f(x) for x in reals], flags)
        want = None
        if ls:
            want = riter_iteration(onets["g"], onets["r"], onets["d"], gs, rs, ds, torch.from_numpy(g["z"]),
                                   [torch.from_numpy(x) for x in reals], lr, lam, R, flags, ls=True)
        ltol = FWD_TOL if it == 0 else (5e-4 if precision == "fp32" else 3e-3)
        for hop, rec in enumerate(got):
            assert (rec is None) == (not flags[hop])
            if rec is None:
                continue
            ref = want[hop] if ls else {k: float(v) for k, v in _grp(s, "it%d/hop%d" % (it, hop)).items()}
            assert sorted(rec) == sorted(ref)
            for k in rec:
                tol = ltol if hop == 0 else 10 * ltol       # later hops inherit the sign-like RMSprop drift (lr = 1e-2)
                assert abs(rec[k].item() - ref[k]) <= tol * abs(ref[k]) + 1e-7, (it, hop, k, rec[k].item(), ref[k])
        if not ls:
            ptol = 2e-2 if precision == "fp32" else 5e-2
            for tag in ("g", "r", "d"):
                for k, v in nets[tag].state_dict().items():
                    assert rel_err(v, torch.from_numpy(g[tag + "/" + k])) <= ptol, (it, tag, k)


def test_r_iterative_graph_replay_matches_eager(precision):
    """GraphedRIter (one captured graph per do_train schedule) follows the eager trajectory."""
    pm, _ = _product()
    from glis_b200.trainer import GraphedRIter, RIterTrainer
    W = H = 32; nf, nl, code, B, R = 16, 3, 32, 4, 2
    def make():
        torch.manual_seed(91)
        return RIterTrainer(pm.build_generator(W, H, nf, nl, code, "weight").to(DEV),
                            pm.build_reverser(W, H, nf // 2, nl, code, "weight", 0).to(DEV),
                            pm.build_discriminator(W, H, nf, nl, "weight").to(DEV), lr=2e-5, r_iterations=R)
    eager, graphed = make(), GraphedRIter(make(), B, H, W, code, DEV)
    gen = torch.Generator().manual_seed(92)
    for it, flags in enumerate([[True, True, True], [False, True, True], [True, True, True], [False, False, True]]):
        z = torch.randn(B, code, generator=gen).to(DEV)
        reals = [torch.rand(B, 3, H, W, generator=gen).to(DEV) for _ in range(sum(flags))]
        a = eager.step(z, reals, flags)
        b = graphed.step(z, reals, flags)
        for hop, (ra, rb) in enumerate(zip(a, b)):
            assert (ra is None) == (rb is None)
            for k in (ra or {}):
                assert abs(ra[k].item() - rb[k].item()) <= 2e-3 * abs(ra[k].item()) + 1e-7, (it, hop, k)
    for fa_, fb_ in ((eager.gen_flat, graphed.tr.gen_flat), (eager.rev_flat, graphed.tr.rev_flat), (eager.dis_flat, graphed.tr.dis_flat)):
        assert (fa_.p - fb_.p).abs().max().item() <= 4 * 3 * 6.4 * 2e-5 + 1e-7


def test_two_backward_passes_before_zero_grad_accumulate():
    """The reference's own D pattern — `loss_d_real.backward()` then `loss_d_fake.backward()` into the same
    gradients (g_lis/main.py:555-565) — on FlatParams-owned weights: the raw filter-gradient scratch the kernels
    add into must not carry the first pass into the second one's projection."""
    pm, _ = _product()
    from glis_b200.trainer import FlatParams
    torch.manual_seed(95)
    od = oracle.build_discriminator(32, 32, 16, 3, "weight", 0)
    pd = pm.build_discriminator(32, 32, 16, 3, "weight", 0)
    copy_params(pd, od)
    od, pd = od.double(), pd.to(DEV)
    flat = FlatParams(pd)
    gen = torch.Generator().manual_seed(96)
    xs = [torch.rand(4, 3, 32, 32, generator=gen) for _ in range(2)]
    flat.zero_grad()
    for t, x in zip((1.0, 0.0), xs):
        p = pd(x.to(DEV))
        torch.nn.functional.binary_cross_entropy(p, torch.full_like(p, t)).backward()
        q = od(x.double())
        torch.nn.functional.binary_cross_entropy(q, torch.full_like(q, t)).backward()
    flat.rebind_grads()
    torch.cuda.synchronize()
    for (n, po), gp in zip(od.named_parameters(), _flat_grads(flat)):
        assert rel_err(gp, po.grad) <= 2e-2, (n, rel_err(gp, po.grad))     # (not flip-aware: a loose bound suffices —
        # stale scratch would double the first pass's contribution, an error of order 1)


def test_r_separate_golden(golden_dir, precision):
    """RSeparateTrainer (g_lis/train_r.py:406-436: R trained alone against frozen G-LIS + D) against two iterations
    run on the reference's own builders with a stock RMSprop (tests/golden/rsep_steps.npz)."""
    pm, _ = _product()
    from glis_b200.trainer import RSeparateTrainer
    s = _golden(golden_dir, "rsep_steps.npz")
    cfg = _grp(s, "cfg")
    W, H, B, code, nf, nl, n_lis = (int(cfg[k]) for k in ("W", "H", "B", "code", "nf", "nl", "n_lis"))
    nets = {"g": pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis, "fractional"),
            "r": pm.build_reverser(W, H, nf // 2, nl, code, "weight", 0),
            "d": pm.build_discriminator(W, H, nf, nl, "weight", 0)}
    for tag, net in nets.items():
        net.load_state_dict({k: torch.from_numpy(v).float() for k, v in _grp(s, "init/" + tag).items()})
        nets[tag] = net.to(DEV)
    tr = RSeparateTrainer(nets["g"], nets["r"], nets["d"], lr=float(cfg["lr"]), r_iterations=n_lis)
    g0 = [p.detach().clone() for p in nets["g"].parameters()]
    for it in range(int(cfg["iters"])):
        g = _grp(s, "it%d" % it)
        out = tr.step(torch.from_numpy(g["z"]).float().to(DEV))
        ltol = FWD_TOL if it == 0 else (5e-4 if precision == "fp32" else 3e-3)
        assert abs(out["stage1"].item() - float(g["stage1"])) <= FWD_TOL * abs(float(g["stage1"]))   # frozen nets
        assert abs(out["r"].item() - float(g["r"])) <= ltol * abs(float(g["r"])), (it, out["r"].item(), float(g["r"]))
        assert abs(out["stage2"].item() - float(g["stage2"])) <= 10 * ltol * abs(float(g["stage2"]))
        ptol = 1e-2 if precision == "fp32" else 5e-2
        for k, v in nets["r"].state_dict().items():
            assert rel_err(v, torch.from_numpy(g["r/" + k])) <= ptol, (it, k)
    for a, b in zip(g0, nets["g"].parameters()):
        assert torch.equal(a, b)              # the generator stays frozen


def test_cli_train_r(tmp_path, precision):
    """g_lis/train_r.py: loads a G-LIS / D checkpoint written by g_lis/main.py, trains R alone, saves the
    reference's file names (`net_archive/{prefix}_{r,r_opt,state}.pt`)."""
    import importlib.util
    from conftest import PKG
    def load(name, *parts):
        spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, *parts))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    glis, train_r = load("glis_main_gpu4", "g_lis", "main.py"), load("train_r_gpu", "g_lis", "train_r.py")
    base = str(tmp_path / "gan")
    shared = ["--synthetic", "--image_size", "32", "--nfeature", "16", "--code_size", "32", "--norm", "weight",
              "--r_iterations", "1", "--batch_size", "8", "--lr", "0.0002", "--precision", precision]
    glis.main(shared + ["--niter", "3", "--save_path", base, "--vis_interval", "100", "--save_interval", "100"])
    out = str(tmp_path / "r")
    train_r.main(shared + ["--niter", "4", "--load_path", base, "--save_path_r", out, "--vis_interval", "2",
                           "--vis_size", "2", "--save_interval", "2"])
    arch = os.path.join(out, "net_archive")
    for f in ("2_r.pt", "2_r_opt.pt", "2_state.pt", "4_r.pt"):
        assert os.path.exists(os.path.join(arch, f)), f
    assert os.path.exists(os.path.join(out, "samples_both", "sample_2_both.jpg"))
    assert torch.load(os.path.join(arch, "4_state.pt"), weights_only=False)["current_iter"] == 4


# ---------------------------------------------------------------- forward-only consumers
def _load_cli(name, *parts):
    import importlib.util
    from conftest import PKG
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, *parts))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reconstruction_test_matches_oracle():
    """`test()` of g_lis/main.py:398-453 — held-out images embedded by 50 RMSprop steps on the latent code with G
    frozen: G forward + data-gradient-only backward.  The same function on the product's generator (GPU) and on
    the oracle's (CPU, fp64) must follow the same trajectory."""
    import argparse
    pm, _ = _product()
    m = _load_cli("glis_main_f3", "g_lis", "main.py")
    torch.manual_seed(101)
    og = oracle.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 2, "fractional")
    pg = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 2, "fractional")
    copy_params(pg, og)
    targets = torch.rand(6, 3, 32, 32)
    opt = argparse.Namespace(batch_size=4, code_size=32, test_lr=0.01, test_steps=8)
    want = m.reconstruction_test(og.double(), targets.double(), opt)
    got = m.reconstruction_test(pg.to(DEV), targets.to(DEV), opt)
    assert abs(got - want) <= 1e-3 * abs(want), (got, want)
    assert all(p.requires_grad for p in pg.parameters()) and pg.training


def test_samplers_match_oracle(tmp_path):
    """g_lis/sample_images.py: image grids per LIS depth (with an R-separate repair), interpolations, perturbations
    and the embedding of given images — the product's generator / reverser against the oracle's through the SAME
    sampler functions (uint8 images: at most one grey level apart; embedded codes: 1e-3), and the command line."""
    pm, _ = _product()
    S = _load_cli("glis_sample_images", "g_lis", "sample_images.py")
    torch.manual_seed(102)
    og = oracle.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 2, "fractional").eval()
    orv = oracle.build_reverser(32, 32, 8, 3, 32, "weight", 0).eval()
    pg = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 2, "fractional").eval()
    prv = pm.build_reverser(32, 32, 8, 3, 32, "weight", 0).eval()
    copy_params(pg, og); copy_params(prv, orv)
    pg, prv = pg.to(DEV), prv.to(DEV)
    code = torch.randn(10, 32)

    def close(a, b):
        d = (a.cpu().int() - b.cpu().int()).abs()
        return d.max().item() <= 1 and (d > 0).float().mean().item() < 0.02      # rounding to grey levels
    for depth in (0, 1, 2):
        a, af = S.generate_images(pg, prv, code.to(DEV), depth, 4)
        b, bf = S.generate_images(og, orv, code, depth, 4)
        assert a.shape == (10, 32, 32, 3) and close(a, b) and close(af, bf), depth
        assert close(S.generate_interpolations(pg, code[0].to(DEV), code[1].to(DEV), 8, depth),
                     S.generate_interpolations(og, code[0], code[1], 8, depth))
        assert close(S.generate_perturbations(pg, code[2].to(DEV), 7, 1.0, 9, depth),
                     S.generate_perturbations(og, code[2], 7, 1.0, 9, depth))
    real = torch.rand(3, 3, 32, 32)
    ca = S.embed_real_images(pg, prv, real.to(DEV), lr=1e-2, test_steps=6)
    cb = S.embed_real_images(og, orv, real, lr=1e-2, test_steps=6)
    # (Adam's step is lr * m / (sqrt(v) + 1e-8): sign-like where |g| is tiny, so rounding-level gradient differences
    #  move single elements by up to lr per step; the codes agree to a few 1e-3 of their scale, not to 1e-6)
    assert rel_err(ca, cb) <= 2e-2 and rel_l2(ca, cb) <= 5e-3
    gp, rp = str(tmp_path / "g.pt"), str(tmp_path / "r.pt")
    torch.save(pg.state_dict(), gp); torch.save(prv.state_dict(), rp); torch.save(real, str(tmp_path / "real.pt"))
    out = str(tmp_path / "samples")
    S.main(["--image_size", "32", "--nfeature", "16", "--code_size", "32", "--norm", "weight", "--r_iterations", "2",
            "--load_path_g", gp, "--load_path_r", rp, "--save_path", out, "--rounds", "1", "--with_real_images",
            "--real_images", str(tmp_path / "real.pt"), "--embed_steps", "3"])
    for f in ("sampled_images_r2/r2_full_0000.jpg", "sampled_images_rsep_r1_both/rsep_r1_chain_full_0000.jpg",
              "sampled_images_chains/chain_small_0000.jpg", "sampled_images_interpolations_all/interp_all_0000.jpg",
              "sampled_images_perturbations_r0/pert_r0_0000.jpg", "sampled_images_real_images_perturbations/pert_real_0002.jpg"):
        assert os.path.exists(os.path.join(out, f)), f


# ---------------------------------------------------------------- input pipeline
def _augment_reference(x, q):
    """The arithmetic of glis_augment restated with torch gathers (fp32): inverse affine map + bilinear sampling with
    constant / symmetric border, horizontal flip of the output, multiply, contrast about 0.5, clamp (no noise)."""
    n, c, h, w = x.shape
    out = torch.empty(n, c, h, w)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    for i in range(n):
        a = q[i]
        fx = (w - 1 - xs) if a[9] != 0 else xs
        sx, sy = a[0] * fx + a[1] * ys + a[2], a[3] * fx + a[4] * ys + a[5]
        x0, y0 = torch.floor(sx), torch.floor(sy)
        wx, wy = sx - x0, sy - y0

        def fetch(yy, xx):
            yy, xx = yy.long(), xx.long()
            if a[10] != 0:
                xx = xx % (2 * w); xx = torch.where(xx >= w, 2 * w - 1 - xx, xx)
                yy = yy % (2 * h); yy = torch.where(yy >= h, 2 * h - 1 - yy, yy)
                return x[i][:, yy, xx]
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            return x[i][:, yy.clamp(0, h - 1), xx.clamp(0, w - 1)] * ok
        v = (1 - wy) * ((1 - wx) * fetch(y0, x0) + wx * fetch(y0, x0 + 1)) + wy * ((1 - wx) * fetch(y0 + 1, x0) + wx * fetch(y0 + 1, x0 + 1))
        out[i] = (0.5 + a[7] * (v * a[6] - 0.5)).clamp(0, 1)
    return out


def test_augment_kernel_matches_torch_restatement():
    import random
    from glis_b200 import data
    g = torch.Generator().manual_seed(201)
    x = torch.rand(6, 3, 20, 28, generator=g)
    for name in ("none", "flowers102", "cifar10", "10kcats", "lsun_churches"):
        q = data.augment_params(name, 6, 20, 28, random.Random(5))
        sigma = q[:, 8].clone()
        q[:, 8] = 0                      # noise apart: compared exactly
        got = data.augment(x.to(DEV), q.to(DEV), seed=1)
        assert got.is_contiguous(memory_format=torch.channels_last)
        assert rel_err(got, _augment_reference(x, q)) <= 2e-6, name
        if name == "none":               # only the flip: every image is the input or its mirror
            for i in range(6):
                assert torch.equal(got[i].cpu(), x[i].flip(-1) if q[i, 9] else x[i])
        if sigma.max() > 0:              # additive Gaussian noise: one draw per pixel, shared by the channels
            q2 = torch.zeros(6, 12); q2[:, 0] = q2[:, 4] = q2[:, 6] = q2[:, 7] = 1; q2[:, 8] = 0.03
            y = data.augment(torch.full((6, 3, 64, 64), 0.5, device=DEV), q2.to(DEV), seed=3).cpu() - 0.5
            assert abs(y.std().item() - 0.03) < 2e-3 and abs(y.mean().item()) < 1e-3
            assert torch.equal(y[:, 0], y[:, 1]) and torch.equal(y[:, 0], y[:, 2])
            y3 = data.augment(torch.full((6, 3, 64, 64), 0.5, device=DEV), q2.to(DEV), seed=4).cpu() - 0.5
            assert not torch.equal(y, y3)


def test_prefetch_loader_order_sharding_and_position():
    """The input pipeline delivers the reference's sample order (a permutation consumed front to back, redrawn at the
    wrap) from worker processes, shards by rank, and reports the position of the batch the trainer holds."""
    from glis_b200.data import PrefetchLoader, ShuffledOrder
    n, B = 37, 8
    images = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 3, 4, 6).contiguous() / 100.0   # image k is all k/100
    ds = torch.utils.data.TensorDataset(images, torch.zeros(n))
    train_index = torch.arange(n)
    for rank, world in ((0, 1), (1, 2)):
        mine = train_index[rank::world]
        ref = ShuffledOrder(len(mine), B, 9 * 7919 + rank)
        want = iter(ref)
        loader = PrefetchLoader(ds, train_index, B, torch.device(DEV, 0), rank, world, workers=2, augment_set="none", seed=9)
        pos0 = loader.position()
        assert pos0["current_sample"] == 0 and sorted(pos0["index_shuffle"].tolist()) == list(range(len(mine)))
        for it in range(9):                       # > one epoch of this rank's share: crosses the reshuffle
            batch = loader.next_batch()
            ids = (batch[:, 0, 0, 0] * 100).round().long().cpu().tolist()
            assert ids == [int(mine[k]) for k in next(want)], (rank, it)
            pos = loader.position()
            assert pos["current_sample"] == ref.log[it][1] and torch.equal(pos["index_shuffle"], ref.log[it][0])
        # resume from the reported position: the stream continues where it stopped
        resumed = PrefetchLoader(ds, train_index, B, torch.device(DEV, 0), rank, world, workers=0, augment_set="none",
                                 seed=9, shuffle=pos["index_shuffle"], current=pos["current_sample"])
        ids = (resumed.next_batch()[:, 0, 0, 0] * 100).round().long().cpu().tolist()
        perm, cur = pos["index_shuffle"], pos["current_sample"]
        assert ids[:min(B, len(mine) - cur)] == [int(mine[int(perm[cur + k])]) for k in range(min(B, len(mine) - cur))]
        del loader, resumed


def test_cli_trains_from_an_image_folder_with_augmentation(tmp_path, precision):
    """g_lis/main.py on a real (tiny) image folder: `data_index.pt` split, decoder workers, device-side
    `--augment flowers102`, checkpoint with the data position, resume."""
    if precision == "fp32":
        pytest.skip("host-side test: once is enough")
    from PIL import Image
    root = tmp_path / "data"
    (root / "cls").mkdir(parents=True)
    g = torch.Generator().manual_seed(7)
    for i in range(40):
        arr = (torch.rand(40, 48, 3, generator=g) * 255).to(torch.uint8).numpy()
        Image.fromarray(arr).save(str(root / "cls" / ("img%03d.png" % i)))
    perm = torch.randperm(40, generator=g)
    torch.save({"running_test": perm[:4], "final_test": perm[4:8], "train": perm[8:]}, str(root / "data_index.pt"))
    m = _load_cli("glis_main_data", "g_lis", "main.py")
    save = str(tmp_path / "exp")
    common = ["--dataset", "folder", "--dataroot", str(root), "--crop_size", "40", "--image_size", "32", "--nfeature", "16",
              "--code_size", "32", "--norm", "weight", "--r_iterations", "1", "--batch_size", "8", "--lr", "0.0002",
              "--vis_interval", "100", "--save_interval", "100", "--test_interval", "1000", "--vis_size", "2",
              "--augment", "flowers102", "--workers", "2"]
    m.main(common + ["--niter", "6", "--save_path", save])
    state = torch.load(os.path.join(save, "net_archive", "last_state.pt"), weights_only=False)
    assert state["current_iter"] == 6 and state["index_shuffle"].numel() == 32
    assert state["current_sample"] == (6 * 8) % 32          # 48 samples drawn from a 32-image training set
    m.main(common + ["--niter", "8", "--load_path", save])
    state = torch.load(os.path.join(save, "net_archive", "last_state.pt"), weights_only=False)
    assert state["current_iter"] == 8 and state["current_sample"] == (8 * 8) % 32
    with pytest.raises(Exception):
        m.main(common + ["--niter", "1", "--save_path", save, "--augment", "nonsense"])


def test_tensor_core_weight_gradients_are_bit_reproducible(monkeypatch):
    """tcgen05 weight gradients split the pixel contraction over many CTAs (36 for D's level 1 at the 2B batch, 147
    for the image-side level 0).  Every split stores its partial sums into its own slab and the weight-norm projection
    adds the slabs in a fixed order (glis_conv_wgrad_bf16_slabs / glis_wn_project_slabs): two runs on the same inputs
    give bit-identical weight gradients (with atomically added partial sums the last bits changed from run to run)."""
    import ctypes as C
    from glis_b200 import _lib, ops
    _lib.set_precision("bf16x3")
    _, pmod = _product()
    monkeypatch.setattr(ops, "DETERMINISTIC_WGRAD", True)      # opt-in (GLIS_DETERMINISTIC_WGRAD=1): +3 % per iteration
    for ci, co, h, n in ((64, 128, 40, 128), (3, 64, 80, 128), (256, 512, 10, 64)):
        torch.manual_seed(301)
        conv = pmod.WeightNormalizedConv2d(ci, co, 4, 2, 1, scale=False, bias=False).to(DEV)
        x = torch.rand(n, ci, h, h, device=DEV)
        r = torch.randn(n, co, h // 2, h // 2, device=DEV)
        grads = []
        for _ in range(3):
            conv.weight.grad = None
            (conv(x) * r).sum().backward()
            grads.append(conv.weight.grad.clone())
        assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2]), (ci, co)
        # ... and equal to the default (atomic) form up to summation order
        monkeypatch.setattr(ops, "DETERMINISTIC_WGRAD", False)
        conv.weight.grad = None
        (conv(x) * r).sum().backward()
        assert rel_err(conv.weight.grad, grads[0]) <= 1e-5
        monkeypatch.setattr(ops, "DETERMINISTIC_WGRAD", True)
        if ci >= 32:
            g = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1)).geom(_lib.CONV, n, h, h, ci, h // 2, h // 2, co)
            assert _lib.load().glis_wgrad_tc_splits(C.byref(g)) > 1          # the contraction really is split
