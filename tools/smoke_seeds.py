"""Worst gradient error of the smoke() configuration per seed (to pick one without a TPReLU mask flip)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
import oracle
from oracle.step import GLISOracleTrainer
import common.model as pm
from glis_b200.trainer import GLISTrainer

W = H = 32
B, nf, nl, code = 8, 16, 3, 32
for seed in range(8):
    torch.manual_seed(seed)
    og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
    od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
    pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
    pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
    pg.load_state_dict(og.state_dict())
    pd.load_state_dict(od.state_dict())
    og, od = og.double(), od.double()
    ot = GLISOracleTrainer(og, od, lr=1e-3)
    real, zd, zg = torch.rand(B, 3, H, W), torch.randn(B, code), torch.randn(B, code)
    lo = ot.step(real.double(), zd.double(), zg.double(), 1, 1)
    worsts = []
    for rep in range(3):
        pg2 = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
        pd2 = pm.build_discriminator(W, H, nf, nl, "weight", 0)
        pg2.load_state_dict(pg.state_dict())
        pd2.load_state_dict(pd.state_dict())
        pt = GLISTrainer(pg2.cuda(), pd2.cuda(), lr=1e-3)
        lp = pt.step(real.cuda(), zd.cuda(), zg.cuda(), 1, 1)
        torch.cuda.synchronize()
        worst = 0.0
        for p, o, po in zip(pt.gen_flat.params, pt.gen_flat.offsets, og.parameters()):
            got = pt.gen_flat.g[o:o + p.numel()].view(p.shape).cpu().double()
            want = po.grad if po.grad is not None else torch.zeros_like(po)
            worst = max(worst, ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item())
        worsts.append(worst)
    lerr = max(abs(lp[k].item() - lo[k]) / abs(lo[k]) for k in ("d_real", "d_fake", "g"))
    print("seed %d  worst gradient rel err over 3 runs: %s   loss err %.1e" % (seed, ["%.1e" % w for w in worsts], lerr))
