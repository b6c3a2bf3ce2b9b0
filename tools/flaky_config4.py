"""Hunt for the intermittent config-4 loss deviation: same nets, same inputs, N product iterations per
schedule; prints the distribution of |g loss - oracle| / oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                                         # noqa: E402
from oracle.step import GLISOracleTrainer            # noqa: E402
from glis_b200 import _lib, ops                      # noqa: E402
from glis_b200.trainer import GLISTrainer            # noqa: E402
import common.model as pm                            # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
_lib.set_precision("bf16x3")
W = H = 160
nf, nl, code, B = 32, 5, 64, 2
torch.manual_seed(31)
og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
sg, sd = og.state_dict(), od.state_dict()
gen = torch.Generator().manual_seed(5)
real = torch.rand(B, 3, H, W, generator=gen)
zd, zg = torch.randn(B, code, generator=gen), torch.randn(B, code, generator=gen)
ot = GLISOracleTrainer(og.double(), od.double(), lr=2e-5, lambda_r=0.9)
lo = ot.step(real.double(), zd.double(), zg.double(), 1, 1)
print("oracle", {k: lo[k] for k in ("d_real", "d_fake", "g")})


def once():
    pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
    pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
    pg.load_state_dict(sg)
    pd.load_state_dict(sd)
    pt = GLISTrainer(pg.cuda(), pd.cuda(), lr=2e-5, lambda_r=0.9)
    lp = pt.step(real.cuda(), zd.cuda(), zg.cuda(), 1, 1)
    torch.cuda.synchronize()
    return [abs(lp[k].item() - lo[k]) / abs(lo[k]) for k in ("d_real", "d_fake", "g")]


for label, ov, sk, lf in (("default", True, True, True), ("single stream", False, True, True),
                         ("no split-K fwd", True, False, True), ("no fused LIS", True, True, False),
                         ("default again", True, True, True)):
    ops.Overlap.enabled, ops.SPLIT_K_FORWARD, ops.LIS_FUSED = ov, sk, lf
    errs = [once() for _ in range(N)]
    g = sorted(e[2] for e in errs)
    print("%-16s g: median %.1e max %.1e  over 1e-4: %d/%d   d_real max %.1e d_fake max %.1e" % (
        label, g[len(g) // 2], g[-1], sum(1 for v in g if v > 1e-4), N, max(e[0] for e in errs), max(e[1] for e in errs)))
ops.Overlap.enabled, ops.SPLIT_K_FORWARD, ops.LIS_FUSED = True, True, True
