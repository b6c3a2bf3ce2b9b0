import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, oracle
import common.model as pm
from glis_b200 import _lib
from glis_b200.trainer import GLISTrainer
from oracle.step import GLISOracleTrainer
from util import copy_params, rel_err

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
_lib.set_precision(mode)
W = H = 32; nf = 64; nl = 3; code = 256; B = 32
torch.manual_seed(11)
og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
copy_params(pg, og); copy_params(pd, od)
og, od = og.double(), od.double()
ot = GLISOracleTrainer(og, od, lr=2e-5); pt = GLISTrainer(pg.cuda(), pd.cuda(), lr=2e-5)
gen = torch.Generator().manual_seed(5)
real = torch.rand(B, 3, H, W, generator=gen); zd = torch.randn(B, code, generator=gen); zg = torch.randn(B, code, generator=gen)
lo = ot.step(real.double(), zd.double(), zg.double(), 1, 1)
lp = pt.step(real.cuda(), zd.cuda(), zg.cuda(), 1, 1)
print(mode, {k: (lp[k].item(), lo[k]) for k in ("d_real", "d_fake", "g")})
for tag, onet, flat in (("gen", og, pt.gen_flat), ("dis", od, pt.dis_flat)):
    for (n, po), pp, o in zip(onet.named_parameters(), flat.params, flat.offsets):
        gp = flat.g[o:o + pp.numel()].view(pp.shape)
        print("  %s %-40s rel_err %.3e |g|max %.3e" % (tag, n, rel_err(gp, po.grad), po.grad.abs().max().item()))
