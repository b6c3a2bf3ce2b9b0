"""Replays tests/test_gpu_parity.py::test_step_parity_cfg1_shape[fp32] and prints the structure of the
worst gradient error per tensor (a TPReLU mask flip shows as ONE row / channel off, everything else exact)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, oracle
from oracle.step import GLISOracleTrainer
import common.model as pm
from glis_b200 import _lib
from glis_b200.trainer import GLISTrainer
from util import copy_params, rel_err

_lib.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
W = H = 32; nf = 64; nl = 3; code = 256; B = 32
torch.manual_seed(11)
og = oracle.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
od = oracle.build_discriminator(W, H, nf, nl, "weight", 0)
pg = pm.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", 1, "fractional")
pd = pm.build_discriminator(W, H, nf, nl, "weight", 0)
copy_params(pg, og); copy_params(pd, od)
og, od, pg, pd = og.double(), od.double(), pg.cuda(), pd.cuda()
ot = GLISOracleTrainer(og, od, lr=2e-5, lambda_r=0.9)
pt = GLISTrainer(pg, pd, lr=2e-5, lambda_r=0.9)
gen = torch.Generator().manual_seed(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
SUMMARY = len(sys.argv) > 3
for it, (kd, kg) in enumerate([(1, 1), (0, 1), (1, 0)] + ([(1, 0), (0, 0), (1, 0), (0, 0), (1, 1), (1, 0)] if SUMMARY else [])):
    worst = 0.0
    real = torch.rand(B, 3, H, W, generator=gen)
    zd, zg = torch.randn(B, code, generator=gen), torch.randn(B, code, generator=gen)
    lo = ot.step(real.double(), zd.double(), zg.double(), kd, kg)
    lp = pt.step(real.cuda(), zd.cuda(), zg.cuda(), kd, kg)
    for tag, onet, flat in (("gen", og, pt.gen_flat), ("dis", od, pt.dis_flat)):
        for (n, po), pp, o in zip(onet.named_parameters(), flat.params, flat.offsets):
            go = po.grad if po.grad is not None else torch.zeros_like(po)
            gp = flat.g[o:o + pp.numel()].view(pp.shape).double().cpu()
            e = (gp - go).abs()
            gm = go.abs().max().item()
            if gm == 0:
                continue
            r = e.max().item() / gm
            worst = max(worst, r)
            if r > 2e-4 and not SUMMARY:
                e2 = e.reshape(e.shape[0], -1)
                rows = e2.max(dim=1).values / gm
                top = torch.topk(rows, min(4, rows.numel()))
                print("it%d %s %-40s rel %.2e | rows>1e-4: %d of %d | top rows %s" % (
                    it, tag, n, r, int((rows > 1e-4).sum()), rows.numel(),
                    ["%d:%.1e" % (i, v) for v, i in zip(top.values.tolist(), top.indices.tolist())]))
    if SUMMARY:
        print("seed %s it%d depths (%d,%d): worst gradient rel err %.2e" % (sys.argv[2], it, kd, kg, worst))
    with torch.no_grad():
        for onet, flat, state in ((og, pt.gen_flat, ot.gen_state), (od, pt.dis_flat, ot.dis_state)):
            for po, pp, o in zip(onet.parameters(), flat.params, flat.offsets):
                pp.copy_(po.float())
                v = state.get(po)
                flat.v[o:o + po.numel()].copy_((v if v is not None else torch.zeros_like(po)).reshape(-1).float())
print("done")
