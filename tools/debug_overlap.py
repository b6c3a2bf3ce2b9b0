"""Where do the side-stream and the single-stream schedules part?  Runs the same three iterations under
several schedules (twice each) and prints, per iteration, the largest difference of the flat gradient and
parameter buffers against the first single-stream run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
from glis_b200 import _lib, ops                      # noqa: E402
from glis_b200.trainer import GLISTrainer           # noqa: E402
from common import model as pm                      # noqa: E402

DEV = torch.device("cuda:0")
_lib.set_precision(sys.argv[1] if len(sys.argv) > 1 else "bf16x3")
_orig_refresh = ops.refresh_packs


def run(overlap, side_packs, iters=3):
    ops.Overlap.enabled = overlap
    ops.refresh_packs = (lambda flat, part="all", side=True: _orig_refresh(flat, part, side and side_packs))
    import glis_b200.trainer as T
    T.ops.refresh_packs = ops.refresh_packs
    torch.manual_seed(25)
    g = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 1, "fractional").to(DEV)
    d = pm.build_discriminator(32, 32, 16, 3, "weight", 0).to(DEV)
    tr = GLISTrainer(g, d, lr=1e-4)
    gen = torch.Generator().manual_seed(26)
    snaps = []
    for _ in range(iters):
        o = tr.step(torch.rand(8, 3, 32, 32, generator=gen).to(DEV), torch.randn(8, 32, generator=gen).to(DEV),
                    torch.randn(8, 32, generator=gen).to(DEV), 1, 1)
        torch.cuda.synchronize()
        snaps.append(dict(gg=tr.gen_flat.g.clone(), dg=tr.dis_flat.g.clone(), gp=tr.gen_flat.p.clone(),
                          dp=tr.dis_flat.p.clone(), loss=[o[k].item() for k in ("d_real", "d_fake", "g")]))
    names = [n for n, _ in g.named_parameters()]
    return snaps, tr, names


def worst(tr, names, a, b):
    """name of the generator parameter whose gradient differs most"""
    best = (0.0, None)
    for n, p, o in zip(names, tr.gen_flat.params, tr.gen_flat.offsets):
        d = (a[o:o + p.numel()] - b[o:o + p.numel()]).abs().max().item()
        s = b[o:o + p.numel()].abs().max().item()
        if d / (s + 1e-30) > best[0]:
            best = (d / (s + 1e-30), n)
    return best


base, tr0, names = run(False, False)
for label, (ov, sp) in (("single again", (False, False)), ("fork only", (True, False)), ("fork+packs", (True, True)),
                        ("fork+packs again", (True, True))):
    snaps, tr, _ = run(ov, sp)
    for i, (a, b) in enumerate(zip(snaps, base)):
        rel = lambda k: ((a[k] - b[k]).abs().max() / b[k].abs().max()).item()
        print("%-18s it%d  gen.g %.2e dis.g %.2e gen.p %.2e dis.p %.2e  worst gen grad %s  loss %s" % (
            label, i, rel("gg"), rel("dg"), rel("gp"), rel("dp"), worst(tr, names, a["gg"], b["gg"]), a["loss"]))
