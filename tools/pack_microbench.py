#!/usr/bin/env python
"""Time glis_wn_prepare_bf16_perm (norm + bf16 hi/lo packs) per layer of config 2: wide-store kernel vs the element-wise
one (GLIS_PACK_WIDE=0), checking both packs against torch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L

CASES = [("conv 64->128", 0, 128, 64, 16, 0, 0), ("conv 128->256", 0, 256, 128, 16, 0, 0),
         ("conv 256->512", 0, 512, 256, 16, 0, 0), ("tconv 512->256", 1, 256, 512, 16, 0, 0),
         ("tconv 128->64", 1, 64, 128, 16, 0, 0), ("head linear 256->12800", 0, 12800, 256, 1, 512, 25)]
for name, axis, co, ci, t, pc, pp in CASES:
    shape = (co, ci, t) if axis == 0 else (ci, co, t)
    w = torch.randn(shape, device="cuda") * 0.05
    norm = torch.empty(co, device="cuda")
    mk = lambda *s: torch.empty(s, device="cuda", dtype=torch.bfloat16)
    fh, fl, bh, bl = mk(t, co, ci), mk(t, co, ci), mk(t, ci, co), mk(t, ci, co)
    c = 0.25 if axis == 1 else 1.0

    def run():
        L.call("glis_wn_prepare_bf16_perm", L.ptr(w), None, axis, co, ci, t, c, L.ptr(norm), L.ptr16(fh), L.ptr16(fl),
               L.ptr16(bh), L.ptr16(bl), pc, pp, L.stream(), kernels=2)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    wd = w.double()
    red = (1, 2) if axis == 0 else (0, 2)
    n = (wd.pow(2).sum(red) * c + 1e-6).sqrt()
    eff = wd / (n.view(-1, 1, 1) if axis == 0 else n.view(1, -1, 1))           # master layout
    oit = eff if axis == 0 else eff.permute(1, 0, 2)                            # [o][i][t]
    if pc:
        rows = torch.tensor([(o % pc) * pp + o // pc for o in range(co)], device="cuda")
        oit = oit[rows]
    want_f, want_b = oit.permute(2, 0, 1), oit.permute(2, 1, 0)
    ef = ((fh.double() + fl.double()) - want_f).abs().max().item() / want_f.abs().max().item()
    eb = ((bh.double() + bl.double()) - want_b).abs().max().item() / want_b.abs().max().item()
    print("%-26s %7.1f us per norm+pack  err fwd %.1e bwd %.1e" % (name, e0.elapsed_time(e1) * 50, ef, eb))
