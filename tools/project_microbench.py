#!/usr/bin/env python
"""Time glis_wn_project (weight-norm backward) and the norm kernel per layer of config 2, 20 launches per CUDA graph,
with the bytes each moves (G, w read; dw read + written when accumulating) -> achieved GB/s; checks against torch fp64."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L

CASES = [("D conv 64->128", 0, 128, 64, 16), ("D conv 128->256", 0, 256, 128, 16), ("D conv 256->512", 0, 512, 256, 16),
         ("D final conv 512x5x5->1", 0, 1, 512, 25), ("G tconv 512->256", 1, 256, 512, 16), ("G tconv 256->128", 1, 128, 256, 16),
         ("G tconv 128->64", 1, 64, 128, 16), ("G head linear 256->12800", 0, 12800, 256, 1), ("LIS linear 256->256", 0, 256, 256, 1)]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, axis, co, ci, t in CASES:
    shape = (co, ci, t) if axis == 0 else (ci, co, t)
    w = torch.randn(shape, device="cuda") * 0.05
    G = torch.randn(shape, device="cuda")
    scale = torch.rand(co, device="cuda") + 0.5
    c = 0.25 if axis == 1 else 1.0
    red = (1, 2) if axis == 0 else (0, 2)
    norm = (w.double().pow(2).sum(red) * c + 1e-6).sqrt().float().contiguous()
    dw = torch.zeros_like(w); ds = torch.zeros(co, device="cuda")
    nbytes = w.numel() * 4
    for acc in (0, 1):
        tt = timeit(lambda: L.call("glis_wn_project", L.ptr(G), L.ptr(w), L.ptr(scale), L.ptr(norm), axis, co, ci, t, c,
                                   L.ptr(dw), L.ptr(ds), acc, L.stream()))
        moved = nbytes * (4 if acc else 3)
        print("%-28s project%s %7.1f us  %6.0f GB/s" % (name, " (accumulate)" if acc else "             ", tt, moved / tt / 1e3))
    # check (overwrite mode)
    L.call("glis_wn_project", L.ptr(G), L.ptr(w), L.ptr(scale), L.ptr(norm), axis, co, ci, t, c, L.ptr(dw), L.ptr(ds), 0, L.stream())
    torch.cuda.synchronize()
    Gd, wd, nd, sd = G.double(), w.double(), norm.double(), scale.double()
    bshape = (-1, 1, 1) if axis == 0 else (1, -1, 1)
    dot = (Gd * wd).sum(red)
    want = (sd / nd).view(bshape) * (Gd - c * wd * (dot / nd ** 2).view(bshape))
    err = (dw.double() - want).abs().max().item() / want.abs().max().item()
    errs = (ds.double() - dot / nd).abs().max().item() / (dot / nd).abs().max().item()
    nrm = torch.empty(co, device="cuda")
    tn = timeit(lambda: L.call("glis_wn_prepare", L.ptr(w), None, axis, co, ci, t, c, L.ptr(nrm), None, None, L.stream()))
    print("%-28s norm                  %7.1f us  %6.0f GB/s   project err dw %.1e dscale %.1e" % (name, tn, nbytes / tn / 1e3, err, errs))
