#!/usr/bin/env python
"""Time the image-side 1x1 products (csrc/tc_pm.cu vs the channel-major kernel: GLIS_TC_PM=0) with their real
epilogues, CUDA events around a 20-launch graph; checks the result against a torch fp32 product."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

CASES = [  # name, images, h, w, K, N, epilogue
    ("D level 0, 2B images (unfold product)", 128, 40, 40, 48, 64, "tprelu+preact+planes"),
    ("D level 0, B images", 64, 40, 40, 48, 64, "tprelu+preact+planes"),
    ("D level 0, B images, no_grad", 64, 40, 40, 48, 64, "tprelu+planes"),
    ("G level 0 (fold product)", 64, 40, 40, 64, 48, "f32"),
    ("G level 0 K=128", 64, 40, 40, 128, 48, "f32"),
]


def main():
    dev = "cuda"
    spec = ops.ContractionSpec(False, (1, 1), (1, 1), (0, 0), (1, 1))
    torch.manual_seed(0)
    for name, n, h, w, k, co, epi in CASES:
        g = spec.geom(L.CONV, n, h, w, k, h, w, co)
        x = torch.randn(n, h, w, k, device=dev)
        wt = torch.randn(co, k, device=dev) * 0.1
        xh, xl = ops.split_bf16(x)
        wh, wl = ops.split_bf16(wt)
        a, b = torch.rand(co, device=dev), torch.randn(co, device=dev) * 0.1
        out = torch.empty(n, h, w, co, device=dev) if "f32" in epi else None
        pre = torch.empty(n, h, w, co, device=dev) if "preact" in epi else None
        hi = torch.empty(n, h, w, co, device=dev, dtype=torch.bfloat16) if "planes" in epi else None
        lo = torch.empty_like(hi) if hi is not None else None
        act = L.ACT_TPRELU if "tprelu" in epi else L.ACT_NONE
        ep = L.Epilogue(None, act, L.ptr(a), L.ptr(b), L.ptr(pre), None, None, 0)

        def run():
            L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xh), L.ptr16(xl), L.ptr16(wh), L.ptr16(wl),
                   C.byref(ep), L.ptr(out), L.ptr16(hi), L.ptr16(lo), L.PREC_BF16X3, L.stream())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                run()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        y = (x.double().view(-1, k) @ wt.double().t()).view(n, h, w, co)
        errs = []
        if act == L.ACT_TPRELU:
            t = y - b.double()
            want = torch.where(t > 0, t, a.double() * t) + b.double()
        else:
            want = y
        if pre is not None:
            errs.append(((pre - y).abs().max() / y.abs().max()).item())
        if out is not None:
            errs.append(((out - want).abs().max() / want.abs().max()).item())
        if hi is not None:
            errs.append((((hi.double() + lo.double()) - want).abs().max() / want.abs().max()).item())
        nbytes = x.numel() * 4 + sum(t.numel() * t.element_size() for t in (out, pre, hi, lo) if t is not None)
        us = e0.elapsed_time(e1) * 50
        print("%-42s %-22s %6.1f us  %5.2f TB/s  max rel err %.1e" % (name, epi, us, nbytes / us / 1e6, max(errs)))


if __name__ == "__main__":
    main()
