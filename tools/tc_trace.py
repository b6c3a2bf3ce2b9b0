#!/usr/bin/env python
"""Per-k-step timeline of CTA 0 of the tcgen05 conv kernel (GLIS_TC_TRACE)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

dev = "cuda"
spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
for name, rel, n, hi, wi, ci, ho, wo, co in [("D1", L.CONV, 64, 40, 40, 64, 20, 20, 128),
                                              ("G1", L.TCONV, 64, 20, 20, 128, 40, 40, 64),
                                              ("D3", L.CONV, 64, 10, 10, 256, 5, 5, 512)]:
    g = spec.geom(rel, n, hi, wi, ci, ho, wo, co)
    x = torch.randn(n, hi, wi, ci, device=dev)
    xh, xl = ops.split_bf16(x)
    w = torch.randn(16, co, ci, device=dev) * 0.05
    wh, wl = ops.split_bf16(w)
    out = torch.empty(n, ho, wo, co, device=dev)
    ep = L.Epilogue(None, 0, None, None, None)
    trace = torch.zeros(1088, dtype=torch.int64, device=dev)
    def run():
        L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xh), L.ptr16(xl), L.ptr16(wh), L.ptr16(wl),
               C.byref(ep), L.ptr(out), None, None, L.PREC_BF16X3, L.stream())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    os.environ["GLIS_TC_TRACE"] = "%x" % trace.data_ptr()
    run()
    torch.cuda.synchronize()
    del os.environ["GLIS_TC_TRACE"]
    t = trace.cpu().tolist()
    t0 = t[1087]
    prod = [v - t0 for v in t[:512] if v]
    mma = [v - t0 for v in t[512:1024] if v]
    epi = [v - t0 for v in t[1024:1087] if v]
    print(name, "k-steps", len(prod))
    print("  producer issue (ns):", prod[:40])
    print("  mma full-seen  (ns):", mma[:40])
    print("  epilogue start/end (ns):", epi)
