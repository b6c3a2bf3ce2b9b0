#!/usr/bin/env python
"""Timeline of CTA 0 of the halo kernel (csrc/tc_conv2.cu, GLIS_TC_TRACE): when each weight tile / pixel box was
requested and when the MMA warp saw it."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

dev = "cuda"
spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
for name, rel, n, hi, wi, ci, ho, wo, co in [("D1 2B", L.CONV, 128, 40, 40, 64, 20, 20, 128),
                                              ("D2 2B", L.CONV, 128, 20, 20, 128, 10, 10, 256)]:
    g = spec.geom(rel, n, hi, wi, ci, ho, wo, co)
    x = torch.randn(n, hi, wi, ci, device=dev)
    xh, xl = ops.split_bf16(x)
    w = torch.randn(16, co, ci, device=dev) * 0.05
    wh, wl = ops.split_bf16(w)
    out = torch.empty(n, ho, wo, co, device=dev)
    ep = L.Epilogue(None, 0, None, None, None)
    trace = torch.zeros(1280, dtype=torch.int64, device=dev)
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
    t0 = t[1216]
    rel_ = lambda a: [v - t0 for v in a if v]
    print(name)
    print("  W issue  :", rel_(t[0:512])[:40])
    print("  W seen   :", rel_(t[512:1024])[:40])
    print("  X issue  :", rel_(t[1024:1088])[:12])
    print("  X seen   :", rel_(t[1088:1152])[:12])
    print("  epilogue :", rel_(t[1152:1216]))
