#!/usr/bin/env python
"""One tensor-core conv launch shape in a loop (for `ncu --set full -k regex:tc_conv`):
    python tools/halo_one.py [D1|D2|D3|G1] [reps]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

SH = {"D1": (L.CONV, 128, 40, 40, 64, 20, 20, 128), "D2": (L.CONV, 128, 20, 20, 128, 10, 10, 256),
      "D3": (L.CONV, 128, 10, 10, 256, 5, 5, 512), "G1": (L.TCONV, 64, 20, 20, 128, 40, 40, 64)}
rel, n, hi, wi, ci, ho, wo, co = SH[sys.argv[1] if len(sys.argv) > 1 else "D1"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
g = spec.geom(rel, n, hi, wi, ci, ho, wo, co)
x = torch.randn(n, hi, wi, ci, device="cuda")
xh, xl = ops.split_bf16(x)
w = torch.randn(16, co, ci, device="cuda") * 0.05
wh, wl = ops.split_bf16(w)
out = torch.empty(n, ho, wo, co, device="cuda")
ep = L.Epilogue(None, 0, None, None, None)
for _ in range(reps):
    L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xh), L.ptr16(xl), L.ptr16(wh), L.ptr16(wl), C.byref(ep),
           L.ptr(out), None, None, L.PREC_BF16X3, L.stream())
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
