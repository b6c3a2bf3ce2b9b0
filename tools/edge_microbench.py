#!/usr/bin/env python
"""GPU time of the thin fp32 kernels per shape: 20 launches replayed as one CUDA graph."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

dev = "cuda"


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


def conv(name, rel, n, hi, wi, ci, ho, wo, co, k, s, p, act=0, preact=False, planes=False):
    spec = ops.ContractionSpec(False, (k, k), (s, s), (p, p), (1, 1))
    g = spec.geom(rel, n, hi, wi, ci, ho, wo, co)
    x = torch.randn(n, hi, wi, ci, device=dev)
    w = torch.randn(k * k, ci, co, device=dev) * 0.05
    out = torch.empty(n, ho, wo, co, device=dev)
    pre = torch.empty_like(out) if preact else None
    hi_ = torch.empty_like(out, dtype=torch.bfloat16) if planes else None
    lo_ = torch.empty_like(out, dtype=torch.bfloat16) if planes else None
    a = torch.full((co,), 0.25, device=dev); b = torch.zeros(co, device=dev)
    ep = L.Epilogue(None, act, L.ptr(a), L.ptr(b), L.ptr(pre), L.ptr16(hi_), L.ptr16(lo_))
    t = timeit(lambda: L.call("glis_conv_forward", C.byref(g), L.ptr(x), L.ptr(w), C.byref(ep), L.ptr(out), 0, L.stream()))
    print("%-44s %7.1f us" % (name, t))


def wgrad(name, n, hi, wi, cb, ho, wo, ca, k, s, p):
    spec = ops.ContractionSpec(False, (k, k), (s, s), (p, p), (1, 1))
    g = spec.geom(L.CONV, n, hi, wi, cb, ho, wo, ca)
    small = torch.randn(n, ho, wo, ca, device=dev); big = torch.randn(n, hi, wi, cb, device=dev)
    G = torch.zeros(ca, cb, k, k, device=dev)
    t = timeit(lambda: L.call("glis_conv_wgrad", C.byref(g), L.ptr(small), L.ptr(big), L.ptr(G), 0, L.stream()))
    print("%-44s %7.1f us" % (name, t))


conv("LIS linear 64x256x256 (+tprelu,preact)", L.CONV, 64, 1, 1, 256, 1, 1, 256, 1, 1, 0, act=1, preact=True)
conv("LIS linear 64x256x256", L.CONV, 64, 1, 1, 256, 1, 1, 256, 1, 1, 0)
conv("G initial linear 64x12800x256", L.CONV, 64, 1, 1, 256, 1, 1, 12800, 1, 1, 0)
conv("G initial linear dgrad 64x256x12800", L.TCONV, 64, 1, 1, 12800, 1, 1, 256, 1, 1, 0)
conv("D head 128x1x12800", L.CONV, 128, 5, 5, 512, 1, 1, 1, 5, 1, 0)
conv("D head dgrad 128x12800x1", L.TCONV, 128, 1, 1, 1, 5, 5, 512, 5, 1, 0)
conv("D0 conv 3->64 80->40 2B (+tprelu,preact,planes)", L.CONV, 128, 80, 80, 3, 40, 40, 64, 4, 2, 1, act=1, preact=True, planes=True)
conv("D0 conv 3->64 80->40 B", L.CONV, 64, 80, 80, 3, 40, 40, 64, 4, 2, 1, act=1, preact=True, planes=True)
conv("G0 tconv 64->3 40->80 (sigmoid)", L.TCONV, 64, 40, 40, 64, 80, 80, 3, 4, 2, 1, act=2)
conv("dD0 tconv 64->3 40->80", L.TCONV, 64, 40, 40, 64, 80, 80, 3, 4, 2, 1)
conv("dG0 conv 3->64 80->40", L.CONV, 64, 80, 80, 3, 40, 40, 64, 4, 2, 1)
wgrad("D0 wgrad small=dy64 big=x3 2B", 128, 80, 80, 3, 40, 40, 64, 4, 2, 1)
wgrad("G0 wgrad small=x64 big=dy3", 64, 80, 80, 3, 40, 40, 64, 4, 2, 1)
wgrad("linear wgrad 12800x256 K=64", 64, 1, 1, 256, 1, 1, 12800, 1, 1, 0)
wgrad("LIS wgrad 256x256 K=64", 64, 1, 1, 256, 1, 1, 256, 1, 1, 0)


def wgrad_project(name, m, ca, cb, perm=(0, 0)):
    """linear weight gradient + weight-norm projection as one kernel (glis_linear_wgrad_project)"""
    dy = torch.randn(m, ca, device=dev); x = torch.randn(m, cb, device=dev)
    w = torch.randn(ca, cb, device=dev) * 0.05
    scale = torch.ones(ca, device=dev); norm = w.norm(dim=1).contiguous()
    dw = torch.zeros(ca, cb, device=dev); ds = torch.zeros(ca, device=dev)
    for acc in (0, 1):
        t = timeit(lambda: L.call("glis_linear_wgrad_project", L.ptr(dy), L.ptr(x), L.ptr(w), L.ptr(scale), L.ptr(norm),
                                  L.ptr(dw), L.ptr(ds), m, ca, cb, perm[0], perm[1], acc, 0, ca, L.stream()))
        print("%-44s %7.1f us" % (name + (" (accumulate)" if acc else ""), t))


wgrad_project("wgrad+project 12800x256 K=64", 64, 12800, 256)
wgrad_project("wgrad+project 12800x256 K=64 perm(512,25)", 64, 12800, 256, (512, 25))
wgrad_project("wgrad+project LIS 256x256 K=64", 64, 256, 256)
