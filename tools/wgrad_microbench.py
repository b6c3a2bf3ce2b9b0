#!/usr/bin/env python
"""Time the tcgen05 weight-gradient kernel per layer shape (GLIS_WG_DEBUG / GLIS_WG_COLS variants)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

SHAPES = [  # name, N, Hi, Wi, Cb (fine), Ho, Wo, Ca (coarse)
    ("D1 wgrad 2B 64->128 40->20", 128, 40, 40, 64, 20, 20, 128),
    ("D2 wgrad 2B 128->256 20->10", 128, 20, 20, 128, 10, 10, 256),
    ("D3 wgrad 2B 256->512 10->5", 128, 10, 10, 256, 5, 5, 512),
    ("G3 wgrad 512->256 5->10", 64, 10, 10, 256, 5, 5, 512),
    ("G2 wgrad 256->128 10->20", 64, 20, 20, 128, 10, 10, 256),
    ("G1 wgrad 128->64 20->40", 64, 40, 40, 64, 20, 20, 128),
]
SHAPES_1X1 = [  # image-side weight gradients as 1x1 products: name, N, H, W, Cb, Ca
    ("D0 wgrad 2B (unfolded)", 128, 40, 40, 48, 64),
    ("G0 wgrad (unfolded)", 64, 40, 40, 48, 64),
]


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


def main():
    dev = "cuda"
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    variants = sys.argv[1:] or ["0:128", "1:128", "2:128", "4:128", "8:128", "0:256", "1:256", "0:64"]
    for name, n, hi, wi, cb, ho, wo, ca in SHAPES:
        g = spec.geom(L.CONV, n, hi, wi, cb, ho, wo, ca)
        small = torch.randn(n, ho, wo, ca, device=dev)
        big = torch.randn(n, hi, wi, cb, device=dev)
        sp, bp = ops.split_bf16(small), ops.split_bf16(big)
        G = torch.zeros(ca, cb, 4, 4, device=dev)
        res = []
        for v in variants:
            parts = v.split(":")
            dbg, cols = parts[0], parts[1]
            os.environ["GLIS_WG_DEBUG"] = dbg
            os.environ["GLIS_WG_COLS"] = cols
            if len(parts) > 2:
                os.environ["GLIS_WG_EPI"] = parts[2]
            else:
                os.environ.pop("GLIS_WG_EPI", None)
            t = timeit(lambda: L.call("glis_conv_wgrad_bf16", C.byref(g), L.ptr16(sp[0]), L.ptr16(sp[1]),
                                      L.ptr16(bp[0]), L.ptr16(bp[1]), L.ptr(G), L.PREC_BF16X3, L.stream()))
            res.append("%s %6.1f" % (v, t))
        flop = 2.0 * n * ho * wo * ca * cb * 16
        print("%-30s %5.2f GF | %s" % (name, flop / 1e9, "  ".join(res)), flush=True)
    spec1 = ops.ContractionSpec(False, (1, 1), (1, 1), (0, 0), (1, 1))
    for name, n, h, w, cb, ca in SHAPES_1X1:
        g = spec1.geom(L.CONV, n, h, w, cb, h, w, ca)
        sp = ops.split_bf16(torch.randn(n, h, w, ca, device=dev))
        bp = ops.split_bf16(torch.randn(n, h, w, cb, device=dev))
        G = torch.zeros(ca, cb, device=dev)
        res = []
        for v in variants:
            parts = v.split(":")
            os.environ["GLIS_WG_DEBUG"] = parts[0]
            os.environ["GLIS_WG_COLS"] = parts[1]
            if len(parts) > 2:
                os.environ["GLIS_WG_EPI"] = parts[2]
            else:
                os.environ.pop("GLIS_WG_EPI", None)
            t = timeit(lambda: L.call("glis_conv_wgrad_bf16", C.byref(g), L.ptr16(sp[0]), L.ptr16(sp[1]),
                                      L.ptr16(bp[0]), L.ptr16(bp[1]), L.ptr(G), L.PREC_BF16X3, L.stream()))
            res.append("%s %6.1f" % (v, t))
        print("%-30s          | %s" % (name, "  ".join(res)), flush=True)
    os.environ["GLIS_WG_DEBUG"] = "0"
    os.environ.pop("GLIS_WG_COLS", None)
    os.environ.pop("GLIS_WG_EPI", None)


if __name__ == "__main__":
    main()
