#!/usr/bin/env python
"""Time the tcgen05 conv kernel per layer shape with CUDA events (GLIS_TC_DEBUG variants)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gan-error-avoidance_b200"))
import torch
from glis_b200 import _lib as L, ops

SHAPES = [  # name, relation, N, Hi, Wi, Ci, Ho, Wo, Co
    ("D1 conv 2B", L.CONV, 128, 40, 40, 64, 20, 20, 128),
    ("D2 conv 2B", L.CONV, 128, 20, 20, 128, 10, 10, 256),
    ("D3 conv 2B", L.CONV, 128, 10, 10, 256, 5, 5, 512),
    ("dD3 tconv 2B", L.TCONV, 128, 5, 5, 512, 10, 10, 256),
    ("dD2 tconv 2B", L.TCONV, 128, 10, 10, 256, 20, 20, 128),
    ("dD1 tconv 2B", L.TCONV, 128, 20, 20, 128, 40, 40, 64),
    ("D1 conv 64->128 40->20", L.CONV, 64, 40, 40, 64, 20, 20, 128),
    ("D2 conv 128->256 20->10", L.CONV, 64, 20, 20, 128, 10, 10, 256),
    ("D3 conv 256->512 10->5", L.CONV, 64, 10, 10, 256, 5, 5, 512),
    ("G3 tconv 512->256 5->10", L.TCONV, 64, 5, 5, 512, 10, 10, 256),
    ("G2 tconv 256->128 10->20", L.TCONV, 64, 10, 10, 256, 20, 20, 128),
    ("G1 tconv 128->64 20->40", L.TCONV, 64, 20, 20, 128, 40, 40, 64),
    ("dG1 conv 64->128 40->20", L.CONV, 64, 40, 40, 64, 20, 20, 128),
    ("dD1 tconv 128->64 20->40", L.TCONV, 64, 20, 20, 128, 40, 40, 64),
]


def main():
    dev = "cuda"
    spec = ops.ContractionSpec(False, (4, 4), (2, 2), (1, 1), (1, 1))
    for name, rel, n, hi, wi, ci, ho, wo, co in SHAPES:
        g = spec.geom(rel, n, hi, wi, ci, ho, wo, co)
        x = torch.randn(n, hi, wi, ci, device=dev)
        xh, xl = ops.split_bf16(x)
        w = torch.randn(16, co, ci, device=dev) * 0.05
        wh, wl = ops.split_bf16(w)
        out = torch.empty(n, ho, wo, co, device=dev)
        ep = L.Epilogue(None, 0, None, None, None)
        out_f32, out_hi, out_lo = out, None, None
        if os.environ.get("TCMB_EPILOGUE") == "1":    # the fused-chain forward epilogue: bias + TPReLU, pre-activation + bf16 planes
            bias = torch.randn(co, device=dev); ta = torch.full((co,), 0.25, device=dev); tb = torch.zeros(co, device=dev)
            pre = torch.empty_like(out)
            out_hi = torch.empty(n, ho, wo, co, device=dev, dtype=torch.bfloat16); out_lo = torch.empty_like(out_hi)
            ep = L.Epilogue(L.ptr(bias), L.ACT_TPRELU, L.ptr(ta), L.ptr(tb), L.ptr(pre))
            out_f32 = None
        res = []
        ref = None
        for cl in (sys.argv[1:] or ["1", "2", "4"]):
            dbg = "0"
            os.environ["GLIS_TC_DEBUG"] = dbg
            os.environ.pop("GLIS_TC_HALO", None)
            os.environ.pop("GLIS_T2_DEBUG", None)
            os.environ.pop("GLIS_TC_HALO_BK", None)
            os.environ.pop("GLIS_TC_HALO_MINW", None)
            os.environ.pop("GLIS_TC_PAIR", None)
            if cl.startswith("p"):      # "p0" / "p1": one-CTA kernels / cta_group::2 pairs where they apply
                os.environ["GLIS_TC_PAIR"] = cl[1:]
                os.environ["GLIS_TC_CLUSTER"] = "1"
                os.environ.pop("GLIS_TC_KSPLIT", None)
                cl = "P" + cl[1:]
            if cl.startswith("b"):      # "b32" / "b64": halo kernel on every map width with that many channels per stage
                os.environ["GLIS_TC_HALO_BK"] = cl[1:]
                os.environ["GLIS_TC_HALO_MINW"] = "1"
                os.environ["GLIS_TC_CLUSTER"] = "1"
                os.environ.pop("GLIS_TC_KSPLIT", None)
                cl = "x"
            if cl.startswith("c"):      # "c2d6": tc_conv.cu, cluster 2 (weight multicast), debug bits 6 (weight loads only)
                os.environ["GLIS_TC_HALO"] = "0"
                os.environ["GLIS_TC_CLUSTER"], os.environ["GLIS_TC_DEBUG"] = cl[1:].split("d")
                os.environ["GLIS_TC_KSPLIT"] = "1"
            elif cl.startswith("t"):      # "t1" ...: halo kernel with GLIS_T2_DEBUG bits
                os.environ["GLIS_T2_DEBUG"] = cl[1:]
                os.environ["GLIS_TC_CLUSTER"] = "1"
                os.environ.pop("GLIS_TC_KSPLIT", None)
            elif cl.startswith("d"):      # "d1" / "d2" / "d4" / "d6": tc_conv.cu with debug bits (1 no stores, 2 no MMA, 4 no pixel loads)
                os.environ["GLIS_TC_HALO"] = "0"
                os.environ["GLIS_TC_DEBUG"] = cl[1:]
                os.environ["GLIS_TC_CLUSTER"] = "1"
                os.environ.pop("GLIS_TC_KSPLIT", None)
            elif cl.startswith("h"):      # "h0" / "h1": one box per tap (tc_conv.cu) / halo kernel (tc_conv2.cu)
                os.environ["GLIS_TC_HALO"] = cl[1:]
                os.environ["GLIS_TC_CLUSTER"] = "1"
                os.environ.pop("GLIS_TC_KSPLIT", None)
            elif cl.startswith("P"):
                pass
            elif cl.startswith("k"):      # "k1", "k8": upper bound on the K split, no cluster
                os.environ["GLIS_TC_CLUSTER"] = "1"
                os.environ["GLIS_TC_KSPLIT"] = cl[1:]
            else:
                os.environ["GLIS_TC_CLUSTER"] = cl
                os.environ["GLIS_TC_KSPLIT"] = "1"
            for prec in (L.PREC_BF16X3,):
                def run():
                    L.call("glis_conv_forward_bf16", C.byref(g), L.ptr16(xh), L.ptr16(xl), L.ptr16(wh), L.ptr16(wl),
                           C.byref(ep), L.ptr(out_f32), L.ptr16(out_hi), L.ptr16(out_lo), prec, L.stream())
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()       # replay 20 launches: GPU time, not host launch time
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
                chk = out if out_f32 is not None else pre
                chk.zero_()
                run()
                torch.cuda.synchronize()
                if ref is None:
                    ref = chk.clone()
                err = ((chk - ref).abs().max() / ref.abs().max()).item()
                res.append("cl%s %6.1fus (dev %.1e)" % (cl, e0.elapsed_time(e1) * 50, err))
        flop = 2.0 * n * (ho * wo if rel == L.CONV else hi * wi) * co * ci * (16 if rel == L.CONV else 16)
        print("%-28s %5.2f GFLOP | %s" % (name, flop / 1e9, "  ".join(res)))
    os.environ["GLIS_TC_DEBUG"] = "0"
    os.environ.pop("GLIS_TC_CLUSTER", None)
    os.environ.pop("GLIS_TC_KSPLIT", None)


if __name__ == "__main__":
    main()
