#!/usr/bin/env python
"""Probe (torchrun, N ranks): all-reduce of a SLICE of one flat symmetric-memory buffer through torch's
symmetric-memory kernels (multimem / two-shot / one-shot) against NCCL — agreement, CUDA-graph capture, time per call
at the sizes the gradient exchange uses (13.1 MB head bucket, 11 MB, 2 MB)."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    group = dist.group.WORLD
    gname = group.group_name
    total = 12 * (1 << 20)          # floats: 48 MB flat buffer
    flat = symm_mem.empty(total, dtype=torch.float32, device="cuda")
    hdl = symm_mem.rendezvous(flat, group)
    if rank == 0:
        print("symmetric memory: world %d, multicast %s, buffer %d MB" % (hdl.world_size, hdl.has_multicast_support, total * 4 >> 20), flush=True)
    ref = torch.empty(total, device="cuda")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-error-avoidance_b200"))
    import ctypes as C
    from glis_b200 import dp
    px = dp._PeerExchange(world, group)
    bufs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
    base = flat.data_ptr()

    def ours(t):
        px.all_reduce(bufs, (t.data_ptr() - base) // 4, t.numel())

    ops = {"nccl": lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM), "ours": ours,
           "two_shot": lambda t: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname)}
    for mb, off in ((13.1, 1024 * 32), (11.0, 4 * (1 << 20)), (2.0, 64)):
        n = int(mb * (1 << 20) / 4) // 64 * 64
        for name, fn in ops.items():
            torch.manual_seed(1234 + rank)
            flat.normal_()
            ref.copy_(flat)
            dist.all_reduce(ref[off:off + n], op=dist.ReduceOp.SUM)
            view = flat[off:off + n]
            try:
                fn(view)
                torch.cuda.synchronize()
            except Exception as e:       # noqa: BLE001
                if rank == 0:
                    print("%-9s %5.1f MB: FAILED eagerly: %s" % (name, mb, str(e)[:200]), flush=True)
                continue
            err = (view - ref[off:off + n]).abs().max().item() / ref[off:off + n].abs().max().item()
            untouched = torch.equal(flat[:off], ref[:off]) and torch.equal(flat[off + n:], ref[off + n:])
            # CUDA graph: 10 calls per replay
            try:
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for _ in range(3):
                        fn(view)
                    torch.cuda.synchronize()
                    dist.barrier()
                    with torch.cuda.graph(g, stream=s):
                        for _ in range(10):
                            fn(view)
                g.replay(); torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 100.0
                graph = "%.1f us per call in a CUDA graph" % us
            except Exception as e:       # noqa: BLE001
                graph = "graph capture FAILED: %s" % str(e)[:160]
            if rank == 0:
                print("%-9s %5.1f MB at offset %d: rel err vs NCCL %.1e, rest untouched %s, %s" % (name, mb, off, err, untouched, graph), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
