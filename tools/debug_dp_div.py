#!/usr/bin/env python
"""Which parameters diverge between data-parallel replicas, and after which iteration (torchrun, 2 ranks)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gan-error-avoidance_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import common.model as pm
from glis_b200 import dp
from glis_b200.trainer import GLISTrainer, GraphedStep


def build(dev, seed=77):
    torch.manual_seed(seed)
    g = pm.GeneratorLearnedInputSpace(32, 32, 16, 3, 32, "weight", 2, "fractional").to(dev)
    d = pm.build_discriminator(32, 32, 16, 3, "weight", 0).to(dev)
    return g, d


def main():
    rank, world, local = dp.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B, code, lr = 4, 32, 2e-5
    gen = torch.Generator().manual_seed(5)
    batches = [(torch.rand(B * world, 3, 32, 32, generator=gen), torch.randn(B * world, code, generator=gen),
                torch.randn(B * world, code, generator=gen)) for _ in range(3)]
    depths = [(2, 1), (0, 2), (2, 1)]
    modes = [m == "1" for m in (sys.argv[1] if len(sys.argv) > 1 else "01")]
    for use_graph in modes:
        g, d = build(dev)
        tr = GLISTrainer(g, d, lr=lr, grad_sync=dp.OverlappedGradSync(world, bucket_mb=0.05))
        stepper = GraphedStep(tr, B, 32, 32, code, dev, warmup=1) if use_graph else tr
        sl = slice(rank * B, (rank + 1) * B)
        for it, ((real, zd, zg), dep) in enumerate(zip(batches, depths)):
            stepper.step(real[sl].to(dev), zd[sl].to(dev), zg[sl].to(dev), *dep)
            torch.cuda.synchronize()
            for tag, net, flat in (("gen", g, tr.gen_flat), ("dis", d, tr.dis_flat)):
                for (name, p), o in zip(net.named_parameters(), flat.offsets):
                    for what, buf in (("p", flat.p), ("g", flat.g), ("v", flat.v)):
                        mine = buf[o:o + p.numel()].clone()
                        ref = mine.clone()
                        dist.broadcast(ref, 0)
                        diff = (mine - ref).abs().max().item()
                        if rank == 1 and diff > 0:
                            print("graph=%s it%d %s %-36s %s differs by %.3e (max |ref| %.3e)"
                                  % (use_graph, it, tag, name, what, diff, ref.abs().max().item()), flush=True)
        dist.barrier()
        if rank == 1:
            print("graph=%s done" % use_graph, flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
