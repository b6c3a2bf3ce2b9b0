#!/usr/bin/env python
"""N-rank data-parallel step == 1-rank step at the global batch (SURVEY.md §4 "DP tests").

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_equivalence_gpu.py

Each rank trains its shard of the same global batch for three iterations through the
side-stream overlapped gradient exchange (also replayed as a CUDA graph); rank 0 runs the
whole batch on one GPU and compares losses and parameters.
"""
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
    for use_graph in (False, True):
        g, d = build(dev)
        tr = GLISTrainer(g, d, lr=lr, grad_sync=dp.OverlappedGradSync(world, bucket_mb=0.05, split_mb=0.05))
        stepper = GraphedStep(tr, B, 32, 32, code, dev, warmup=1) if use_graph else tr
        sl = slice(rank * B, (rank + 1) * B)
        losses = []
        for (real, zd, zg), dep in zip(batches, depths):
            out = stepper.step(real[sl].to(dev), zd[sl].to(dev), zg[sl].to(dev), *dep)
            t = torch.stack([out["d_real"], out["d_fake"], out["g"]]).clone()
            dist.all_reduce(t)
            losses.append((t / world).tolist())
        # replicas must stay bit-identical
        flat = torch.cat([tr.gen_flat.p, tr.dis_flat.p])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "replicas diverged"
        if rank == 0:
            g1, d1 = build(dev)
            t1 = GLISTrainer(g1, d1, lr=lr)
            for i, ((real, zd, zg), dep) in enumerate(zip(batches, depths)):
                o = t1.step(real.to(dev), zd.to(dev), zg.to(dev), *dep)
                for j, k in enumerate(("d_real", "d_fake", "g")):
                    a, b = losses[i][j], o[k].item()
                    assert abs(a - b) <= 2e-3 * abs(b), (use_graph, i, k, a, b)
            worst = 0.0
            for f_dp, f_1 in ((tr.gen_flat, t1.gen_flat), (tr.dis_flat, t1.dis_flat)):
                worst = max(worst, (f_dp.p - f_1.p).abs().max().item())
            assert worst <= 3 * 6.4 * lr + 1e-7, worst   # one sign-like RMSprop step per iteration at most
            print("dp equivalence ok (graph=%s): losses %s, worst |dp - single| = %.2e, %d bytes all-reduced"
                  % (use_graph, [round(v, 5) for v in losses[-1]], worst, tr.grad_sync.bytes_reduced))
        dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)   # captured NCCL work keeps the communicator busy at teardown; nothing left to clean up


if __name__ == "__main__":
    main()
