"""Data parallelism for the G-LIS step: one process per GPU, NCCL over NVLink.

The step has no batch statistics (norm='weight') and every loss is a batch mean, so N
ranks at B images each equal one rank at N*B if the gradients are averaged
(SURVEY.md §8e).  Each network's gradients live in ONE flat fp32 buffer
(``trainer.FlatParams``); the exchange is a sum all-reduce of contiguous bucket views of
that buffer and the 1/N goes into the fused RMSprop kernel (``gscale``).  The stochastic
LIS depth must agree across ranks: every rank seeds ``gen.rng`` identically.
"""
import os
import random

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Join the torchrun world (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def plan_buckets(numel, bucket_elems):
    """Split [0, numel) into contiguous (offset, length) buckets of at most ``bucket_elems``
    elements, walking from the END of the buffer (the last layers' gradients are produced
    first by backward) — so bucket 0 is the first one ready."""
    if numel <= 0:
        return []
    bucket_elems = max(1, int(bucket_elems))
    out, end = [], numel
    while end > 0:
        start = max(0, end - bucket_elems)
        out.append((start, end - start))
        end = start
    return out


class GradSync(object):
    """Callable ``(flat_grad, tag) -> gscale`` installed into ``GLISTrainer``."""

    def __init__(self, world, bucket_mb=32.0, group=None):
        self.world = world
        self.group = group
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.bytes_reduced = 0

    def __call__(self, flat_grad, tag):
        if self.world <= 1:
            return 1.0
        handles = []
        for off, n in plan_buckets(flat_grad.numel(), self.bucket_elems):
            handles.append(dist.all_reduce(flat_grad[off:off + n], op=dist.ReduceOp.SUM, group=self.group,
                                           async_op=True))
            self.bytes_reduced += 4 * n
        for h in handles:
            h.wait()
        return 1.0 / self.world


def seed_everything(seed, rank):
    """Weights and LIS-depth draws identical on every rank; data streams distinct."""
    random.seed(seed)
    torch.manual_seed(seed)
    return seed + 1 + rank  # the per-rank data seed
