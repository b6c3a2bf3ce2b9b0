"""Data-parallel host logic on CPU: bucket planning, the gloo world-size-2 gradient exchange and
the cross-rank agreement of the stochastic LIS depth."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def test_plan_buckets_cover_buffer_back_to_front():
    from glis_b200.dp import plan_buckets
    assert plan_buckets(0, 4) == []
    assert plan_buckets(10, 4) == [(6, 4), (2, 4), (0, 2)]
    assert plan_buckets(8, 100) == [(0, 8)]
    for numel, b in ((1, 1), (17, 5), (1000, 64), (9_000_001, 1 << 20)):
        plan = plan_buckets(numel, b)
        assert plan[0][0] + plan[0][1] == numel and plan[-1][0] == 0          # starts at the end
        assert sum(n for _, n in plan) == numel and all(0 < n <= b for _, n in plan)
        for (o1, n1), (o2, n2) in zip(plan, plan[1:]):
            assert o2 + n2 == o1                                                  # contiguous, no overlap


def test_param_buckets_hold_whole_parameters():
    """A bucket of the overlapped exchange may only be sent once EVERY parameter overlapping it has its
    gradient, so buckets are unions of whole parameters (a large weight is never cut)."""
    from glis_b200.dp import plan_param_buckets
    offsets, sizes = [0, 8, 12, 1036, 1040, 1104], [6, 4, 1024, 4, 64, 300]   # 4-aligned layout, total 1404
    plan = plan_param_buckets(offsets, sizes, 1404, 100)
    assert plan[0] == (1104, 300, [5])                       # last parameter first
    assert plan[1] == (12, 1092, [4, 3, 2])                  # the 1024-element weight stays whole
    assert plan[2] == (0, 12, [1, 0])
    assert sum(n for _, n, _ in plan) == 1404
    assert sorted(i for _, _, m in plan for i in m) == list(range(6))
    one = plan_param_buckets(offsets, sizes, 1404, 10 ** 9)
    assert one == [(0, 1404, [5, 4, 3, 2, 1, 0])]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from glis_b200 import dp
    from common.model import GeneratorLearnedInputSpace
    r, w, _ = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    data_seed = dp.seed_everything(1234, rank)
    # identical weights on every rank, distinct data seeds
    gen = GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 3, "fractional")
    wsum = float(sum(p.double().sum() for p in gen.parameters()))
    depths = [gen.lis_depth(None) for _ in range(64)]
    # gradient exchange: sum over ranks, scale 1/world, bucketed
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    sync = dp.GradSync(world, bucket_mb=0.001)  # 262 elements per bucket -> 4 buckets
    gscale = sync(flat, "gen")
    results[rank] = dict(data_seed=data_seed, wsum=wsum, depths=depths, flat=flat.clone(), gscale=gscale,
                         bytes=sync.bytes_reduced)
    dist.destroy_process_group()


def test_gloo_world2_gradient_exchange_and_depth_agreement():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    a, b = results[0], results[1]
    assert a["data_seed"] != b["data_seed"]
    assert a["wsum"] == b["wsum"]                       # same initial weights
    assert a["depths"] == b["depths"] and len(set(a["depths"])) > 1   # same stochastic schedule, not constant
    want = torch.arange(1000, dtype=torch.float32) * 3  # (1 + 2) * i
    assert torch.equal(a["flat"], want) and torch.equal(b["flat"], want)
    assert a["gscale"] == 0.5 and a["bytes"] == 4000


def test_overlapped_sync_counts_each_parameter_once(monkeypatch):
    """A parameter is announced twice per backward — by the operator that added its gradient in place and
    by autograd's post-accumulate hook, which torch also runs for parameter inputs a Function returned
    None for.  A bucket must be sent only when every one of its parameters has reported (regression:
    double counting sent buckets before their last gradients were written and the replicas diverged)."""
    from glis_b200 import dp

    class Flat(object):
        def __init__(self, sizes):
            self.params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
            self.offsets, off = [], 0
            for n in sizes:
                self.offsets.append(off)
                off += n
            self.numel = off
            self.g = torch.zeros(off)

    flat = Flat([4, 4, 8, 4, 4, 8])
    sync = dp.OverlappedGradSync(world=2, bucket_mb=12 * 4 / float(1 << 20))   # 12 elements per bucket
    sent = []
    monkeypatch.setattr(sync, "_send", lambda st, b: (sent.append(b), st["sent"].__setitem__(b, True)))
    sync.register("net", flat)
    st = sync.sets["net"]
    assert [sorted(m) for m in ([i for i, o in enumerate(st["owner"]) if o == b] for b in range(len(st["buckets"])))] \
        == [[4, 5], [2, 3], [0, 1]]
    assert st["unit_of"] == [[0], [1], [2], [3], [4], [5]]         # small parameters: one unit each
    sync.begin("net")
    hooks = [p._glis_grad_hooks[0] for p in flat.params]
    for idx in (5, 5, 5):                 # the same parameter announced three times: still one report
        hooks[idx](flat.params[idx])
    assert sent == []
    hooks[4](flat.params[4])
    assert sent == [0]
    hooks[4](flat.params[4])              # late duplicate of a sent bucket: ignored
    for idx in (3, 3, 2, 2):
        hooks[idx](flat.params[idx])
    assert sent == [0, 1]
    sync.begin("net")                     # next backward: counters re-armed
    sent.clear()
    for idx in (5, 4, 3, 2, 1, 0):
        hooks[idx](flat.params[idx])
        hooks[idx](flat.params[idx])
    assert sent == [0, 1, 2]


def test_large_linear_weight_is_exchanged_in_row_chunks(monkeypatch):
    """The generator's initial linear (12800 x 256, the LAST gradient backward produces) is cut into row chunks that
    are buckets of their own: each chunk leaves as soon as the operator announces it, the whole-parameter
    announcement (autograd's hook, or an operator that does not work in chunks) completes whatever is left."""
    from glis_b200 import dp

    assert dp.split_rows(12800, 256, 800_000) == [(0, 3200), (3200, 3200), (6400, 3200), (9600, 3200)]
    assert dp.split_rows(100, 10, 10 ** 9) == [(0, 100)]
    parts = dp.split_rows(1000, 256, 40_000)
    assert sum(n for _, n in parts) == 1000 and all(r % 32 == 0 for r, _ in parts) and len(parts) <= 8

    class Flat(object):
        def __init__(self):
            self.params = [torch.nn.Parameter(torch.zeros(16)), torch.nn.Parameter(torch.zeros(12800, 256)),
                           torch.nn.Parameter(torch.zeros(512))]
            self.offsets = [0, 16, 16 + 12800 * 256]
            self.numel = 16 + 12800 * 256 + 512
            self.g = torch.zeros(self.numel)

    flat = Flat()
    sync = dp.OverlappedGradSync(world=8, bucket_mb=2, split_mb=6)      # (off by default: GLIS_DP_SPLIT_MB)
    sent = []
    monkeypatch.setattr(sync, "_send", lambda st, b: (sent.append(st["buckets"][b]), st["sent"].__setitem__(b, True)))
    sync.register("gen", flat)
    st = sync.sets["gen"]
    big = flat.params[1]
    parts = big._glis_grad_parts
    assert parts == dp.split_rows(12800, 256, 2 * (1 << 20) // 4) and 4 <= len(parts) <= 8
    assert len(st["unit_of"][1]) == len(parts) and sum(n for _, n in st["buckets"]) == flat.numel
    sync.begin("gen")
    hook = big._glis_grad_hooks[0]
    flat.params[2]._glis_grad_hooks[0](flat.params[2])          # the small trailing parameter first
    last = len(parts) - 1
    hook(big, last)                                             # last chunk ready: its bucket (with the tail) leaves
    assert sent == [(16 + parts[last][0] * 256, parts[last][1] * 256 + 512)]
    hook(big, 1); hook(big, 1)
    assert sent[-1] == (16 + parts[1][0] * 256, parts[1][1] * 256) and len(sent) == 2
    hook(big)                                                   # whole-parameter announcement: every other chunk
    assert len(sent) == len(parts)
    flat.params[0]._glis_grad_hooks[0](flat.params[0])          # (the 16-element head shares chunk 0's bucket)
    assert len(set(sent)) == len(sent) and sum(n for _, n in sent) == flat.numel
