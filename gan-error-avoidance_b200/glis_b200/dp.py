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


def plan_param_buckets(offsets, sizes, total, bucket_elems):
    """Buckets of WHOLE parameters for the overlapped exchange: walk the parameters from the last
    one (whose gradient backward produces first) towards the first, closing a bucket once it holds
    at least ``bucket_elems`` elements.  Returns [(offset, length, [param indices])], first-ready
    first; the buckets tile [0, total) (alignment padding rides along with its parameter)."""
    buckets, members, end = [], [], total
    for i in range(len(offsets) - 1, -1, -1):
        members.append(i)
        if end - offsets[i] >= bucket_elems or i == 0:
            start = offsets[i] if i > 0 else 0
            buckets.append((start, end - start, members))
            members, end = [], start
    return buckets


def plan_unit_buckets(units, total, bucket_elems):
    """As plan_param_buckets over UNITS — (start, length) pieces of the flat buffer in ascending order, a whole
    parameter or one row chunk of a large one: walk them from the last (first ready) to the first, closing a bucket
    once it holds at least ``bucket_elems`` elements.  Returns [(offset, length, [unit indices])]; the buckets tile
    [0, total)."""
    buckets, members, end = [], [], total
    for u in range(len(units) - 1, -1, -1):
        members.append(u)
        if end - units[u][0] >= bucket_elems or u == 0:
            start = units[u][0] if u > 0 else 0
            buckets.append((start, end - start, members))
            members, end = [], start
    return buckets


def split_rows(rows, row_elems, chunk_elems, max_parts=8):
    """[(row_begin, row_count)] cutting a (rows x row_elems) matrix into at most ``max_parts`` chunks of about
    ``chunk_elems`` elements (whole rows, multiples of 32 rows: the fused linear weight-gradient kernel works in
    32-row blocks)."""
    parts = max(1, min(max_parts, (rows * row_elems) // max(1, chunk_elems)))
    per = -(-rows // parts)
    per = -(-per // 32) * 32
    out, r = [], 0
    while r < rows:
        n = min(per, rows - r)
        out.append((r, n))
        r += n
    return out


class _PeerExchange(object):
    """State of the peer-memory all-reduce (``glis_peer_allreduce``): a symmetric flag array, this rank's epoch
    counters, and per adopted flat buffer the ranks' base pointers.  PyTorch's symmetric memory does the allocation
    and the handle exchange; the kernel is ours."""

    BLOCKS = 64

    def __init__(self, world, group):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self.world, self.group = world, group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.flags = symm_mem.empty(128 * 8, dtype=torch.int32, device=dev)
        self.flags.zero_()
        handle = symm_mem.rendezvous(self.flags, self.group)
        self.flag_ptrs = (C.c_void_p * world)(*[int(p) for p in handle.buffer_ptrs])
        self.epochs = torch.zeros(128, dtype=torch.int32, device=dev)
        self._keep = [handle]
        self._L, self._C = L, C
        torch.cuda.synchronize()
        dist.barrier(self.group)          # nobody signals a flag array that its owner has not zeroed yet

    def adopt(self, flat):
        handle = flat.to_symmetric(self.group)
        self._keep.append(handle)
        torch.cuda.synchronize()
        dist.barrier(self.group)
        return (self._C.c_void_p * self.world)(*[int(p) for p in handle.buffer_ptrs])

    def all_reduce(self, bufs, off, n):
        """Sum floats [off, off + n) of the adopted buffer across ranks, on the current stream."""
        L, C = self._L, self._C
        work = n // (4 * self.world)                      # float4 per rank
        blocks = max(1, min(self.BLOCKS, (work + 2047) // 2048))
        L.call("glis_peer_allreduce", bufs, self.flag_ptrs, self.rank, self.world, int(off), int(n),
               C.c_void_p(self.epochs.data_ptr()), blocks, L.stream())


class OverlappedGradSync(object):
    """Bucketed gradient all-reduce launched from a side stream WHILE backward is still running.

    The all-reduce of a bucket is ``glis_peer_allreduce`` — our own kernel over NVLink peer memory, the gradient
    buffers living in symmetric memory (``_PeerExchange``) — on 2, 4 or 8 ranks of one node, ``ncclAllReduce`` otherwise.

    One instance serves several flat buffers (``register(tag, flat)``).  Every parameter gets a
    post-accumulate-grad hook; when the last parameter of a bucket has its gradient, the main
    stream records an event, the side stream waits for it and issues that bucket's all-reduce, so
    the exchange of the late layers overlaps the backward of the early ones.  A LARGE 2-D parameter
    (a linear weight of ``split_mb`` or more: the generator's initial linear, 13 MB at config 2, whose gradient
    is the LAST one backward produces) is cut into row chunks that are buckets of their own: the operator
    that makes the gradient produces it chunk by chunk (``_glis_grad_parts``, ops._layer_backward) and the
    all-reduce of chunk i runs under the kernel of chunk i+1.  ``finish(tag)``
    flushes the buckets that never completed (a LIS module skipped by the stochastic depth keeps
    its zero-filled gradient — identically on every rank, so the flush order is the same
    everywhere), makes the main stream wait for the side stream and returns 1/world for the
    fused RMSprop.  Everything is stream-ordered, so the whole thing can sit inside a CUDA graph.
    """

    def __init__(self, world, bucket_mb=None, group=None, split_mb=None):
        self.world, self.group = world, group
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("GLIS_DP_BUCKET_MB", "16"))
        if split_mb is None:
            # 0 = off (default).  Measured on 8 B200 at config 2 (tools/dp_bench.sh): cutting the 13 MB gradient of
            # G's initial linear into 4-6 chunks is SLOWER (2.19 vs 2.12 ms): its kernel takes ~10 us, the exchange
            # ~100 us, so there is nothing to pipeline with — what is exposed is the all-reduce of the LAST gradient
            # of backward itself, and the optimizer needs it whole.
            split_mb = float(os.environ.get("GLIS_DP_SPLIT_MB", "0"))
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.split_elems = int(split_mb * (1 << 20) / 4)
        self.side = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.sets = {}
        self.bytes_reduced = 0
        # The exchange itself: our own all-reduce over NVLink peer memory (csrc/peer_allreduce.cu) when the gradient
        # buffers can live in symmetric memory (2, 4 or 8 ranks on one node; GLIS_DP_PEER=0: ncclAllReduce).
        self.peer = None
        if (world in (2, 4, 8) and torch.cuda.is_available() and os.environ.get("GLIS_DP_PEER", "1") != "0"
                and dist.is_initialized() and dist.get_backend(group) == "nccl"):
            try:
                self.peer = _PeerExchange(world, group)
            except Exception as e:       # noqa: BLE001 — no symmetric memory on this system: NCCL carries the exchange
                import warnings
                warnings.warn("glis_b200.dp: peer-memory gradient exchange unavailable (%s); using ncclAllReduce" % e)
                self.peer = None

    def register(self, tag, flat):
        peer_bufs = None
        if self.peer is not None:
            try:
                peer_bufs = self.peer.adopt(flat)
            except Exception as e:       # noqa: BLE001 — this buffer stays where it is: ncclAllReduce carries its exchange
                import warnings
                warnings.warn("glis_b200.dp: %s gradients not in symmetric memory (%s); using ncclAllReduce" % (tag, e))
        # units: whole parameters, or the row chunks of a large linear weight
        units, unit_of = [], []          # (start, length); per parameter: its unit indices, in part order
        for idx, (p, off) in enumerate(zip(flat.params, flat.offsets)):
            parts = None
            if self.split_elems > 0 and p.dim() == 2 and p.numel() >= self.split_elems and self.world > 1:
                parts = split_rows(p.shape[0], p.shape[1], max(self.split_elems // 4, self.bucket_elems))
                if len(parts) < 2:
                    parts = None
            if parts is None:
                nxt = flat.offsets[idx + 1] if idx + 1 < len(flat.offsets) else flat.numel
                unit_of.append([len(units)])
                units.append((off, nxt - off))
            else:
                p._glis_grad_parts = parts
                mine = []
                for k, (r0, rc) in enumerate(parts):
                    start = off + r0 * p.shape[1]
                    length = rc * p.shape[1]
                    if k == len(parts) - 1:      # alignment padding rides along with the last chunk
                        nxt = flat.offsets[idx + 1] if idx + 1 < len(flat.offsets) else flat.numel
                        length = nxt - start
                    mine.append(len(units))
                    units.append((start, length))
                unit_of.append(mine)
        plan = plan_unit_buckets(units, flat.numel, self.bucket_elems)
        buckets = [(o, n) for o, n, _ in plan]
        owner = [None] * len(units)
        for b, (_, _, members) in enumerate(plan):
            for u in members:
                owner[u] = b
        st = {"flat": flat, "buckets": buckets, "owner": owner, "unit_of": unit_of,
              "need": [len(m) for _, _, m in plan], "left": [0] * len(buckets), "sent": [False] * len(buckets),
              "armed": False, "seen": set(), "peer_bufs": peer_bufs}
        self.sets[tag] = st
        for idx, p in enumerate(flat.params):
            hook = self._make_hook(st, idx)
            p.register_post_accumulate_grad_hook(hook)          # gradients accumulated by autograd
            hooks = getattr(p, "_glis_grad_hooks", None)        # gradients the kernels add in place (ops._touch_hooks)
            if hooks is None:
                hooks = p._glis_grad_hooks = []
            hooks.append(hook)

    def _make_hook(self, st, idx):
        def hook(_param, part=None):
            if not st["armed"] or self.world <= 1:
                return
            # A unit reports ONCE per backward pass.  A parameter may be announced twice — by the operator
            # that added its gradient in place (ops._touch_hooks) and again by autograd, which runs the
            # post-accumulate hooks of every parameter input of a Function even when the Function
            # returned None for it — and either call comes after the gradient work was enqueued.
            # ``part``: one row chunk of a split parameter; None: the whole parameter (every chunk not yet seen).
            mine = st["unit_of"][idx]
            for u in (mine if part is None else [mine[part]]):
                if u in st["seen"]:
                    continue
                st["seen"].add(u)
                b = st["owner"][u]
                st["left"][b] -= 1
                if st["left"][b] == 0 and not st["sent"][b]:
                    self._send(st, b)
        return hook

    def _send(self, st, b):
        off, n = st["buckets"][b]
        ready = torch.cuda.Event()
        ready.record()                      # gradients of this bucket are complete on the current stream ...
        self.side.wait_event(ready)
        from . import ops                   # ... and on the other one of (main, weight-gradient side stream)
        for other in [ops.Overlap.main_stream()] + ops.Overlap.side_streams():
            if other is not None and other != torch.cuda.current_stream():
                ev = torch.cuda.Event()
                ev.record(other)
                self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            if st["peer_bufs"] is not None:
                self.peer.all_reduce(st["peer_bufs"], off, n)
            else:
                dist.all_reduce(st["flat"].g[off:off + n], op=dist.ReduceOp.SUM, group=self.group)
        st["sent"][b] = True
        self.bytes_reduced += 4 * n

    def begin(self, tag):
        """Call after zero_grad and before the backward pass that fills ``tag``'s gradients."""
        st = self.sets[tag]
        st["left"] = list(st["need"])
        st["sent"] = [False] * len(st["buckets"])
        st["seen"] = set()
        st["armed"] = True

    def finish(self, tag):
        st = self.sets[tag]
        st["armed"] = False
        if self.world <= 1:
            return 1.0
        for b in range(len(st["buckets"])):
            if not st["sent"][b]:
                self._send(st, b)
        torch.cuda.current_stream().wait_stream(self.side)
        return 1.0 / self.world


def seed_everything(seed, rank):
    """Weights and LIS-depth draws identical on every rank; data streams distinct."""
    random.seed(seed)
    torch.manual_seed(seed)
    return seed + 1 + rank  # the per-rank data seed
