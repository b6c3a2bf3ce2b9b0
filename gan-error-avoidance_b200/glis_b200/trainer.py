"""The G-LIS adversarial training iteration on the sm_100a kernels.

``GLISTrainer.step`` performs exactly one iteration of the reference's loop
(g_lis/main.py:526-589): D on a real batch, D on a generated batch (G under no-grad),
RMSprop on D, then G + LIS against D with the latent-reconstruction losses, RMSprop on
G.  Parameters, gradients and RMSprop state of each network live in three flat fp32
buffers so that ``zero_grad`` is one memset, the optimizer one kernel and — under data
parallelism — the gradient exchange a handful of large all-reduces.

Legacy semantics kept: gradients are zero-filled, never dropped, so a
LIS module skipped by the stochastic depth still decays its ``square_avg``; a parameter
that never received a gradient has g = v = 0 and the update leaves it untouched.
"""
import torch
import torch.nn.functional as F

from . import ops


class FlatParams(object):
    """Re-homes a module's parameters (and their ``.grad``) into contiguous flat buffers."""

    ALIGN = 32  # floats (128 bytes): 16-byte kernels, and any run of whole parameters splits into 8 ranks x float4

    def __init__(self, module):
        self.params = [p for p in module.parameters()]
        if not self.params:
            raise ValueError("module has no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("glis_b200: move the networks to a CUDA device first (no CPU path)")
        self.offsets = []
        total = 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = total
        self.p = torch.zeros(total, device=dev, dtype=torch.float32)
        # gradients and the raw-filter-gradient scratch share one allocation: one memset clears both
        self.gs = torch.zeros(2 * total, device=dev, dtype=torch.float32)
        self.g, self.scratch = self.gs[:total], self.gs[total:]
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.p[o:o + n].view(p.shape)
            p.grad = self.g[o:o + n].view(p.shape)
            p._glis_direct_grad = True                      # kernels may add into .grad in place
            p._glis_scratch = self.scratch[o:o + n].view(p.shape) if p.dim() >= 2 else None

    def to_symmetric(self, group):
        """Re-home the gradient buffer (+ scratch) in symmetric memory — every rank's copy mapped into every rank's
        address space — so that the gradient exchange can run over NVLink peer loads / stores
        (``glis_peer_allreduce``).  Collective: every rank of ``group`` calls it, in the same order.  Returns the
        symmetric-memory handle (``buffer_ptrs``: the ranks' copies)."""
        import torch.distributed._symmetric_memory as symm_mem
        new = symm_mem.empty(2 * self.numel, dtype=torch.float32, device=self.p.device)
        new.copy_(self.gs)
        handle = symm_mem.rendezvous(new, group)
        self.gs = new
        self.g, self.scratch = new[:self.numel], new[self.numel:]
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            p.grad = self.g[o:o + n].view(p.shape)
            p._glis_direct_grad = True
            p._glis_scratch = self.scratch[o:o + n].view(p.shape) if p.dim() >= 2 else None
        return handle

    def _scratch_clean(self):
        for p in self.params:
            p._glis_scratch_dirty = False

    def zero_grad(self):
        self.gs.zero_()
        self._scratch_clean()

    def zero_grad_async(self):
        """Clear the gradients on the side stream (they are not touched before the next backward); returns the
        event the stream that runs that backward has to wait for, or None when done in place."""
        self._scratch_clean()
        if not (ops.Overlap.enabled and torch.cuda.is_available()):
            self.gs.zero_()
            return None
        side, main = ops.Overlap.stream(), torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)                 # everything that read the old gradients has been enqueued
        side.wait_event(ready)
        with torch.cuda.stream(side):
            self.gs.zero_()
            done = torch.cuda.Event()
            done.record(side)
        return done

    def rebind_grads(self):
        """Autograd may have replaced ``.grad`` (it does not for in-place accumulation, but a
        user could); make sure every parameter still accumulates into the flat buffer."""
        for p, o in zip(self.params, self.offsets):
            want = self.g[o:o + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != want.data_ptr():
                if p.grad is not None:
                    want.add_(p.grad)
                p.grad = want
                p._glis_direct_grad = True

    def optimizer_state_dict(self, lr, alpha=0.9, eps=1e-6):
        """RMSprop state in ``torch.optim.RMSprop.state_dict()`` form — what the reference saves as
        ``*_opt.pt`` (g_lis/main.py:345-349): loads into a stock ``torch.optim.RMSprop`` over the same
        parameters; ``load_optimizer_state_dict`` reads it back, and also the 2017 format (below)."""
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            state[i] = {"step": torch.tensor(0.0),
                        "square_avg": self.v[o:o + p.numel()].view(p.shape).detach().cpu().clone()}
        group = {"lr": lr, "momentum": 0, "alpha": alpha, "eps": eps, "centered": False, "weight_decay": 0,
                 "capturable": False, "foreach": None, "maximize": False, "differentiable": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        """Accepts both ``Optimizer.state_dict()`` formats: today's (state keyed by the parameter's index in
        ``param_groups[*]['params']``) and the reference era's (mid-2017: ``'params'`` lists ``id(p)`` of the
        saving process and ``state`` is keyed by those ids).  Either way the i-th entry of the concatenated
        ``params`` lists is the key of the i-th parameter."""
        keys = [k for group in sd["param_groups"] for k in group["params"]]
        if len(keys) != len(self.params):
            raise ValueError("optimizer state has %d parameters, the network has %d" % (len(keys), len(self.params)))
        matched = 0
        for key, p, o in zip(keys, self.params, self.offsets):
            entry = sd["state"].get(key)
            seg = self.v[o:o + p.numel()]
            if entry is None:          # never stepped in the saved run: legacy "skip" == zero state
                seg.zero_()
                continue
            sq = entry["square_avg"]
            if sq.numel() != p.numel():
                raise ValueError("optimizer state entry %r has %d elements, its parameter %d"
                                 % (key, sq.numel(), p.numel()))
            seg.copy_(sq.reshape(-1).to(seg.device, dtype=torch.float32))
            matched += 1
        if sd["state"] and not matched:
            raise ValueError("optimizer state is not empty but none of its entries matched a parameter")

    def rmsprop_step(self, lr, alpha=0.9, eps=1e-6, gscale=1.0):
        ops.rmsprop_(self.p, self.g, self.v, lr, alpha, eps, gscale, params=self.params)


def _split_head(dis):
    """(layers up to the logits, True) when ``dis`` ends in Sigmoid + View(1) as build_discriminator
    makes it (common/model.py:56-62), so that the loss can run on logits in one fused kernel."""
    mods = list(dis.children()) if isinstance(dis, torch.nn.Sequential) else []
    if len(mods) >= 3 and isinstance(mods[-2], torch.nn.Sigmoid) and type(mods[-1]).__name__ == "View":
        return mods[:-2]
    return None


def dis_bce(dis, x, targets, ls=False):
    """[mean loss(dis(x)[chunk_i], targets[i])] for equal batch chunks of ``x``, through autograd — the path for a
    discriminator WITHOUT the standard Sigmoid + View(1) head (``bce_on_logits`` serves the standard one)."""
    n = x.shape[0] // len(targets)
    p = dis(x)
    lossfunc = F.mse_loss if ls else F.binary_cross_entropy
    return [lossfunc(p[i * n:(i + 1) * n], torch.full_like(p[i * n:(i + 1) * n], t)) for i, t in enumerate(targets)]


def dis_logits(dis, x):
    """Logits (N,) of a standard discriminator (common/model.py:56-62) — everything up to the Sigmoid —
    or None when ``dis`` does not end in Sigmoid + View(1)."""
    from common.model import run_layers
    body = _split_head(dis)
    if body is None:
        return None
    return run_layers(body, x).reshape(-1)


def bce_on_logits(logits, targets, ls=False, gscale=1.0):
    """([mean loss of chunk i against targets[i]], gscale * d(sum of those means)/d(logits)): one fused kernel
    per chunk — sigmoid + BCE (g_lis/main.py:311,555,564,578) or, with ``ls``, sigmoid + squared error (--ls,
    :308-311) — that also leaves the gradient; nothing for autograd to trace: the caller starts backward from
    the logits with the returned gradient."""
    lg = logits.detach()
    n = lg.numel() // len(targets)
    dl = torch.empty_like(lg)
    losses = torch.empty(len(targets), device=lg.device, dtype=torch.float32)
    entry = "glis_lsq_logits" if ls else "glis_bce_logits"
    for i, t in enumerate(targets):
        ops.L.call(entry, ops.L.ptr(lg[i * n:(i + 1) * n]), float(t), n, float(gscale), ops.L.ptr(losses[i:i + 1]),
                   ops.L.ptr(dl[i * n:(i + 1) * n]), None, ops.L.stream())
    return [losses[i] for i in range(len(targets))], dl


def _has_dropout(*nets):
    return any(isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d)) and m.p > 0 for net in nets for m in net.modules())


class PairedBatch(object):
    """The 2B-image batch of a D update, real images in the first half and generated ones in the second, as ONE
    persistent NHWC buffer: the generator's last kernel writes its images straight into the second half and a
    caller that fills ``real_view`` in place (GraphedStep's static input is that view) costs no copy at all."""

    def __init__(self):
        self.buf = None

    def views(self, real):
        B = real.shape[0]
        shape = (2 * B,) + tuple(real.shape[1:])
        if self.buf is None or tuple(self.buf.shape) != shape or self.buf.device != real.device:
            self.buf = torch.empty(shape, device=real.device, dtype=torch.float32).contiguous(
                memory_format=torch.channels_last)
        return self.buf, self.buf[:B], self.buf[B:]


class GLISTrainer(object):
    """One object per (G-LIS, D) pair; ``step`` = one reference training iteration."""

    def __init__(self, gen, dis, lr, lambda_r=0.9, alpha=0.9, eps=1e-6, grad_sync=None, ls=False):
        self.gen, self.dis = gen, dis
        self.lr, self.lambda_r, self.alpha, self.eps = lr, lambda_r, alpha, eps
        self.ls = bool(ls)                      # --ls: squared error on D's sigmoid output instead of BCE
        self.paired = PairedBatch()
        self.dropout = _has_dropout(gen, dis)
        self.gen_flat = FlatParams(gen)
        self.dis_flat = FlatParams(dis)
        # data parallelism: either a callable (flat_grad_tensor, tag) -> gscale run after backward
        # (dp.GradSync) or an object with register/begin/finish that overlaps the exchange with
        # backward from a side stream (dp.OverlappedGradSync)
        self.grad_sync = grad_sync
        self._overlapped = hasattr(grad_sync, "register")
        if self._overlapped:
            grad_sync.register("gen", self.gen_flat)
            grad_sync.register("dis", self.dis_flat)
        self.launches = 0

    def _sync_begin(self, tag):
        if self._overlapped:
            self.grad_sync.begin(tag)

    def _sync_done(self, tag, flat):
        if self.grad_sync is None:
            return 1.0
        if self._overlapped:
            return self.grad_sync.finish(tag)
        return self.grad_sync(flat.g, tag)

    def _set_dis_requires_grad(self, flag):
        for p in self.dis_flat.params:
            p.requires_grad_(flag)

    def step(self, real, z_d, z_g, depth_d=None, depth_g=None):
        gen, dis = self.gen, self.dis
        B = real.shape[0]
        if self.dropout:
            ops.DropoutClock.tick(real.device)     # fresh dropout masks for this iteration (also under graph replay)

        # G's data-gradient packs are not needed before the G update's backward: rebuilt on the side stream
        ops.refresh_packs(self.gen_flat, part="backward", side=True)

        # ---- D step: real and generated batches go through D as ONE batch of 2B images.  The two
        # BCE means are taken over their own halves, so the gradients are exactly the sum of the
        # reference's two backward passes (g_lis/main.py:555-565) at half the kernel launches.
        self._set_dis_requires_grad(True)
        # both networks' gradients were consumed by the previous iteration's optimizer steps: cleared on the
        # side stream, under G's forward
        zeroed_d = self.dis_flat.zero_grad_async()
        zeroed_g = self.gen_flat.zero_grad_async()
        both, real_half, fake_half = self.paired.views(real)
        if real.data_ptr() != real_half.data_ptr():
            real_half.copy_(real)                   # (a caller that fills `paired` in place skips this)
        with torch.no_grad():
            fake, lis_d = gen(z_d, n_execute_lis_layers=depth_d, out=fake_half)
        if fake.data_ptr() != fake_half.data_ptr():
            fake_half.copy_(fake)                   # generator tail off the fused path: one copy
        for attr in ("_glis_unfolded", "_glis_planes"):
            both.__dict__.pop(attr, None)           # planes cached for the buffer's previous contents
        logits = dis_logits(dis, both)
        self._sync_begin("dis")
        if zeroed_d is not None:
            torch.cuda.current_stream().wait_event(zeroed_d)
        ops.Overlap.begin()
        if logits is not None:     # losses and d(loss)/d(logits) from one kernel per half; backward starts at the logits
            (loss_d_real, loss_d_fake), dl = bce_on_logits(logits, [1.0, 0.0], self.ls)
            logits.backward(dl)
        else:
            loss_d_real, loss_d_fake = dis_bce(dis, both, [1.0, 0.0], self.ls)
            (loss_d_real + loss_d_fake).backward()
        ops.Overlap.join()
        self.dis_flat.rebind_grads()
        gs = self._sync_done("dis", self.dis_flat)
        self.dis_flat.rmsprop_step(self.lr, self.alpha, self.eps, gs)
        # D's weight packs for its new parameters: rebuilt on the side stream under G's forward
        ops.refresh_packs(self.dis_flat, part="all", side=True)

        # ---- G step
        self._set_dis_requires_grad(False)
        fake, lis_g = gen(z_g, n_execute_lis_layers=depth_g)
        logits = dis_logits(dis, fake)
        loss_r = []
        self._sync_begin("gen")
        if zeroed_g is not None:
            torch.cuda.current_stream().wait_event(zeroed_g)
        ops.Overlap.begin()
        if logits is not None:
            # roots of backward: the logits and every LIS output, each with its kernel-computed gradient
            (loss_g,), dl = bce_on_logits(logits, [1.0], self.ls)
            roots, grads = [logits], [dl]
            if self.lambda_r > 0:
                zc = z_g.detach().contiguous()
                for i, u in enumerate(lis_g):
                    du = torch.empty_like(u, memory_format=torch.contiguous_format)
                    l = ops.mse_scaled(u, zc, self.lambda_r ** (i + 1), du)
                    loss_r.append(l.reshape(()))
                    roots.append(u)
                    grads.append(du)
            torch.autograd.backward(roots, grads)
        else:
            (loss_g,) = dis_bce(dis, fake, [1.0], self.ls)
            total = loss_g
            if self.lambda_r > 0:
                for i, u in enumerate(lis_g):
                    l = F.mse_loss(u, z_g) * (self.lambda_r ** (i + 1))
                    loss_r.append(l.detach())
                    total = total + l
            total.backward()
        ops.Overlap.join()
        self.gen_flat.rebind_grads()
        gs = self._sync_done("gen", self.gen_flat)
        self.gen_flat.rmsprop_step(self.lr, self.alpha, self.eps, gs)
        ops.refresh_packs(self.gen_flat, part="forward", side=False)   # ready for the next iteration's first kernel
        self._set_dis_requires_grad(True)

        return {"d_real": loss_d_real.detach(), "d_fake": loss_d_fake.detach(), "g": loss_g.detach(),
                "r": loss_r, "depth_d": len(lis_d), "depth_g": len(lis_g)}


class GraphedStep(object):
    """CUDA-graph replay of ``GLISTrainer.step``: one captured graph per LIS-depth pair.

    The step is ~200 short kernels; launched from Python it is bound by the host.  Captured
    once per ``(depth_d, depth_g)`` (the stochastic depth changes the kernel sequence, so each
    pair is its own graph — 4 graphs for one LIS module, 16 for three) and replayed from static
    input buffers, the host cost drops to one launch.  Losses come back as device tensors that
    stay valid until the next replay of the same graph.
    """

    def __init__(self, trainer, batch, height, width, code, device, warmup=2):
        self.tr = trainer
        # the static image input IS the first half of the trainer's 2B buffer: filling it in place costs no copy
        _, self.real, _ = trainer.paired.views(torch.empty(batch, 3, height, width, device=device).contiguous(
            memory_format=torch.channels_last))
        self.real.zero_()
        self.z_d = torch.zeros(batch, code, device=device)
        self.z_g = torch.zeros(batch, code, device=device)
        self.warmup = warmup
        self.graphs = {}
        self.pool = None

    def _capture(self, key):
        depth_d, depth_g = key
        tr = self.tr
        flats = [tr.gen_flat, tr.dis_flat]
        # the warm-up iterations (lazy CUDA init, allocator) must not count as training:
        # snapshot parameters, gradients and optimizer state, restore them before capturing
        saved = [(f.p.clone(), f.g.clone(), f.v.clone()) for f in flats]
        rng_state = tr.gen.rng.getstate() if hasattr(tr.gen.rng, "getstate") else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                tr.step(self.real, self.z_d, self.z_g, depth_d, depth_g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for f, (p, g, v) in zip(flats, saved):
            f.p.copy_(p); f.g.copy_(g); f.v.copy_(v)
        if rng_state is not None:
            tr.gen.rng.setstate(rng_state)
        # Weight packs live in persistent buffers and are rebuilt INSIDE the graph right after each
        # optimizer step; the iteration therefore starts from packs that match the parameters: build
        # them here for the restored parameters (eagerly, outside the capture).
        ops.bump_param_epoch()
        # (the iteration itself rebuilds G's data-gradient packs at its start: leave them stale)
        ops.refresh_packs(tr.gen_flat, part="forward", side=False)
        ops.refresh_packs(tr.dis_flat, part="all", side=False)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, pool=self.pool):
            out = tr.step(self.real, self.z_d, self.z_g, depth_d, depth_g)
        if self.pool is None:
            self.pool = graph.pool()
        ops.bump_param_epoch()
        self.graphs[key] = (graph, out)
        return self.graphs[key]

    def step(self, real=None, z_d=None, z_g=None, depth_d=None, depth_g=None):
        """Inputs left as ``None`` are taken from the static buffers as they are (the caller
        filled ``self.real`` / ``self.z_d`` / ``self.z_g`` in place, e.g. with the device RNG)."""
        if depth_d is None:
            depth_d = self.tr.gen.lis_depth(None)
        if depth_g is None:
            depth_g = self.tr.gen.lis_depth(None)
        key = (depth_d, depth_g)
        entry = self.graphs.get(key)
        if entry is None:
            if real is not None:
                self.real.copy_(real)
            if z_d is not None:
                self.z_d.copy_(z_d)
            if z_g is not None:
                self.z_g.copy_(z_g)
            entry = self._capture(key)     # capturing records the kernels, it does not run them
            entry[0].replay()
            ops.bump_param_epoch()
            return entry[1]
        if real is not None:
            self.real.copy_(real, non_blocking=True)
        if z_d is not None:
            self.z_d.copy_(z_d, non_blocking=True)
        if z_g is not None:
            self.z_g.copy_(z_g, non_blocking=True)
        entry[0].replay()
        ops.bump_param_epoch()   # the replay updated the parameters behind torch's back
        return entry[1]


class RIterTrainer(object):
    """The R-iterative trainer's outer iteration (r_iterative/main.py:428-535) on the sm_100a
    kernels: plain generator, reverser R and discriminator, three flat RMSprop states.

    Per hop r of the chain: G update, (r > 0) R update on
    ``λ^r·MSE(code, first_code) + (1-λ^r)·loss(dis(gen(code)), 1)``, D update on a fresh real batch and the hop's
    (pre-update) generated images as one 2B batch.  The stochastic ``do_train`` schedule (:445-451) is drawn from
    ``self.rng`` unless ``train_flags`` is given.  Losses and their gradients come from the fused loss kernels and
    every backward starts at the logits / the code (no autograd loss glue); weight gradients fork onto the side
    stream; under data parallelism each of the three networks has its own bucket set.  During the R
    update the generator's parameter gradients are not needed (the reference lets them pile up and zero-fills
    them before the next G step), so they are not computed.
    """

    def __init__(self, gen, rev, dis, lr, lambda_r=0.9, r_iterations=3, alpha=0.9, eps=1e-6, rng=None,
                 grad_sync=None, ls=False):
        import random as _random
        self.gen, self.rev, self.dis = gen, rev, dis
        self.lr, self.lambda_r, self.r_iterations, self.alpha, self.eps = lr, lambda_r, r_iterations, alpha, eps
        self.ls = bool(ls)
        self.gen_flat, self.rev_flat, self.dis_flat = FlatParams(gen), FlatParams(rev), FlatParams(dis)
        self.rng = rng if rng is not None else _random
        self.paired = PairedBatch()
        self.dropout = _has_dropout(gen, rev, dis)
        self.grad_sync = grad_sync
        self._overlapped = hasattr(grad_sync, "register")
        if self._overlapped:
            for tag, flat in (("gen", self.gen_flat), ("rev", self.rev_flat), ("dis", self.dis_flat)):
                grad_sync.register(tag, flat)

    def draw_train_flags(self, always_train_all=False):
        hops = 1 + self.r_iterations
        if always_train_all:
            return [True] * hops
        flags, last = [], False
        for r_idx in range(hops):
            hit = self.rng.random() <= (r_idx + 1) / float(hops)
            last = last or (r_idx == hops - 1) or hit
            flags.append(last)
        return flags

    @staticmethod
    def _requires_grad(flat, flag):
        for p in flat.params:
            p.requires_grad_(flag)

    def _backward_and_update(self, tag, flat, roots, grads):
        """Backward from ``roots`` (weight gradients on the side stream, gradient exchange overlapped), RMSprop on
        ``flat``, its weight packs rebuilt on the side stream under whatever runs next."""
        if self._overlapped:
            self.grad_sync.begin(tag)
        ops.Overlap.begin()
        torch.autograd.backward(roots, grads)
        ops.Overlap.join()
        flat.rebind_grads()
        if self.grad_sync is None:
            gs = 1.0
        elif self._overlapped:
            gs = self.grad_sync.finish(tag)
        else:
            gs = self.grad_sync(flat.g, tag)
        flat.rmsprop_step(self.lr, self.alpha, self.eps, gs)
        ops.refresh_packs(flat, part="all", side=True)

    def _adv(self, x, targets, gscale=1.0):
        """(losses, logits, d(sum of losses)/d(logits) * gscale) of ``dis`` on ``x`` — or (losses, None, None) with
        the losses differentiable through autograd when D has no standard head."""
        logits = dis_logits(self.dis, x)
        if logits is None:
            return dis_bce(self.dis, x, targets, self.ls), None, None
        losses, dl = bce_on_logits(logits, targets, self.ls, gscale)
        return losses, logits, dl

    def step(self, first_code, reals, train_flags=None):
        gen, rev = self.gen, self.rev
        hops = 1 + self.r_iterations
        if train_flags is None:
            train_flags = [True] * hops
        reals = list(reals)
        if self.dropout:
            ops.DropoutClock.tick(first_code.device)
        first = first_code.detach()
        out, last_images, last_code = [], None, None
        for r_idx in range(hops):
            code = first_code if last_images is None else rev(last_images.detach())
            if not train_flags[r_idx]:
                with torch.no_grad():
                    last_images = gen(code.detach())
                last_code = code
                out.append(None)
                continue
            rec = {}
            # ---- G (:475-485)
            self.gen_flat.zero_grad()
            self._requires_grad(self.dis_flat, False)
            generated = gen(code.detach())
            (loss_g,), logits, dl = self._adv(generated, [1.0])
            if logits is not None:
                self._backward_and_update("gen", self.gen_flat, [logits], [dl])
            else:
                self._backward_and_update("gen", self.gen_flat, [loss_g], [None])
            rec["g"] = loss_g.detach()
            # ---- R (:487-497), through the updated G and the frozen D
            if last_code is not None:
                self.rev_flat.zero_grad()
                self._requires_grad(self.gen_flat, False)
                lar = self.lambda_r ** r_idx
                (loss_g2,), logits, dl = self._adv(gen(code), [1.0], gscale=1.0 - lar)
                if logits is not None:
                    du = torch.empty_like(code, memory_format=torch.contiguous_format)
                    loss_r = ops.mse_scaled(code, first, lar, du, unscaled_loss=True).reshape(())
                    self._backward_and_update("rev", self.rev_flat, [logits, code], [dl, du])
                else:
                    loss_r = F.mse_loss(code, first)
                    self._backward_and_update("rev", self.rev_flat, [lar * loss_r + (1 - lar) * loss_g2], [None])
                self._requires_grad(self.gen_flat, True)
                rec["r"] = loss_r.detach()
            # ---- D (:502-526): the real batch and the hop's pre-update generated batch as one 2B batch
            self.dis_flat.zero_grad()
            self._requires_grad(self.dis_flat, True)
            real = reals.pop(0)
            both, real_half, fake_half = self.paired.views(real)
            if real.data_ptr() != real_half.data_ptr():
                real_half.copy_(real)
            fake_half.copy_(generated.detach())
            for attr in ("_glis_unfolded", "_glis_planes"):
                both.__dict__.pop(attr, None)
            (loss_d_real, loss_d_fake), logits, dl = self._adv(both, [1.0, 0.0])
            if logits is not None:
                self._backward_and_update("dis", self.dis_flat, [logits], [dl])
            else:
                self._backward_and_update("dis", self.dis_flat, [loss_d_real + loss_d_fake], [None])
            rec["d_real"], rec["d_fake"] = loss_d_real.detach(), loss_d_fake.detach()
            last_images, last_code = generated, code
            out.append(rec)
        if ops.Overlap.enabled and torch.cuda.is_available():
            # the last pack rebuild runs on the side stream: join it (a captured iteration must end on ONE stream)
            torch.cuda.current_stream().wait_stream(ops.Overlap.stream())
            for flat in (self.gen_flat, self.rev_flat, self.dis_flat):
                ops.forget_pending(flat)
        return out


class RSeparateTrainer(object):
    """The R-separate trainer's iteration (g_lis/train_r.py:406-436): a reverser trained alone on
    ``MSE(rev(gen(z)), z)`` against a FROZEN G-LIS, with the frozen D scoring the generated images before and after
    the round trip through R (two forward-only passes).  One flat RMSprop state (R's)."""

    def __init__(self, gen, rev, dis, lr, r_iterations, alpha=0.9, eps=1e-6, grad_sync=None, ls=False):
        self.gen, self.rev, self.dis = gen, rev, dis
        self.lr, self.depth, self.alpha, self.eps, self.ls = lr, r_iterations, alpha, eps, bool(ls)
        for net in (gen, dis):
            for p in net.parameters():
                p.requires_grad_(False)
        self.rev_flat = FlatParams(rev)
        self.dropout = _has_dropout(rev)
        self.grad_sync = grad_sync
        self._overlapped = hasattr(grad_sync, "register")
        if self._overlapped:
            grad_sync.register("rev", self.rev_flat)

    def _score(self, images):
        logits = dis_logits(self.dis, images)
        if logits is None:
            return dis_bce(self.dis, images, [0.0], self.ls)[0]
        return bce_on_logits(logits, [0.0], self.ls)[0][0]

    def step(self, z):
        if self.dropout:
            ops.DropoutClock.tick(z.device)
        with torch.no_grad():
            generated, _ = self.gen(z, n_execute_lis_layers=self.depth)
            stage1 = self._score(generated)
        self.rev_flat.zero_grad()
        code_fixed = self.rev(generated)
        du = torch.empty_like(code_fixed, memory_format=torch.contiguous_format)
        loss_r = ops.mse_scaled(code_fixed, z, 1.0, du).reshape(())
        if self._overlapped:
            self.grad_sync.begin("rev")
        ops.Overlap.begin()
        code_fixed.backward(du)
        ops.Overlap.join()
        self.rev_flat.rebind_grads()
        gs = 1.0 if self.grad_sync is None else (self.grad_sync.finish("rev") if self._overlapped
                                                 else self.grad_sync(self.rev_flat.g, "rev"))
        self.rev_flat.rmsprop_step(self.lr, self.alpha, self.eps, gs)
        ops.refresh_packs(self.rev_flat, part="all", side=False)
        with torch.no_grad():
            fixed, _ = self.gen(code_fixed.detach(), n_execute_lis_layers=self.depth)
            stage2 = self._score(fixed)
        return {"stage1": stage1.detach(), "r": loss_r.detach(), "stage2": stage2.detach()}


class GraphedRIter(object):
    """CUDA-graph replay of ``RIterTrainer.step``: one captured graph per ``do_train`` schedule (the sticky rule
    leaves ``1 + r_iterations`` distinct ones), replayed from static inputs ``first_code`` and ``reals[hop]``."""

    def __init__(self, trainer, batch, height, width, code, device, warmup=1):
        self.tr = trainer
        hops = 1 + trainer.r_iterations
        self.first_code = torch.zeros(batch, code, device=device)
        self.reals = [torch.zeros(batch, 3, height, width, device=device).contiguous(memory_format=torch.channels_last)
                      for _ in range(hops)]
        self.warmup, self.graphs, self.pool = warmup, {}, None

    def _capture(self, flags):
        tr = self.tr
        flats = [tr.gen_flat, tr.rev_flat, tr.dis_flat]
        run = lambda: tr.step(self.first_code, self.reals[:sum(flags)], list(flags))
        graph, out, self.pool = capture_step(flats, run, self.warmup, self.pool)
        self.graphs[flags] = (graph, out)
        return self.graphs[flags]

    def step(self, first_code=None, reals=None, train_flags=None):
        flags = tuple(bool(f) for f in (train_flags if train_flags is not None else self.tr.draw_train_flags()))
        if first_code is not None:
            self.first_code.copy_(first_code, non_blocking=True)
        for dst, src in zip(self.reals, reals or ()):
            dst.copy_(src, non_blocking=True)
        entry = self.graphs.get(flags) or self._capture(flags)
        entry[0].replay()
        ops.bump_param_epoch()
        return entry[1]


def capture_step(flats, run, warmup, pool):
    """Capture ``run()`` (one training iteration over the FlatParams ``flats``) into a CUDA graph.  The warm-up
    iterations (lazy CUDA init, allocator) must not count as training: parameters, gradients and optimizer state
    are snapshotted and restored, and the weight packs — which live in persistent buffers and are rebuilt INSIDE
    the graph after each optimizer step — are rebuilt eagerly for the restored parameters before the capture.
    Returns (graph, outputs of the captured run, memory pool)."""
    saved = [(f.p.clone(), f.g.clone(), f.v.clone()) for f in flats]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for f, (p, g, v) in zip(flats, saved):
        f.p.copy_(p); f.g.copy_(g); f.v.copy_(v)
    ops.bump_param_epoch()
    for f in flats:
        ops.refresh_packs(f, part="all", side=False)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, pool=pool):
        out = run()
    ops.bump_param_epoch()
    return graph, out, (pool if pool is not None else graph.pool())


class HostFedStepper(object):
    """Drives ``GraphedStep`` from HOST batches without ever idling the device.

    ``submit(real, z_d, z_g)`` takes pinned host tensors and (1) enqueues their host-to-device copy
    on a copy stream into one of two staging sets — it overlaps the previous iteration, which is
    still running on the main stream; (2) enqueues the training iteration (device-to-device copy
    into the graph's static inputs, graph replay) behind the copy; (3) enqueues an asynchronous
    read-back of the iteration's losses into pinned memory; (4) waits for the PREVIOUS
    iteration's losses — normally already there — and returns them as Python floats (``None`` on
    the first call).  ``flush()`` returns the last iteration's losses.  Every iteration's inputs
    cross PCIe and every iteration's losses are read on the host; only the waiting is pipelined.
    """

    def __init__(self, graphed):
        self.g = graphed
        dev = graphed.real.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [(torch.empty_like(graphed.real), torch.empty_like(graphed.z_d), torch.empty_like(graphed.z_g))
                      for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]     # H2D of set i finished
        self.consumed = [torch.cuda.Event() for _ in range(2)]   # set i has been read by its iteration
        self.loss_ready = [torch.cuda.Event() for _ in range(2)]
        self.h_loss = [torch.zeros(8).pin_memory() for _ in range(2)]
        self.n_r = [0, 0]
        self.count = 0
        self.h2d_bytes = sum(t.numel() * 4 for t in self.stage[0])

    def _read(self, slot):
        self.loss_ready[slot].synchronize()
        v = self.h_loss[slot].tolist()
        return {"d_real": v[0], "d_fake": v[1], "g": v[2], "r": v[3:3 + self.n_r[slot]]}

    def submit(self, real, z_d, z_g, depth_d=None, depth_g=None):
        slot = self.count & 1
        main = torch.cuda.current_stream()
        if self.count >= 2:
            self.copy_stream.wait_event(self.consumed[slot])     # the iteration two back has read this set
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(self.stage[slot], (real, z_d, z_g)):
                dst.copy_(src, non_blocking=True)
            self.copied[slot].record(self.copy_stream)
        main.wait_event(self.copied[slot])
        out = self.g.step(self.stage[slot][0], self.stage[slot][1], self.stage[slot][2], depth_d, depth_g)
        self.consumed[slot].record(main)
        vals = [out["d_real"], out["d_fake"], out["g"]] + list(out["r"])[:5]
        self.n_r[slot] = len(vals) - 3
        self.h_loss[slot][:len(vals)].copy_(torch.stack([v.reshape(()) for v in vals]), non_blocking=True)
        self.loss_ready[slot].record(main)
        self.count += 1
        return self._read(slot ^ 1) if self.count >= 2 else None

    def flush(self):
        return self._read((self.count - 1) & 1) if self.count else None
