#!/usr/bin/env python
"""Generate the golden vectors that pin the oracle to the reference's OWN sources.

Run in the build container (needs ``/root/reference``; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference (python-2.7 / mid-2017 PyTorch) does not import under torch 2.x, so its
module sources are executed here under a *legacy-semantics shim* that changes nothing
in the arithmetic it asks PyTorch for:

* ``Tensor.sum(int)`` keeps the reduced dimension (the 2017 default the code relies on);
* ``_ConvNd`` accepts the 2017 10-argument constructor; ``_ConvTransposeMixin`` provides
  the 2-argument ``_output_padding``;
* ``Module.add_module`` accepts dotted names (stored with '·', exported back as '.');
* leading tabs are expanded (``TPReLU.py`` mixes tabs and spaces — a py3 TabError).

What is written (all float64, tiny shapes, ``.npz``):

* ``modules.npz``   – each WN layer / TPReLU: params, input, output, all gradients;
* ``models.npz``    – D, R, plain G, G-LIS (incl. a pad-triggering non-square size and the
                      ``nearest`` upscaling variant): state_dict, input, output, param grads;
* ``glis_steps.npz``– three full training iterations (g_lis/main.py:537-589 semantics) on
                      reference-built G-LIS + D with stock ``torch.optim.RMSprop`` /
                      ``nn.BCELoss`` / ``nn.MSELoss``: losses and every parameter after
                      each iteration, with stochastic LIS depth forced per iteration
                      (including a skipped module, exercising the zero-fill rule);
* ``glis_steps_ls.npz`` – two such iterations with ``--ls`` (``lossfunc = nn.MSELoss()`` on D's sigmoid
                      output, g_lis/main.py:308-311);
* ``rsep_steps.npz``  – two iterations of the R-separate trainer (g_lis/train_r.py:406-436): reverser trained alone
                      against reference-built frozen G-LIS + D;
* ``riter_steps.npz`` – two outer iterations of the R-iterative trainer (r_iterative/main.py:428-535) on
                      reference-built plain G + reverser R + D with three stock RMSprops: all hops trained,
                      then the schedule [skip, train, train].
"""
import contextlib
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("GLIS_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
DOT = "·"


# ----------------------------------------------------------------------------- shim
@contextlib.contextmanager
def legacy_torch():
    from torch.nn.modules import conv as convmod

    orig_sum = torch.Tensor.sum
    orig_convnd = convmod._ConvNd
    orig_mixin = getattr(convmod, "_ConvTransposeMixin", None)
    orig_add = nn.Module.add_module

    def sum_keepdim(self, *args, **kw):
        if len(args) == 1 and isinstance(args[0], int) and not kw:
            return orig_sum(self, args[0], keepdim=True)
        return orig_sum(self, *args, **kw)

    class LegacyConvNd(orig_convnd):
        def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation,
                     transposed, output_padding, groups, bias):
            super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation,
                             transposed, output_padding, groups, bias, "zeros")

        def _conv_forward(self, *a, **k):  # abstract in modern torch; unused by the reference
            raise NotImplementedError

    class LegacyTransposeMixin(object):
        def _output_padding(self, input, output_size):
            if output_size is None:
                return self.output_padding
            raise NotImplementedError("golden vectors do not use output_size")

    def add_dotted(self, name, module):
        return orig_add(self, name.replace(".", DOT), module)

    torch.Tensor.sum = sum_keepdim
    convmod._ConvNd = LegacyConvNd
    convmod._ConvTransposeMixin = LegacyTransposeMixin
    nn.Module.add_module = add_dotted
    try:
        yield
    finally:
        torch.Tensor.sum = orig_sum
        convmod._ConvNd = orig_convnd
        if orig_mixin is not None:
            convmod._ConvTransposeMixin = orig_mixin
        nn.Module.add_module = orig_add


def _exec_reference(relpath, modname, extra=None):
    with open(os.path.join(REF, relpath)) as fh:
        src = fh.read().expandtabs(4)
    mod = types.ModuleType(modname)
    mod.__file__ = os.path.join(REF, relpath)
    if extra:
        mod.__dict__.update(extra)
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def load_reference():
    """Execute the reference's module + model sources; returns the model namespace."""
    mods = {}
    for name in ("WeightNormalizedConv", "WeightNormalizedLinear", "TPReLU", "View"):
        mods[name] = _exec_reference("common/modules/%s.py" % name, "modules." + name)
    pkg = types.ModuleType("modules")
    pkg.__path__ = []
    saved = {k: sys.modules.get(k) for k in ["modules"] + ["modules." + n for n in mods]}
    sys.modules["modules"] = pkg
    for n, m in mods.items():
        sys.modules["modules." + n] = m
        setattr(pkg, n, m)
    try:
        model = _exec_reference("common/model.py", "ref_model")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return model


def undot(state):
    return {k.replace(DOT, "."): v for k, v in state.items()}


# ----------------------------------------------------------------------------- helpers
def t64(*shape, lo=-1.0, hi=1.0, gen=None):
    return (torch.rand(*shape, dtype=torch.float64, generator=gen) * (hi - lo) + lo)


def randomize_(module, gen):
    """Give every parameter a non-trivial value (scale/bias/TPReLU included)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            leaf = name.split(".")[-1]
            if leaf == "weight" and p.dim() == 1:          # TPReLU slope: inside and outside [0,1]
                p.copy_(t64(*p.shape, lo=-0.2, hi=1.2, gen=gen))
            elif leaf == "bias" and p.dim() == 1:          # TPReLU translation
                p.copy_(t64(*p.shape, lo=-0.3, hi=0.3, gen=gen))
            elif leaf == "scale":
                p.copy_(t64(*p.shape, lo=0.5, hi=1.5, gen=gen))
            elif leaf == "bias":
                p.copy_(t64(*p.shape, lo=-0.2, hi=0.2, gen=gen))
            else:
                p.copy_(t64(*p.shape, lo=-0.5, hi=0.5, gen=gen))


def put(store, prefix, **arrays):
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        store["%s/%s" % (prefix, k)] = np.array(v, copy=True)


def run_module(store, prefix, module, x, gen):
    module.double()
    randomize_(module, gen)
    x = x.clone().requires_grad_(True)
    y = module(x)
    r = t64(*y.shape, gen=gen)
    (y * r).sum().backward()
    put(store, prefix, x=x, y=y, r=r, dx=x.grad)
    for n, p in module.named_parameters():
        put(store, prefix, **{"p." + n: p, "g." + n: p.grad})


def run_model(store, prefix, net, x, gen, forward=None):
    net.double()
    randomize_(net, gen)
    net.eval()  # no dropout anywhere in these fixtures; LIS depth is forced explicitly
    x = x.clone().requires_grad_(True)
    y = forward(net, x) if forward else net(x)
    outs = y if isinstance(y, (tuple, list)) else (y,)
    flat = [outs[0]] + list(outs[1]) if len(outs) > 1 else [outs[0]]
    loss = 0
    for i, o in enumerate(flat):
        r = t64(*o.shape, gen=gen)
        put(store, prefix, **{"y%d" % i: o, "r%d" % i: r})
        loss = loss + (o * r).sum()
    loss.backward()
    put(store, prefix, x=x, dx=x.grad, n_out=len(flat))
    grads = {n.replace(DOT, "."): p.grad for n, p in net.named_parameters()}
    for k, v in undot(net.state_dict()).items():
        put(store, prefix, **{"p." + k: v})
        put(store, prefix, **{"g." + k: grads[k] if grads.get(k) is not None else torch.zeros_like(v)})


# ----------------------------------------------------------------------------- main
def main():
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(20260101)
    with legacy_torch():
        ref = load_reference()

        # ---- module-level vectors
        mods = {}
        run_module(mods, "conv_s2", ref.WeightNormalizedConv2d(3, 5, 4, 2, (1, 1), scale=False, bias=False),
                   t64(2, 3, 8, 12, gen=gen), gen)
        run_module(mods, "conv_s2_pad2_affine", ref.WeightNormalizedConv2d(4, 6, 4, 2, (2, 1)),
                   t64(2, 4, 10, 8, gen=gen), gen)
        run_module(mods, "conv_head", ref.WeightNormalizedConv2d(6, 1, (3, 5)),
                   t64(3, 6, 3, 5, gen=gen), gen)
        run_module(mods, "conv_3x3_s1", ref.WeightNormalizedConv2d(4, 3, 3, 1, (1, 1), scale=False, bias=False),
                   t64(2, 4, 6, 6, gen=gen), gen)
        run_module(mods, "deconv_s2", ref.WeightNormalizedConvTranspose2d(6, 4, 4, 2, (1, 1), scale=False, bias=False),
                   t64(2, 6, 5, 3, gen=gen), gen)
        run_module(mods, "deconv_s2_pad2_affine", ref.WeightNormalizedConvTranspose2d(4, 3, 4, 2, (2, 1)),
                   t64(2, 4, 5, 4, gen=gen), gen)
        run_module(mods, "linear_plain", ref.WeightNormalizedLinear(7, 10, scale=False, bias=False, init_factor=0.01),
                   t64(4, 7, gen=gen), gen)
        run_module(mods, "linear_affine", ref.WeightNormalizedLinear(7, 5),
                   t64(4, 7, gen=gen), gen)
        run_module(mods, "tprelu_2d", ref.TPReLU(6), t64(5, 6, gen=gen), gen)
        run_module(mods, "tprelu_4d", ref.TPReLU(3), t64(2, 3, 4, 5, gen=gen), gen)
        np.savez_compressed(os.path.join(OUT, "modules.npz"), **mods)

        # ---- model-level vectors
        models = {}
        run_model(models, "D_16", ref.build_discriminator(16, 16, 4, 2, "weight", 0), t64(3, 3, 16, 16, lo=0, gen=gen), gen)
        # 20 wide x 12 high, 3 levels: width pads at level 1 (10 % 4 == 2), height never
        run_model(models, "D_20x12_pad", ref.build_discriminator(20, 12, 4, 3, "weight", 0),
                  t64(2, 3, 12, 20, lo=0, gen=gen), gen)
        run_model(models, "R_16", ref.build_reverser(16, 16, 4, 2, 8, "weight", 0), t64(3, 3, 16, 16, lo=0, gen=gen), gen)
        run_model(models, "G_16", ref.build_generator(16, 16, 4, 2, 8, "weight"), t64(3, 8, gen=gen), gen)
        run_model(models, "G_20x12_pad", ref.build_generator(20, 12, 4, 3, 8, "weight"), t64(2, 8, gen=gen), gen)
        run_model(models, "GLIS_16_k2of3",
                  ref.GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", n_lis_layers=3, upscaling="fractional"),
                  t64(3, 8, gen=gen), gen, forward=lambda n, x: n(x, n_execute_lis_layers=2))
        run_model(models, "GLIS_16_nearest",
                  ref.GeneratorLearnedInputSpace(16, 16, 4, 3, 8, "weight", n_lis_layers=1, upscaling="nearest"),
                  t64(2, 8, gen=gen), gen, forward=lambda n, x: n(x, n_execute_lis_layers="all"))
        run_model(models, "D_16_affine", ref.build_discriminator(16, 16, 4, 2, "weight-affine", 0),
                  t64(2, 3, 16, 16, lo=0, gen=gen), gen)
        np.savez_compressed(os.path.join(OUT, "models.npz"), **models)

        # ---- three training iterations on reference-built nets with stock torch optimizers
        steps = {}
        W = H = 16
        B, code, nf, nl, n_lis = 4, 8, 4, 2, 2
        lr, lam = 1e-2, 0.9
        torch.manual_seed(7)
        g = ref.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis_layers=n_lis, upscaling="fractional").double()
        d = ref.build_discriminator(W, H, nf, nl, "weight", 0).double()
        g.train(); d.train()
        bce, mse = nn.BCELoss(), nn.MSELoss()
        g_opt = torch.optim.RMSprop(g.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        d_opt = torch.optim.RMSprop(d.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        for k, v in undot(g.state_dict()).items():
            put(steps, "init/g", **{k: v})
        for k, v in undot(d.state_dict()).items():
            put(steps, "init/d", **{k: v})
        ones, zeros = torch.ones(B, 1, dtype=torch.float64), torch.zeros(B, 1, dtype=torch.float64)
        depths = [(1, 0), (2, 1), (0, 2)]  # (D-fake forward, G forward); module 1 first trained at it 2
        put(steps, "cfg", W=W, H=H, B=B, code=code, nf=nf, nl=nl, n_lis=n_lis, lr=lr, lam=lam,
            depths=np.array(depths))
        for it, (kd, kg) in enumerate(depths):
            real, zd, zg = t64(B, 3, H, W, lo=0, gen=gen), torch.randn(B, code, dtype=torch.float64, generator=gen), \
                torch.randn(B, code, dtype=torch.float64, generator=gen)
            for p in d.parameters():
                p.requires_grad = True
            d.zero_grad(set_to_none=False)
            l_real = bce(d(real), ones); l_real.backward()
            with torch.no_grad():
                fake, _ = g(zd, n_execute_lis_layers=kd)
            l_fake = bce(d(fake.detach()), zeros); l_fake.backward()
            d_opt.step()
            for p in d.parameters():
                p.requires_grad = False
            g.zero_grad(set_to_none=False)
            fake, lis = g(zg, n_execute_lis_layers=kg)
            l_g = bce(d(fake), ones)
            l_g.backward(retain_graph=len(lis) > 0)
            l_r = []
            for i, u in enumerate(lis):
                l = mse(u, zg) * (lam ** (i + 1))
                l.backward(retain_graph=(i + 1) < len(lis))
                l_r.append(l.item())
            g_opt.step()
            pre = "it%d" % it
            put(steps, pre, real=real, zd=zd, zg=zg, d_real=l_real.item(), d_fake=l_fake.item(),
                g=l_g.item(), r=np.array(l_r, dtype=np.float64))
            for k, v in undot(g.state_dict()).items():
                put(steps, pre + "/g", **{k: v})
            for k, v in undot(d.state_dict()).items():
                put(steps, pre + "/d", **{k: v})
        np.savez_compressed(os.path.join(OUT, "glis_steps.npz"), **steps)

        # ---- --ls: the same iteration with nn.MSELoss() as the adversarial loss (g_lis/main.py:308-311)
        steps = {}
        torch.manual_seed(8)
        g = ref.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis_layers=n_lis, upscaling="fractional").double()
        d = ref.build_discriminator(W, H, nf, nl, "weight", 0).double()
        g.train(); d.train()
        lsq, mse = nn.MSELoss(), nn.MSELoss()
        g_opt = torch.optim.RMSprop(g.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        d_opt = torch.optim.RMSprop(d.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        for k, v in undot(g.state_dict()).items():
            put(steps, "init/g", **{k: v})
        for k, v in undot(d.state_dict()).items():
            put(steps, "init/d", **{k: v})
        depths = [(2, 2), (1, 0)]
        put(steps, "cfg", W=W, H=H, B=B, code=code, nf=nf, nl=nl, n_lis=n_lis, lr=lr, lam=lam, depths=np.array(depths))
        for it, (kd, kg) in enumerate(depths):
            real, zd, zg = t64(B, 3, H, W, lo=0, gen=gen), torch.randn(B, code, dtype=torch.float64, generator=gen), \
                torch.randn(B, code, dtype=torch.float64, generator=gen)
            for p in d.parameters():
                p.requires_grad = True
            d.zero_grad(set_to_none=False)
            l_real = lsq(d(real), ones); l_real.backward()
            with torch.no_grad():
                fake, _ = g(zd, n_execute_lis_layers=kd)
            l_fake = lsq(d(fake.detach()), zeros); l_fake.backward()
            d_opt.step()
            for p in d.parameters():
                p.requires_grad = False
            g.zero_grad(set_to_none=False)
            fake, lis = g(zg, n_execute_lis_layers=kg)
            l_g = lsq(d(fake), ones)
            l_g.backward(retain_graph=len(lis) > 0)
            l_r = []
            for i, u in enumerate(lis):
                l = mse(u, zg) * (lam ** (i + 1))
                l.backward(retain_graph=(i + 1) < len(lis))
                l_r.append(l.item())
            g_opt.step()
            pre = "it%d" % it
            put(steps, pre, real=real, zd=zd, zg=zg, d_real=l_real.item(), d_fake=l_fake.item(),
                g=l_g.item(), r=np.array(l_r, dtype=np.float64))
            for k, v in undot(g.state_dict()).items():
                put(steps, pre + "/g", **{k: v})
            for k, v in undot(d.state_dict()).items():
                put(steps, pre + "/d", **{k: v})
        np.savez_compressed(os.path.join(OUT, "glis_steps_ls.npz"), **steps)

        # ---- R-iterative outer iterations (r_iterative/main.py:428-535; nets :195-215)
        steps = {}
        R = 2
        torch.manual_seed(9)
        g = ref.build_generator(W, H, nf, nl, code, "weight").double()
        rv = ref.build_reverser(W, H, nf // 2, nl, code, "weight", 0).double()
        d = ref.build_discriminator(W, H, nf, nl, "weight", 0).double()   # (:205 passes 5 arguments: a TypeError upstream)
        for net in (g, rv, d):
            net.train()
        lossfunc, lossfunc_r = nn.BCELoss(), nn.MSELoss()
        g_opt = torch.optim.RMSprop(g.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        r_opt = torch.optim.RMSprop(rv.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        d_opt = torch.optim.RMSprop(d.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        for tag, net in (("g", g), ("r", rv), ("d", d)):
            for k, v in undot(net.state_dict()).items():
                put(steps, "init/" + tag, **{k: v})
        schedules = [[True, True, True], [False, True, True]]
        put(steps, "cfg", W=W, H=H, B=B, code=code, nf=nf, nl=nl, R=R, lr=lr, lam=lam,
            schedules=np.array(schedules, dtype=np.int64))
        for it, flags in enumerate(schedules):
            pre = "it%d" % it
            first_code = last_code = last_images = None
            z = torch.randn(B, code, dtype=torch.float64, generator=gen)
            put(steps, pre, z=z)
            n_real = 0
            for r_idx in range(1 + R):
                do_train = flags[r_idx]
                if last_images is None:
                    codev = z
                    first_code = codev
                else:
                    codev = rv(last_images.detach())
                if not do_train:
                    last_images = g(codev.detach())
                    last_code = codev
                    continue
                g.zero_grad(set_to_none=False)
                for p in d.parameters():
                    p.requires_grad = False
                generated = g(codev.detach())
                loss_g = lossfunc(d(generated), ones)
                loss_g.backward()
                g_opt.step()
                rec = {"g": loss_g.item()}
                if last_code is not None:
                    rv.zero_grad(set_to_none=False)
                    loss_g2 = lossfunc(d(g(codev)), ones)
                    loss_r = lossfunc_r(codev, first_code.detach())
                    lar = lam ** r_idx
                    (lar * loss_r + (1 - lar) * loss_g2).backward()
                    r_opt.step()
                    rec["r"] = loss_r.item()
                d.zero_grad(set_to_none=False)
                for p in d.parameters():
                    p.requires_grad = True
                real = t64(B, 3, H, W, lo=0, gen=gen)
                put(steps, pre, **{"real%d" % n_real: real})
                n_real += 1
                loss_d_real = lossfunc(d(real), ones)
                loss_d_real.backward()
                loss_d_fake = lossfunc(d(generated.detach()), zeros)
                loss_d_fake.backward()
                d_opt.step()
                rec["d_real"], rec["d_fake"] = loss_d_real.item(), loss_d_fake.item()
                put(steps, pre + "/hop%d" % r_idx, **rec)
                last_images, last_code = generated, codev
            for tag, net in (("g", g), ("r", rv), ("d", d)):
                for k, v in undot(net.state_dict()).items():
                    put(steps, pre + "/" + tag, **{k: v})
        np.savez_compressed(os.path.join(OUT, "riter_steps.npz"), **steps)

        # ---- R-separate iterations (g_lis/train_r.py:406-436; nets :205-222)
        steps = {}
        torch.manual_seed(10)
        g = ref.GeneratorLearnedInputSpace(W, H, nf, nl, code, "weight", n_lis_layers=n_lis, upscaling="fractional").double()
        rv = ref.build_reverser(W, H, nf // 2, nl, code, "weight", 0).double()
        d = ref.build_discriminator(W, H, nf, nl, "weight", 0).double()
        for net in (g, rv, d):
            net.train()
        for p in list(g.parameters()) + list(d.parameters()):
            p.requires_grad = False
        r_opt = torch.optim.RMSprop(rv.parameters(), lr=lr, eps=1e-6, alpha=0.9)
        lossfunc, lossfunc_r = nn.BCELoss(), nn.MSELoss()
        for tag, net in (("g", g), ("r", rv), ("d", d)):
            for k, v in undot(net.state_dict()).items():
                put(steps, "init/" + tag, **{k: v})
        put(steps, "cfg", W=W, H=H, B=B, code=code, nf=nf, nl=nl, n_lis=n_lis, lr=lr, iters=2)
        for it in range(2):
            pre = "it%d" % it
            z = torch.randn(B, code, dtype=torch.float64, generator=gen)
            generated, _ = g(z, n_execute_lis_layers=n_lis)
            with torch.no_grad():
                l1 = lossfunc(d(generated.detach()), zeros)
            rv.zero_grad(set_to_none=False)
            code_fixed = rv(generated.detach())
            loss_r = lossfunc_r(code_fixed, z)
            loss_r.backward()
            r_opt.step()
            with torch.no_grad():
                fixed, _ = g(code_fixed.detach(), n_execute_lis_layers=n_lis)
                l2 = lossfunc(d(fixed), zeros)
            put(steps, pre, z=z, stage1=l1.item(), r=loss_r.item(), stage2=l2.item())
            for k, v in undot(rv.state_dict()).items():
                put(steps, pre + "/r", **{k: v})
        np.savez_compressed(os.path.join(OUT, "rsep_steps.npz"), **steps)
    for f in ("modules.npz", "models.npz", "glis_steps.npz", "glis_steps_ls.npz", "riter_steps.npz", "rsep_steps.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
