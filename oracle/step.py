"""ORACLE (test infrastructure) — one G-LIS adversarial training iteration, restated.

Follows ``g_lis/main.py:526-589`` of the reference (losses :308-312, optimizers
:313-314, constant targets :508-510) as a function of explicit inputs, with the
legacy semantics of SURVEY.md App. B:

* ``zero_grad`` zero-fills existing gradients (B.4): a parameter that has received a
  gradient once keeps decaying its RMSprop ``square_avg`` on every later step even when
  the stochastic LIS depth skipped it; a parameter that never received one is skipped.
* RMSprop (B.5): ``v = a v + (1-a) g^2 ; p -= lr g / (sqrt(v) + eps)``, a=.9, eps=1e-6.
* BCE (B.6): modern form (log clamped at -100); equal to the 2017 kernel away from
  saturation.  MSE (B.7): mean over all elements, times ``lambda_r**(i+1)``.
* The separate ``.backward()`` calls of :579 and :586 add into the same ``.grad``
  fields, which is what happens here too.
"""
import torch
import torch.nn.functional as F


def rmsprop_update(params, state, lr, alpha=0.9, eps=1e-6):
    """In-place legacy RMSprop over ``params`` (skips ``grad is None``); ``state`` maps
    parameter -> square_avg.  g_lis/main.py:313-314 / torch.optim.RMSprop (no momentum)."""
    with torch.no_grad():
        for p in params:
            if p.grad is None:
                continue
            v = state.get(p)
            if v is None:
                v = state[p] = torch.zeros_like(p)
            g = p.grad
            v.mul_(alpha).addcmul_(g, g, value=1.0 - alpha)
            p.addcdiv_(g, v.sqrt().add_(eps), value=-lr)


def _zero_fill(module):
    for p in module.parameters():
        if p.grad is not None:
            p.grad.detach_()
            p.grad.zero_()


def adversarial_loss(ls):
    """``lossfunc`` of g_lis/main.py:308-311 and r_iterative/main.py:208-211: ``nn.MSELoss()`` on D's
    (sigmoid) output under ``--ls``, else ``nn.BCELoss()``."""
    return F.mse_loss if ls else F.binary_cross_entropy


def glis_iteration(gen, dis, gen_state, dis_state, real, z_d, z_g, lr, lambda_r=0.9,
                   depth_d=None, depth_g=None, alpha=0.9, eps=1e-6, ls=False):
    """One iteration: D on real, D on fake (G under no_grad), D update, G+LIS update.

    ``depth_d`` / ``depth_g`` force the number of LIS modules run in the D-fake and G
    forwards (``None`` = draw from ``gen.rng`` exactly as the reference does).
    Returns a dict of python floats: d_real, d_fake, g, r (list), depth_d, depth_g.
    """
    B = real.size(0)
    ones = torch.ones(B, 1, dtype=real.dtype)
    zeros = torch.zeros(B, 1, dtype=real.dtype)
    lossfunc = adversarial_loss(ls)

    # ---- D step (:537-568)
    for p in dis.parameters():
        p.requires_grad_(True)
    _zero_fill(dis)
    loss_d_real = lossfunc(dis(real), ones)
    loss_d_real.backward()
    with torch.no_grad():
        fake, lis_d = gen(z_d, n_execute_lis_layers=depth_d)
    loss_d_fake = lossfunc(dis(fake.detach()), zeros)
    loss_d_fake.backward()
    rmsprop_update(list(dis.parameters()), dis_state, lr, alpha, eps)

    # ---- G step (:571-589)
    for p in dis.parameters():
        p.requires_grad_(False)
    _zero_fill(gen)
    fake, lis_g = gen(z_g, n_execute_lis_layers=depth_g)
    loss_g = lossfunc(dis(fake), ones)
    loss_g.backward(retain_graph=(lambda_r > 0 and len(lis_g) > 0))
    loss_r = []
    if lambda_r > 0:
        for i, u in enumerate(lis_g):
            l = F.mse_loss(u, z_g) * (lambda_r ** (i + 1))
            l.backward(retain_graph=(i + 1) < len(lis_g))
            loss_r.append(l.item())
    rmsprop_update(list(gen.parameters()), gen_state, lr, alpha, eps)
    for p in dis.parameters():
        p.requires_grad_(True)

    return {"d_real": loss_d_real.item(), "d_fake": loss_d_fake.item(), "g": loss_g.item(),
            "r": loss_r, "depth_d": len(lis_d), "depth_g": len(lis_g)}


class GLISOracleTrainer:
    """Holds G, D and both RMSprop states; ``step`` = :func:`glis_iteration`."""

    def __init__(self, gen, dis, lr, lambda_r=0.9, alpha=0.9, eps=1e-6, ls=False):
        self.gen, self.dis = gen, dis
        self.lr, self.lambda_r, self.alpha, self.eps, self.ls = lr, lambda_r, alpha, eps, ls
        self.gen_state, self.dis_state = {}, {}

    def step(self, real, z_d, z_g, depth_d=None, depth_g=None):
        return glis_iteration(self.gen, self.dis, self.gen_state, self.dis_state, real, z_d, z_g,
                              self.lr, self.lambda_r, depth_d, depth_g, self.alpha, self.eps, self.ls)


def riter_iteration(gen, rev, dis, gen_state, rev_state, dis_state, first_code, reals, lr, lambda_r=0.9,
                    r_iterations=3, train_flags=None, alpha=0.9, eps=1e-6, ls=False):
    """One outer iteration of the R-iterative trainer, restated from ``r_iterative/main.py:428-535``.

    A chain of ``1 + r_iterations`` hops: hop 0 starts from ``first_code`` (noise), hop r > 0 from
    ``rev(images of hop r-1)``.  A hop with ``train_flags[r]`` false only regenerates images
    (:461-470); a trained hop does, in this order, a G update on ``BCE(dis(gen(code.detach())), 1)``
    (:475-485), for r > 0 an R update on ``λ^r·MSE(code, first_code) + (1-λ^r)·BCE(dis(gen(code)), 1)``
    with the already updated G (:487-497), and a D update on a fresh real batch and the hop's
    (pre-update) generated images (:502-526).  ``train_flags=None`` means ``--always_train_all``.
    ``reals``: one real batch per trained hop, in order.  Returns per-hop loss dicts (None if skipped).
    """
    B = first_code.size(0)
    ones = torch.ones(B, 1, dtype=first_code.dtype)
    zeros = torch.zeros(B, 1, dtype=first_code.dtype)
    hops = 1 + r_iterations
    if train_flags is None:
        train_flags = [True] * hops
    reals = list(reals)
    lossfunc = adversarial_loss(ls)
    out = []
    last_images, last_code = None, None
    for r_idx in range(hops):
        code = first_code if last_images is None else rev(last_images.detach())
        if not train_flags[r_idx]:
            with torch.no_grad():      # (:468: the images are only ever used detached, :459)
                last_images = gen(code.detach())
            last_code = code
            out.append(None)
            continue
        rec = {}
        # ---- G
        _zero_fill(gen)
        for p in dis.parameters():
            p.requires_grad_(False)
        generated = gen(code.detach())
        loss_g = lossfunc(dis(generated), ones)
        loss_g.backward()
        rmsprop_update(list(gen.parameters()), gen_state, lr, alpha, eps)
        rec["g"] = loss_g.item()
        # ---- R
        if last_code is not None:
            _zero_fill(rev)
            loss_g2 = lossfunc(dis(gen(code)), ones)
            loss_r = F.mse_loss(code, first_code.detach())
            lar = lambda_r ** r_idx
            (lar * loss_r + (1 - lar) * loss_g2).backward()
            rmsprop_update(list(rev.parameters()), rev_state, lr, alpha, eps)
            rec["r"] = loss_r.item()
        # ---- D
        _zero_fill(dis)
        for p in dis.parameters():
            p.requires_grad_(True)
        loss_d_real = lossfunc(dis(reals.pop(0)), ones)
        loss_d_real.backward()
        loss_d_fake = lossfunc(dis(generated.detach()), zeros)
        loss_d_fake.backward()
        rmsprop_update(list(dis.parameters()), dis_state, lr, alpha, eps)
        rec["d_real"], rec["d_fake"] = loss_d_real.item(), loss_d_fake.item()
        last_images, last_code = generated, code
        out.append(rec)
    return out


def rsep_iteration(gen, rev, dis, rev_state, z, lr, depth, alpha=0.9, eps=1e-6, ls=False):
    """One iteration of the R-separate trainer, restated from ``g_lis/train_r.py:406-436``: a reverser R is trained
    ALONE to recover the latent code of images a frozen G-LIS generates, and the frozen D scores the images before
    and after the round trip.

    ``generated = gen(z)`` with ``depth`` LIS modules (:410); stage-1 loss = lossfunc(dis(generated), zeros) without
    gradients (:413-417); R update on ``MSE(r(generated), z)`` (:420-427 — the reference forgets to keep the
    result of ``generated.detach()`` (:418), so its backward also walks the frozen generator; no parameter but R's
    is ever stepped, which is what is restated); stage-2 loss = lossfunc(dis(gen(r(generated))), zeros) (:430-435).
    Returns {"stage1", "r", "stage2"} as floats."""
    B = z.size(0)
    zeros = torch.zeros(B, 1, dtype=z.dtype)
    lossfunc = adversarial_loss(ls)
    with torch.no_grad():
        generated, _ = gen(z, n_execute_lis_layers=depth)
        stage1 = lossfunc(dis(generated), zeros).item()
    _zero_fill(rev)
    code_fixed = rev(generated)
    loss_r = F.mse_loss(code_fixed, z)
    loss_r.backward()
    rmsprop_update(list(rev.parameters()), rev_state, lr, alpha, eps)
    with torch.no_grad():
        fixed, _ = gen(code_fixed.detach(), n_execute_lis_layers=depth)
        stage2 = lossfunc(dis(fixed), zeros).item()
    return {"stage1": stage1, "r": loss_r.item(), "stage2": stage2}
