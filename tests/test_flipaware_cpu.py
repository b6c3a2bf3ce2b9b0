"""The flip-aware oracle (oracle/flipaware.py) on the CPU: injecting the oracle's OWN branch masks changes
nothing; a mask bit flipped at a pre-activation next to its kink passes `check` and changes only what that
element feeds; a bit flipped far from the kink is reported as an error; bookkeeping errors are loud."""
import pytest
import torch

import oracle
from oracle.flipaware import FlipAware
from oracle.modules import TPReLU


def _nets():
    torch.manual_seed(3)
    g = oracle.GeneratorLearnedInputSpace(16, 16, 4, 2, 8, "weight", 2, "fractional").double()
    d = oracle.build_discriminator(16, 16, 4, 2, "weight", 0).double()
    with torch.no_grad():
        for m in list(g.modules()) + list(d.modules()):
            if isinstance(m, TPReLU):
                m.weight.uniform_(-0.2, 1.2)
                m.bias.uniform_(-0.3, 0.3)
    return g, d


def _loss(g, d, z):
    img, lis = g(z, n_execute_lis_layers="all")
    return (d(img) * torch.linspace(-1, 1, z.shape[0], dtype=z.dtype).view(-1, 1)).sum() + sum(u.pow(2).mean() for u in lis)


def _preacts(g, d, z):
    """Pre-activation (the TPReLU's input) of every TPReLU call, in call order, per module."""
    seen, hooks = {}, []
    for net in (g, d):
        for m in net.modules():
            if isinstance(m, TPReLU):
                hooks.append(m.register_forward_hook(lambda mod, inp, out: seen.setdefault(id(mod), []).append(inp[0].detach())))
    with torch.no_grad():      # (not differentiated: an installed injector stays out of the way)
        _loss(g, d, z)
    for h in hooks:
        h.remove()
    return seen


def _grads(g, d, z):
    for p in list(g.parameters()) + list(d.parameters()):
        p.grad = None
    zz = z.clone().requires_grad_(True)
    _loss(g, d, zz).backward()
    return [zz.grad.clone()] + [p.grad.clone() for p in list(g.parameters()) + list(d.parameters())]


def _masks(mod, xs):
    shape = lambda x: (1, -1) + (1,) * (x.dim() - 2)
    return [~((x - mod.bias.detach().view(shape(x))) > 0) for x in xs]


def test_own_masks_change_nothing_and_near_kink_flips_pass():
    g, d = _nets()
    z = torch.randn(5, 8, dtype=torch.float64)
    plain = _grads(g, d, z)
    seen = _preacts(g, d, z)
    mods = {id(m): m for net in (g, d) for m in net.modules() if isinstance(m, TPReLU)}
    fa = FlipAware(g, d)
    for k, xs in seen.items():
        for mask in _masks(mods[k], xs):
            fa.feed(mods[k], mask)
    injected = _grads(g, d, z)
    flips, total, _ = fa.check(1e-4)
    assert flips == 0 and total == sum(x.numel() for xs in seen.values() for x in xs)
    for a, b in zip(plain, injected):      # (same function, another summation order)
        assert (a - b).abs().max().item() <= 1e-12 * max(b.abs().max().item(), 1e-30)

    # move one pre-activation of D's first TPReLU onto its kink (through the TPReLU's translation), flip its bit:
    # accepted, and the gradient changes
    first = [m for m in d.modules() if isinstance(m, TPReLU)][0]
    x0 = seen[id(first)][0]
    with torch.no_grad():
        first.weight[1] = 0.25
        first.bias[1] = x0[1, 1, 3, 3] - 1e-9        # t = +1e-9 for that element: positive side, |t| tiny
    seen = _preacts(g, d, z)
    for k, xs in seen.items():
        for mask in _masks(mods[k], xs):
            if k == id(first):
                assert not mask[1, 1, 3, 3]
                mask[1, 1, 3, 3] = True
            fa.feed(mods[k], mask)
    flipped = _grads(g, d, z)
    flips, _, worst = fa.check(1e-4)
    assert flips == 1 and worst < 1e-6
    fa.remove()
    assert not any("forward" in m.__dict__ for m in mods.values())
    base = _grads(g, d, z)
    assert any((a - b).abs().max().item() > 1e-6 * b.abs().max().item() for a, b in zip(base, flipped))


def test_far_flip_and_bookkeeping_errors_are_loud():
    g, d = _nets()
    z = torch.randn(4, 8, dtype=torch.float64)
    seen = _preacts(g, d, z)
    mods = {id(m): m for net in (g, d) for m in net.modules() if isinstance(m, TPReLU)}
    fa = FlipAware(g, d)
    victim = [m for m in g.modules() if isinstance(m, TPReLU)][-1]
    for k, xs in seen.items():
        for mask in _masks(mods[k], xs):
            if k == id(victim):
                t = (xs[0] - victim.bias.detach().view(1, -1, 1, 1)).abs()
                idx = (t == t.max()).nonzero()[0]
                mask[tuple(idx)] = ~mask[tuple(idx)]
            fa.feed(mods[k], mask)
    _grads(g, d, z)
    with pytest.raises(AssertionError, match="not a rounding flip"):
        fa.check(1e-4)
    # masks left over
    fa.feed(victim, torch.zeros(4, victim.weight.numel(), 8, 8, dtype=torch.bool))
    with pytest.raises(AssertionError, match="never"):
        fa.check(1e-4)
    fa.remove()
    # a differentiated call without a reported mask
    fa = FlipAware(g, d)
    with pytest.raises(AssertionError, match="more differentiated rows"):
        _grads(g, d, z)
    fa.remove()
