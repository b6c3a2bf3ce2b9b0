"""Shared helpers of the parity tests."""
import numpy as np
import torch


def rel_err(got, want):
    """max |got - want| / max |want|  — the relative error every tolerance in tests/ refers to."""
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    denom = want.abs().max().item()
    return (got - want).abs().max().item() / (denom if denom > 0 else 1.0)


def rel_l2(got, want):
    got = got.detach().double().cpu().reshape(-1)
    want = want.detach().double().cpu().reshape(-1)
    d = want.norm().item()
    return (got - want).norm().item() / (d if d > 0 else 1.0)


def copy_params(dst, src):
    """Copy parameters between two modules with identical state_dict keys (oracle <-> product)."""
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    dst.load_state_dict({k: v.to(next(dst.parameters()).device, dtype=next(dst.parameters()).dtype)
                         for k, v in sd.items()})


def randomize_params_(module, gen, tprelu_range=(-0.2, 1.2)):
    with torch.no_grad():
        for name, p in module.named_parameters():
            leaf = name.split(".")[-1]
            r = torch.rand(p.shape, generator=gen, dtype=torch.float32)
            if leaf == "weight" and p.dim() == 1:
                p.copy_(r * (tprelu_range[1] - tprelu_range[0]) + tprelu_range[0])
            elif leaf == "bias" and p.dim() == 1:
                p.copy_(r * 0.6 - 0.3)
            elif leaf == "scale":
                p.copy_(r + 0.5)
            elif leaf == "bias":
                p.copy_(r * 0.4 - 0.2)
            else:
                p.copy_(p * (0.5 + r))


class FlipAwarePair(object):
    """Couples the product's TPReLU pre-activation tap (``ops.PreactTap``) with the oracle's mask injection
    (``oracle.flipaware``): run the product inside ``with pair.tap():``, call ``feed()``, run the oracle, call
    ``check()``.

    ``pairs``: [(oracle network, product network), ...] with identical parameter order.  The gradient
    comparison that follows may then use the per-op tolerance through whole chains: both sides differentiate
    the same piecewise-linear function, and ``check`` has verified that their masks differ only where the
    oracle's pre-activation lies within ``tol`` (relative to the tensor's max) of the TPReLU kink."""

    def __init__(self, pairs):
        from oracle.flipaware import FlipAware
        from oracle.modules import TPReLU as OracleTPReLU
        self.fa = FlipAware(*[o for o, _ in pairs])
        self.owner = []          # (product slope Parameter, oracle TPReLU module)
        for onet, pnet in pairs:
            slope_owner = {id(m.weight): m for m in onet.modules() if isinstance(m, OracleTPReLU)}
            for po, pp in zip(onet.parameters(), pnet.parameters()):
                if id(po) in slope_owner:
                    self.owner.append((pp, slope_owner[id(po)]))
        self._tap = None

    def tap(self):
        from glis_b200 import ops
        self._tap = ops.PreactTap()
        return self._tap

    def feed(self):
        """Hand the branch masks of everything recorded since ``tap()`` to the oracle's modules."""
        by_ptr = {pp.data_ptr(): m for pp, m in self.owner}     # (parameters may have been re-homed since)
        for a_raw, neg in self._tap.records:
            self.fa.feed(by_ptr[a_raw.data_ptr()], neg)
        self._tap.records = []

    def check(self, tol):
        return self.fa.check(tol)

    def remove(self):
        self.fa.remove()
