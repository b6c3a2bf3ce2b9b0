"""ORACLE (test infrastructure) — TPReLU backward with an injected branch mask.

A TPReLU (``common/modules/TPReLU.py:16-18``) is piecewise linear: its gradient depends on
which side of the kink ``t = x - b = 0`` every pre-activation lies.  Two correct
implementations whose pre-activations agree to the forward tolerance can still disagree on
that side for the handful of elements with ``|t|`` below the rounding distance, and one such
element moves every gradient downstream of it by far more than rounding.  Comparing gradients
at the north star's 1e-3 therefore needs the SAME masks on both sides:

* the implementation under test reports the pre-activation of every TPReLU it differentiates;
* the oracle's forward stays the reference's (its own mask), its backward uses the reported
  mask (``mask = !(t > 0)``);
* ``check()`` then asserts that the two masks differ only where the oracle's own ``|t|`` lies
  within ``tol * max|t|`` of the kink — a flip anywhere else is a real error.

Nothing here changes the pinned oracle modules: ``install`` overrides ``forward`` on module
INSTANCES, ``remove`` restores them.
"""
import torch

from .modules import TPReLU, tprelu


class _TPReLUInjected(torch.autograd.Function):
    """Reference forward; backward of SURVEY.md App. E with ``neg`` as the branch mask."""

    @staticmethod
    def forward(ctx, x, a_raw, b, neg):
        shape = (1, -1) + (1,) * (x.dim() - 2)
        t = x - b.view(shape)
        ctx.save_for_backward(t, a_raw, neg)
        return tprelu(x, a_raw, b)

    @staticmethod
    def backward(ctx, dout):
        t, a_raw, neg = ctx.saved_tensors
        shape = (1, -1) + (1,) * (t.dim() - 2)
        a = a_raw.clamp(0, 1).view(shape)
        dims = [d for d in range(t.dim()) if d != 1]
        dx = torch.where(neg, a * dout, dout)
        gate = ((a_raw >= 0) & (a_raw <= 1)).to(dout.dtype)
        m = neg.to(dout.dtype)
        da = (dout * t * m).sum(dims) * gate
        db = (dout * m * (1 - a)).sum(dims)
        return dx, da, db, None


class _Injector(object):
    """Per-module state: mask rows waiting to be consumed, and what every differentiated call saw."""

    def __init__(self, module, name):
        self.module, self.name = module, name
        self.rows = []          # bool tensors (batch-major); consumed from the front, `pos` rows used of rows[0]
        self.pos = 0
        self.calls = []         # (own t, injected mask) per differentiated call

    def pending(self):
        return sum(r.shape[0] for r in self.rows) - self.pos

    def take(self, n):
        got = []
        while n > 0:
            if not self.rows:
                raise AssertionError("flip-aware oracle: TPReLU %s ran more differentiated rows than the "
                                     "implementation under test reported" % self.name)
            r = self.rows[0]
            k = min(n, r.shape[0] - self.pos)
            got.append(r[self.pos:self.pos + k])
            self.pos += k
            n -= k
            if self.pos == r.shape[0]:
                self.rows.pop(0)
                self.pos = 0
        return got[0] if len(got) == 1 else torch.cat(got, 0)

    def forward(self, x):
        m = self.module
        differentiated = torch.is_grad_enabled() and (x.requires_grad or m.weight.requires_grad
                                                      or m.bias.requires_grad)
        if not differentiated:
            return tprelu(x, m.weight, m.bias)
        neg = self.take(x.shape[0])
        if tuple(neg.shape) != tuple(x.shape):
            raise AssertionError("flip-aware oracle: TPReLU %s got a mask of shape %s for an input of shape %s"
                                 % (self.name, tuple(neg.shape), tuple(x.shape)))
        shape = (1, -1) + (1,) * (x.dim() - 2)
        self.calls.append(((x.detach() - m.bias.detach().view(shape)), neg))
        return _TPReLUInjected.apply(x, m.weight, m.bias, neg)


class FlipAware(object):
    """Mask injection for every TPReLU of the given oracle networks.

    ``feed(module, mask)`` queues the branch mask (``True`` = negative side) of the next
    ``mask.shape[0]`` batch rows that ``module`` differentiates; calls are served in order and a
    mask may span several calls (D on real + fake as one 2B batch against two oracle calls).
    """

    def __init__(self, *nets):
        self.inj = {}
        for net in nets:
            for name, m in net.named_modules():
                if isinstance(m, TPReLU) and id(m) not in self.inj:
                    j = _Injector(m, name)
                    self.inj[id(m)] = j
                    m.forward = j.forward          # instance attribute shadows the class method

    def remove(self):
        for j in self.inj.values():
            if "forward" in j.module.__dict__:
                del j.module.__dict__["forward"]

    def feed(self, module, neg_mask):
        self.inj[id(module)].rows.append(neg_mask.to(torch.bool).cpu())

    def check(self, tol):
        """Every queued mask consumed; masks differ from the oracle's own only within ``tol * max|t|`` of
        the kink.  Returns (number of flipped elements, number of elements, worst |t| / max|t| among flips)
        and forgets the recorded calls."""
        flips = total = 0
        worst = 0.0
        for j in self.inj.values():
            if j.pending():
                raise AssertionError("flip-aware oracle: %d reported mask rows of TPReLU %s were never "
                                     "differentiated by the oracle" % (j.pending(), j.name))
            for t, neg in j.calls:
                own = ~(t > 0)
                diff = own != neg
                n = int(diff.sum())
                total += t.numel()
                if n:
                    flips += n
                    scale = t.abs().max().item()
                    w = t[diff].abs().max().item() / (scale if scale > 0 else 1.0)
                    worst = max(worst, w)
                    if w > tol:
                        raise AssertionError("TPReLU %s: %d mask bits differ from the oracle's, one of them at "
                                             "|t| = %.3g of max|t| (allowed: %.3g) — not a rounding flip"
                                             % (j.name, n, w, tol))
            j.calls = []
        return flips, total, worst
