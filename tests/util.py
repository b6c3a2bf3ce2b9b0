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


def _philox4x32_10(seed, quad, stream):
    """Philox4x32-10 (Salmon et al., SC'11), vectorised over ``quad``: key = seed (64 bit), counter =
    {quad lo, quad hi, stream lo, stream hi} — the host restatement of csrc/pointwise.cu::philox4_stream."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    quad = np.asarray(quad, dtype=np.uint64)
    c = [quad & mask, quad >> np.uint64(32), np.full_like(quad, np.uint64(stream & 0xFFFFFFFF)),
         np.full_like(quad, np.uint64((stream >> 32) & 0xFFFFFFFF))]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def philox_keep_mask(seed, stream, numel, p):
    """Keep mask (bool tensor of ``numel`` elements) of one glis_dropout draw: element e is kept iff
    u01(Philox(seed, {e // 4, stream})[e % 4]) >= p with u01(r) = (r >> 8) / 2^24 compared in fp32."""
    quads = np.arange((numel + 3) // 4, dtype=np.uint64)
    words = np.stack(_philox4x32_10(int(seed), quads, int(stream)), axis=1).reshape(-1)[:numel]
    u = (words >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return torch.from_numpy(u >= np.float32(p))
