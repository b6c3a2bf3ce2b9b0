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
