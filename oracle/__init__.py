"""ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement (modern PyTorch, fp32 / fp64) of the reference's G-LIS hot path:
the weight-normalized modules, the G / D / R / LIS builders and one adversarial
training iteration.  Every function cites the reference file:line it follows
(paths relative to the upstream repository aleju/gan-error-avoidance).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or as
the timed CPU baseline.  Nothing under ``gan-error-avoidance_b200/`` imports it.

Parity pin: the reference ships no golden vectors or tests and
cannot run unmodified under Python 3 / torch 2.x.  The oracle is pinned instead
against the reference's OWN module sources executed in-process under a small
legacy-semantics shim (``tests/golden/make_golden.py``: keep-dim ``sum``, legacy
``_ConvNd`` constructor, dotted module names, tab normalisation).  The vectors
that run produced are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.  Third-party arithmetic (conv / GEMM / PReLU /
BCE / RMSprop) lives in PyTorch (reference pin: commit 065c5986, README.md:106,
not installable offline); torch 2.11.0 CPU is the nearest installable stand-in.
"""
from .modules import (TPReLU, View, WeightNormalizedConv2d,
                      WeightNormalizedConvTranspose2d, WeightNormalizedLinear)
from .model import (GeneratorLearnedInputSpace, build_discriminator,
                    build_generator, build_reverser)
from .step import GLISOracleTrainer, glis_iteration, riter_iteration, rmsprop_update, rsep_iteration

__all__ = [
    "TPReLU", "View", "WeightNormalizedConv2d", "WeightNormalizedConvTranspose2d",
    "WeightNormalizedLinear", "GeneratorLearnedInputSpace", "build_discriminator",
    "build_generator", "build_reverser", "GLISOracleTrainer", "glis_iteration",
    "rmsprop_update", "riter_iteration", "rsep_iteration",
]
