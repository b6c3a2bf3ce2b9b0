"""glis_b200 — host side of the B200-native G-LIS training step.

``_lib``   ctypes binding of the C ABI (``include/glis_b200.h``)
``ops``    autograd operators over the sm_100a kernels
``naming`` dotted child names / reference-compatible ``state_dict`` keys
"""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
