"""ctypes binding of ``libglis_b200.so`` (C ABI declared in ``include/glis_b200.h``).

There is no fallback: if the shared library is missing or a kernel reports an error the
call raises.  Pointers are borrowed from torch tensors for the duration of a call and the
work is enqueued on torch's current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglis_b200.so")

CONV, TCONV = 0, 1
ACT_NONE, ACT_TPRELU, ACT_SIGMOID = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2


class Geom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "relation", "N", "Hi", "Wi", "Ci", "Ho", "Wo", "Co",
        "KH", "KW", "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w")]


class Epilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int32), ("act_a", C.c_void_p),
                ("act_b", C.c_void_p), ("preact", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
                ("act_channels", C.c_int32), ("split_slabs", C.c_int32)]


class WnLayer(C.Structure):
    """glis_wn_layer_t: one layer of a glis_wn_prepare_multi call."""
    _fields_ = ([(n, C.c_void_p) for n in ("w", "scale", "norm", "pack_io", "pack_oi", "fwd_hi", "fwd_lo", "bwd_hi",
                                            "bwd_lo", "mat_hi", "mat_lo", "matt_hi", "matt_lo")]
                + [(n, C.c_int32) for n in ("out_axis", "Cout", "Cin", "T", "perm_c", "perm_p", "mat_rows", "need_norm")]
                + [("c", C.c_float), ("reserved", C.c_int32)])


class WnProj(C.Structure):
    """glis_wn_proj_t: one layer of a glis_wn_project_multi call."""
    _fields_ = ([(n, C.c_void_p) for n in ("G", "w", "scale", "norm", "dw", "dscale")]
                + [(n, C.c_int32) for n in ("out_axis", "Cout", "Cin", "T", "accumulate")] + [("c", C.c_float)])


_vp, _i, _f, _i64, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint64

# name -> argtypes; every symbol include/glis_b200.h declares
SIGNATURES = {
    "glis_wn_prepare": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp],
    "glis_wn_project": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp],
    "glis_wn_prepare_multi": [C.POINTER(WnLayer), _i, _vp],
    "glis_wn_project_multi": [C.POINTER(WnProj), _i, _vp],
    "glis_conv_forward": [C.POINTER(Geom), _vp, _vp, C.POINTER(Epilogue), _vp, _i, _vp],
    "glis_conv_wgrad": [C.POINTER(Geom), _vp, _vp, _vp, _i, _vp],
    "glis_split_bf16": [_vp, _vp, _vp, _i64, _vp],
    "glis_wn_prepare_bf16": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
    "glis_conv_tc_supported": [C.POINTER(Geom)],
    "glis_conv_tc_ksplit": [C.POINTER(Geom)],
    "glis_conv_tc_plan": [C.POINTER(Geom), _i, C.POINTER(C.c_int)],
    "glis_conv_tc_halo_plan": [C.POINTER(Geom), _i, C.POINTER(C.c_int)],
    "glis_conv_tc_pair_plan": [C.POINTER(Geom), _i, C.POINTER(C.c_int)],
    "glis_tprelu_forward_planes_sum": [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "glis_lis_supported": [_i],
    "glis_lis_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "glis_lis_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "glis_sigmoid_backward": [_vp, _vp, _vp, _i64, _vp],
    "glis_tprelu_forward_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "glis_conv_forward_bf16": [C.POINTER(Geom), _vp, _vp, _vp, _vp, C.POINTER(Epilogue), _vp, _vp, _vp, _i, _vp],
    "glis_wgrad_tc_supported": [C.POINTER(Geom)],
    "glis_conv_wgrad_bf16": [C.POINTER(Geom), _vp, _vp, _vp, _vp, _vp, _i, _vp],
    "glis_wgrad_tc_splits": [C.POINTER(Geom)],
    "glis_slab_reduce": [_vp, _i, _i64, _vp, _i64, _vp],
    "glis_conv_wgrad_bf16_slabs": [C.POINTER(Geom), _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "glis_wn_project_slabs": [_vp, _i, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _vp],
    "glis_tprelu_forward": [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "glis_tprelu_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "glis_tprelu_backward_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "glis_channel_sum": [_vp, _vp, _i64, _i, _i, _i, _vp],
    "glis_bce_logits": [_vp, _f, _i, _f, _vp, _vp, _vp, _vp],
    "glis_mse_scaled": [_vp, _vp, _i64, _f, _vp, _vp, _i, _vp],
    "glis_lsq_logits": [_vp, _f, _i, _f, _vp, _vp, _vp, _vp],
    "glis_dropout": [_vp, _vp, _i64, _i, _i, _i64, _i, _f, _u64, _vp, _u64, _vp],
    "glis_counter_add": [_vp, _u64, _vp],
    "glis_augment": [_vp, _vp, _vp, _i, _i, _i, _i, _u64, _vp],
    "glis_rmsprop": [_vp, _vp, _vp, _i64, _f, _f, _f, _f, _vp],
    "glis_randn": [_vp, _i64, _u64, _u64, _vp],
    "glis_uniform": [_vp, _i64, _u64, _u64, _vp],
    "glis_wn_prepare_perm": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp],
    "glis_wn_prepare_bf16_perm": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "glis_linear_wgrad": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "glis_linear_wgrad_project_supported": [_i, _i, _i],
    "glis_linear_wgrad_project": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "glis_unfold4x4s2_bf16": [_vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "glis_fold4x4s2": [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "glis_wn_pack_matrix_bf16": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "glis_peer_allreduce": [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i, _i, _i64, _i64, _vp, _i, _vp],
}

_lib = None


def load():
    """Load the library once; raises ``RuntimeError`` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "glis_b200: %s is missing — build it with `python __graft_entry__.py` "
            "(or `make -C gan-error-avoidance_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.glis_last_error.restype = C.c_char_p
    lib.glis_last_error.argtypes = []
    lib.glis_version.restype = C.c_int
    lib.glis_version.argtypes = []
    lib.glis_set_pdl.restype = C.c_int
    lib.glis_set_pdl.argtypes = [C.c_int]
    lib.glis_set_reserved_sms.restype = C.c_int
    lib.glis_set_reserved_sms.argtypes = [C.c_int]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, load().glis_last_error().decode()))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a dense fp32 CUDA tensor (or NULL for None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("glis_b200 kernels need CUDA tensors; there is no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError("glis_b200 kernels are fp32 at the API boundary, got %s" % t.dtype)
    return C.c_void_p(t.data_ptr())


# kernels enqueued by one successful call of each entry point (memsets are not kernels)
KERNELS_PER_CALL = {name: 1 for name in SIGNATURES}

launch_count = 0     # kernels this process enqueued through the C ABI


def ptr16(t):
    """Device pointer of a dense bf16 CUDA tensor (or NULL for None)."""
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise RuntimeError("glis_b200: expected a CUDA bfloat16 plane")
    return C.c_void_p(t.data_ptr())


PRECISION_NAMES = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}
default_precision = PRECISION_NAMES[os.environ.get("GLIS_PRECISION", "bf16x3")]


def set_precision(name):
    """Contraction arithmetic of the layers that tile for tcgen05: 'fp32' (FFMA everywhere),
    'bf16x3' (split-bf16, fp32-faithful; default) or 'bf16' (single pass)."""
    global default_precision
    default_precision = PRECISION_NAMES[name]


def call(name, *args, kernels=None):
    global launch_count
    check(getattr(load(), name)(*args), name)
    launch_count += KERNELS_PER_CALL[name] if kernels is None else kernels


launch_tags = {}     # tag -> launches (GEMM-shaped launches tag themselves with their M / N / K: bench.py's tally)


class timed(object):
    """``with timed(tag):`` tallies the enclosed launch under ``tag`` (kernel family + GEMM shape).  Counting only:
    per-kernel TIMES come from the ncu launch lists under profiles/ — CUDA events recorded around an asynchronous
    launch bracket its enqueue, not its execution."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        launch_tags[self.tag] = launch_tags.get(self.tag, 0) + 1
        return self

    def __exit__(self, *exc):
        return False
