"""ctypes binding of ``libvmtl_b200.so`` (the C ABI declared in ``include/vmtl_b200.h``).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvmtl_b200.so")
_lib = None

# name -> (restype, argtypes); mirrors include/vmtl_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "vmtl_version": (c_int, []),
    "vmtl_strerror": (c_char_p, [c_int]),
    "vmtl_sm_count": (c_int, []),
    "vmtl_xstitch_fwd": (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P]),
    "vmtl_xstitch_bwd_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int]),
    "vmtl_xstitch_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P, c_size_t, _P]),
    "vmtl_xstitch_cat_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, _P]),
    "vmtl_xstitch_cat_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                        c_int]),
    "vmtl_xstitch_cat_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "vmtl_gate_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int]),
    "vmtl_gate_tc_supported": (c_int, [c_int, c_int]),
    "vmtl_gate_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_float, c_float, c_int, c_int, c_int64,
                              c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_gate_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int64, c_int, c_int,
                              _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_bnrelu_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "vmtl_bnrelu_fwd": (c_int, [_P, _P, _P, _P, _P, c_float, c_float, c_int, c_int, c_int64, c_int, _P, _P, _P, _P,
                                _P, _P, c_size_t, _P]),
    "vmtl_bnrelu_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_bnrelu_pool_fwd": (c_int, [_P, _P, _P, _P, _P, c_float, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                     _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_bnrelu_pool_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P,
                                     c_size_t, _P]),
    "vmtl_loss_workspace_bytes": (c_size_t, [c_int64]),
    "vmtl_head_ce_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int64, _P, _P, _P, _P, _P,
                                 c_size_t, _P]),
    "vmtl_head_ce_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, c_int64, _P, _P, _P, _P, _P, _P,
                                 c_size_t, _P]),
    "vmtl_ce_logits_fwd": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, c_int64, _P, _P, _P, _P, _P,
                                   c_size_t, _P]),
    "vmtl_ce_logits_bwd": (c_int, [_P, _P, c_int64, c_int64, c_int, c_int, c_int64, _P, _P, _P, _P]),
    "vmtl_head_silog_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_head_silog_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_float, _P, _P, _P, _P, _P, _P,
                                    c_size_t, _P]),
    "vmtl_confusion_accum": (c_int, [_P, c_int, _P, c_int64, c_int, c_int64, _P, _P]),
    "vmtl_depth_err_sums": (c_int, [_P, _P, c_int64, c_float, _P, _P, c_size_t, _P]),
    "vmtl_seg_metrics": (c_int, [_P, c_int, _P, _P]),
    # global-batch statistics: the two halves of every op with a batch reduction in its middle
    "vmtl_bn_moments": (c_int, [_P, c_int64, c_int, _P, _P, c_size_t, _P]),
    "vmtl_bnrelu_fwd_global": (c_int, [_P, _P, _P, _P, _P, c_float, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                       _P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_bnrelu_bwd_moments": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t,
                                        _P]),
    "vmtl_bnrelu_bwd_global": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P,
                                       _P, c_size_t, _P]),
    "vmtl_gate_fwd_moments": (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "vmtl_gate_fwd_global": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_float, c_int, c_int64, c_int, c_int, _P,
                                     c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_gate_bwd_moments": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P, _P,
                                      _P, c_size_t, _P]),
    "vmtl_gate_bwd_global": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P,
                                     c_int64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "vmtl_silog_finalize": (c_int, [_P, _P, _P]),
    "vmtl_up2_bilinear_fwd": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int64, _P]),
    "vmtl_up2_bilinear_bwd": (c_int, [_P, c_int64, _P, c_int, c_int, c_int, c_int, _P]),
    "vmtl_head_supported": (c_int, [c_int, c_int, c_int]),
    "vmtl_adam_max_tensors_per_launch": (c_int, []),
    "vmtl_adam_step": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, c_double, _P, _P, c_double, c_double, c_double,
                               c_double, _P]),
}


class VmtlError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library once and type every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise VmtlError(
            f"{_LIB_PATH} is missing: build it with `python -m vision_mtl_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)"
        )
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vmtl_strerror(rc).decode()
        raise VmtlError(f"{what} failed: {msg} (code {rc})")


__all__ = ["load", "check", "lib_path", "VmtlError", "SIGNATURES", "c_double"]
