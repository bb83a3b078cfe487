"""ctypes binding of the C ABI in include/lcbi_b200.h (liblcbi_b200.so, built in-tree by build.py).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblcbi_b200.so")

_lock = threading.Lock()
_lib = None

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p
c_i32p = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); must list every symbol include/lcbi_b200.h declares
SIGNATURES = {
    "lcbi_version": (ctypes.c_int, []),
    "lcbi_last_error": (ctypes.c_char_p, []),
    "lcbi_dense_attn_fwd": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, c_i64p, c_i64p, c_i64p, c_i64p, ctypes.c_float,
                                           c_vp]),
    "lcbi_dense_attn_fwd_state": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_int, c_i64p, c_i64p, c_i64p, c_i64p,
                                                 ctypes.c_float, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp]),
    "lcbi_dense_attn_bwd_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "lcbi_dense_attn_bwd_workspace_bytes_for": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "lcbi_dense_attn_bwd": (ctypes.c_int, [c_vp] * 9 + [ctypes.c_int] * 5 + [c_i64p] * 8 +
                            [ctypes.c_float, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_size_t, c_vp]),
    "lcbi_attn_merge": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, c_vp]),
    "lcbi_patch_embed_fwd": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, c_vp]),
    "lcbi_patch_embed_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, c_i32p, c_i32p, ctypes.c_int]),
    "lcbi_patch_embed_fwd_ws": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, c_vp, ctypes.c_size_t,
                                               c_vp]),
    "lcbi_patch_embed_bwd": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp,
                                            ctypes.c_int, ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, c_vp]),
    "lcbi_patch_embed_bwd_ws": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp,
                                               ctypes.c_int, ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, c_vp,
                                               ctypes.c_size_t, c_vp]),
    "lcbi_win_attn_fwd": (ctypes.c_int, [ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lcbi_win_attn_bwd": (ctypes.c_int, [ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_float] + [c_vp] * 11),
    "lcbi_win_attn_fwd_range": (ctypes.c_int, [ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int,
                                               ctypes.c_int, c_vp]),
    "lcbi_win_attn_bwd_range": (ctypes.c_int, [ctypes.c_int, c_i32p, c_i32p, c_i32p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_float] + [c_vp] * 10 +
                                              [ctypes.c_int, ctypes.c_int, c_vp]),
    "lcbi_layer_norm_fwd": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_int, c_vp, c_vp,
                                           ctypes.c_int64, ctypes.c_int, ctypes.c_float, c_vp]),
    "lcbi_layer_norm_bwd_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int]),
    "lcbi_layer_norm_bwd": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                           c_vp, ctypes.c_size_t, ctypes.c_int64, ctypes.c_int, c_vp]),
    "lcbi_add_layer_norm_fwd": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, ctypes.c_int,
                                               c_vp, c_vp, ctypes.c_int64, ctypes.c_int, ctypes.c_float, c_vp]),
    "lcbi_add_layer_norm_bwd": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                               ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_size_t, ctypes.c_int64,
                                               ctypes.c_int, c_vp]),
    "lcbi_gather_rows": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int, c_vp]),
    "lcbi_scatter_rows": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int64, ctypes.c_int, c_vp]),
    "lcbi_bias_grad": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp, ctypes.c_size_t, ctypes.c_int64, ctypes.c_int, c_vp]),
    "lcbi_set_window_kernel_mode": (ctypes.c_int, [ctypes.c_int]),
    "lcbi_set_reserved_sms": (ctypes.c_int, [ctypes.c_int]),
    "lcbi_get_reserved_sms": (ctypes.c_int, []),
    "lcbi_window_maps": (ctypes.c_int, [ctypes.c_int, c_i32p, c_i32p, c_i32p, c_vp, c_vp, c_vp, c_i32p, c_i32p, c_vp]),
}


def load():
    """Loads (once) and returns the ctypes handle. Raises RuntimeError if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m long_context_biomedical_imaging_b200.build` "
                "(there is no CPU / PyTorch fallback for the attention hot path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().lcbi_last_error().decode("utf-8", "replace")
        if rc in (-1, -2):
            raise ValueError(f"{what} failed ({rc}): {msg}")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def int_array(vals):
    vals = [int(v) for v in vals]
    return (ctypes.c_int32 * len(vals))(*vals)


def int3(vals):
    return (ctypes.c_int32 * 3)(*[int(v) for v in vals])


def strides3(*vals):
    return (ctypes.c_int64 * 3)(*[int(v) for v in vals])
