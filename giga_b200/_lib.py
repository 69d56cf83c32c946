"""ctypes binding of the C ABI declared in include/giga_b200.h (libgiga_b200.so).

There is deliberately no fallback: if the CUDA library is missing or cannot be
loaded the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libgiga_b200.so")

HEAD_QUAL, HEAD_ROT, HEAD_WIDTH, HEAD_TSDF, HEAD_RAW = 1, 2, 4, 8, 16
HEAD_GRASP = HEAD_QUAL | HEAD_ROT | HEAD_WIDTH

# symbol -> (restype, argtypes); must list every function include/giga_b200.h declares
_F = C.POINTER(C.c_float)
SYMBOLS = {
    "giga_version": (C.c_int, []),
    "giga_last_error": (C.c_char_p, []),
    "giga_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "giga_ctx_destroy": (None, [C.c_void_p]),
    "giga_ctx_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long, C.c_int]),
    "giga_ctx_set_params_flat": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_long), C.POINTER(C.c_long),
                                          C.c_void_p, C.c_long, C.c_int]),
    "giga_ctx_commit_params": (C.c_int, [C.c_void_p]),
    "giga_ctx_heads": (C.c_uint, [C.c_void_p]),
    "giga_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "giga_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_sample_feature": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "giga_scene_argmax": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_forward_host_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_forward_host_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "giga_select_params_default": (None, [C.c_void_p]),
    "giga_gaussian_kernel1d": (C.c_int, [C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "giga_select_grasps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_ctx_set_lattice": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "giga_detect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_detect_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_vgn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_loss": (C.c_int, [C.c_void_p] + [C.c_void_p] * 8 + [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_void_p]),
    "giga_adam_step": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_long, C.c_int] + [C.c_double] * 5 + [C.c_void_p]),
    "giga_ctx_commit_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_debug_blob": (C.c_long, [C.c_void_p, C.c_int, C.c_void_p, C.c_long]),
    "giga_train_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "giga_tsdf_integrate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_void_p]),
    "giga_tsdf_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "giga_mise_sweep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]),
    "giga_ctx_launch_count": (C.c_long, [C.c_void_p]),
    "giga_ctx_overflow_count": (C.c_long, [C.c_void_p, C.c_int]),
    "giga_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "giga_ctx_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "giga_ctx_timing_report": (C.c_long, [C.c_void_p, C.c_char_p, C.c_long]),
    "giga_debug_copy": (C.c_long, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long, C.c_void_p]),
}


class SelectParams(C.Structure):
    """giga_select_params (include/giga_b200.h)"""
    _fields_ = [("gaussian_sigma", C.c_double), ("min_width", C.c_float), ("max_width", C.c_float), ("out_th", C.c_float),
                ("lim_x", C.c_int), ("lim_y", C.c_int), ("lim_z", C.c_int), ("low_th", C.c_float), ("threshold", C.c_float),
                ("force_detection", C.c_int), ("max_filter_size", C.c_int)]


class GigaError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the giga_b200 CUDA library is not built. Run "
            "`python giga_b200/build.py` (or __graft_entry__.build()). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        raise GigaError(f"{what}: {lib.giga_last_error().decode()} (code {rc})")
    return rc
