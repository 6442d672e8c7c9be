"""ctypes binding of the C ABI declared in include/walkgen_b200.h.

This module is *plumbing*: it loads ``libwalkgen_b200.so`` (built in-tree by
``jrl_walkgen_b200/csrc/Makefile``) and declares argument types.  It never falls back to a
CPU implementation: if the library is missing, or no CUDA device is present when a context
is requested, it raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwalkgen_b200.so")

WG_OK = 0
WG_ERR_NO_DEVICE = -1
WG_ERR_CUDA = -2
WG_ERR_INVALID = -3
WG_ERR_ALLOC = -4
WG_ERR_NOT_READY = -5
WG_ERR_WINDOW = -6
WG_MEM_HOST = 0
WG_MEM_DEVICE = 1
WG_PREVIEW_MAX_NL = 2048
MODE_WITH_INITIALPOS = 0
MODE_WITHOUT_INITIALPOS = 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)


class WalkgenError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"walkgen_b200 error {code}: {msg}")
        self.code = code


class PreviewGains(C.Structure):
    """Mirror of wg_preview_gains_t."""
    _fields_ = [
        ("A", C.c_double * 9),
        ("B", C.c_double * 3),
        ("C", C.c_double * 3),
        ("Kx", C.c_double * 3),
        ("Ks", C.c_double),
        ("T", C.c_double),
        ("preview_time", C.c_double),
        ("zc", C.c_double),
        ("mode", C.c_int),
        ("NL", C.c_int),
        ("F", C.c_double * WG_PREVIEW_MAX_NL),
    ]


class HerdtParams(C.Structure):
    """Mirror of wg_herdt_params."""
    _fields_ = [("T", C.c_double), ("com_height", C.c_double), ("w_jerk", C.c_double), ("w_vel", C.c_double),
                ("w_cop", C.c_double), ("cop_half_x", C.c_double), ("cop_half_y", C.c_double),
                ("ds_feet_distance", C.c_double), ("foot_hull_x", C.c_double * 5), ("foot_hull_y", C.c_double * 5),
                ("lipm_T", C.c_double)]


HERDT_N = 16
HERDT_MAX_VARS = 36
HERDT_MAX_ROWS = 75


def _np():
    import numpy as np
    return np


def herdt_dtypes():
    """numpy mirrors of wg_herdt_qp_input (784 B) and wg_herdt_qp_output (960 B)."""
    np = _np()
    N = HERDT_N
    qin = np.dtype([
        ("com_x", "f8", 3), ("com_y", "f8", 3), ("ref_x", "f8", N), ("ref_y", "f8", N),
        ("sup_x", "f8", N + 1), ("sup_y", "f8", N + 1), ("sup_yaw", "f8", N + 1),
        ("sup_foot", "i1", N + 1), ("sup_phase", "i1", N + 1), ("sup_step", "i1", N + 1),
        ("sup_changed", "i1", N + 1), ("pad_", "i1", 4)])
    qout = np.dtype([
        ("x", "f8", HERDT_MAX_VARS), ("lagr", "f8", HERDT_MAX_ROWS + 1), ("com_next_x", "f8", 3),
        ("com_next_y", "f8", 3), ("n_vars", "i4"), ("n_rows", "i4"), ("fail", "i4"), ("iterations", "i4")])
    assert qin.itemsize == 784 and qout.itemsize == 960
    return qin, qout


# name -> (restype, argtypes); also the list checked against include/walkgen_b200.h by the tests
SIGNATURES = {
    "wg_version": (C.c_int, []),
    "wg_device_count": (C.c_int, []),
    "wg_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "wg_ctx_destroy": (C.c_int, [C.c_void_p]),
    "wg_sync": (C.c_int, [C.c_void_p]),
    "wg_last_error": (C.c_char_p, [C.c_void_p]),
    "wg_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "wg_malloc_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "wg_free_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "wg_malloc_pinned": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "wg_free_pinned": (C.c_int, [C.c_void_p, C.c_void_p]),
    "wg_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "wg_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "wg_memset_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]),
    "wg_timer_start": (C.c_int, [C.c_void_p]),
    "wg_timer_stop_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "wg_launch_count": (C.c_longlong, [C.c_void_p]),
    "wg_launch_count_reset": (None, [C.c_void_p]),
    "wg_prof_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "wg_prof_end": (C.c_int, [C.c_void_p]),
    "wg_prof_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), c_double_p]),
    "wg_measure_fp64_peak": (C.c_int, [C.c_void_p, c_double_p]),
    "wg_preview_gains": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(PreviewGains)]),
    "wg_preview_set_gains": (C.c_int, [C.c_void_p, C.POINTER(PreviewGains)]),
    "wg_preview_plan_create": (C.c_int, [C.c_void_p, C.c_int, c_i64_p, C.POINTER(C.c_void_p)]),
    "wg_preview_plan_destroy": (C.c_int, [C.c_void_p]),
    "wg_preview_plan_total_steps": (C.c_int64, [C.c_void_p]),
    "wg_preview_plan_total_samples": (C.c_int64, [C.c_void_p]),
    "wg_preview_run_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int]),
    "wg_preview_one_iteration": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                           c_double_p, C.c_int, c_double_p, c_double_p, C.c_int]),
    "wg_herdt_default_params": (None, [C.c_double, C.c_double, C.POINTER(HerdtParams)]),
    "wg_herdt_set_params": (C.c_int, [C.c_void_p, C.POINTER(HerdtParams)]),
    "wg_herdt_qp_solve_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the product library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WalkgenError(
            WG_ERR_NO_DEVICE,
            f"{LIB_PATH} not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the ABI header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
