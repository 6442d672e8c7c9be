"""ctypes access to the CPU oracle (oracle/liboracle.so) and to the reference's own object code
(oracle/_ref/libwalkgen_ref.so).  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline leg import this module."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = C.POINTER(C.c_double)
I64 = C.POINTER(C.c_int64)
_ora = None
_ref = None


def dptr(a):
    return None if a is None else a.ctypes.data_as(D)


def oracle():
    global _ora
    if _ora is None:
        _ora = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _ora.oracle_preview_gains.restype = C.c_int
        _ora.oracle_preview_gains.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, D, D, D, C.c_int, D, D, D]
        _ora.oracle_preview_run.restype = C.c_int
        _ora.oracle_preview_run.argtypes = [D, D, D, D, C.c_double, D, C.c_int, D, C.c_int, D, D, D, C.c_int]
        _ora.oracle_preview_run_batch.restype = C.c_long
        _ora.oracle_preview_run_batch.argtypes = [D, D, D, D, C.c_double, D, C.c_int, C.c_int, I64, D, D, D, D, C.c_int]
        _ora.oracle_preview_run_batch_mt.restype = C.c_long
        _ora.oracle_preview_run_batch_mt.argtypes = _ora.oracle_preview_run_batch.argtypes + [C.c_int]
        _ora.oracle_preview_step.restype = C.c_int
        _ora.oracle_preview_step.argtypes = [D, D, D, D, C.c_double, D, C.c_int, D, D, D, D, D, C.c_int, D, D, C.c_int]
        _ora.oracle_preview_step_1d_wrap.restype = C.c_int
        _ora.oracle_preview_step_1d_wrap.argtypes = [D, D, D, D, C.c_double, D, C.c_int, D, D, D, C.c_int, C.c_int, D, C.c_int]
    return _ora


def ref():
    """The reference's own qld/OptCholesky/PLDP object code; None if it was never built."""
    global _ref
    p = os.path.join(ROOT, "oracle", "_ref", "libwalkgen_ref.so")
    if _ref is None and os.path.exists(p):
        _ref = C.CDLL(p)
    return _ref


class OracleGains:
    def __init__(self, T=0.005, preview_time=1.6, zc=0.814, mode=1):
        self.A = np.zeros(9); self.B = np.zeros(3); self.C = np.zeros(3); self.Kx = np.zeros(3)
        F = np.zeros(4096)
        ks = C.c_double()
        nl = oracle().oracle_preview_gains(T, preview_time, zc, mode, C.byref(ks), dptr(self.Kx), dptr(F), 4096,
                                           dptr(self.A), dptr(self.B), dptr(self.C))
        if nl < 0:
            raise RuntimeError(f"oracle_preview_gains failed: {nl}")
        self.NL = nl
        self.F = F[:nl].copy()
        self.Ks = ks.value


def oracle_preview_batch(g, offsets, zmpref_xy, state, simulation=True, want_out=True, threads=0, out=None):
    """Run the oracle over a ragged batch; returns (com, zmp) with the product's row indexing.
    threads > 0 selects the reference-layout (deque of ZMPPosition) multi-threaded CPU-baseline driver."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = int(offsets[-1])
    if out is not None:
        com, zmp = out
    else:
        com = np.zeros((n, 6)) if want_out else None
        zmp = np.zeros((n, 2)) if want_out else None
    args = [dptr(g.A), dptr(g.B), dptr(g.C), dptr(g.Kx), g.Ks, dptr(g.F), g.NL, len(offsets) - 1,
            offsets.ctypes.data_as(I64), dptr(zmpref_xy), dptr(state), dptr(com), dptr(zmp), int(simulation)]
    if threads > 0:
        steps = oracle().oracle_preview_run_batch_mt(*args, int(threads))
    else:
        steps = oracle().oracle_preview_run_batch(*args)
    return com, zmp, steps
