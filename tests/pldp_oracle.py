"""ctypes access to the PLDP / OptCholesky oracle (oracle/oracle_pldp.cpp) and to the reference's own object code
(oracle/_ref/libwalkgen_ref.so: PLDPSolver.cpp, OptCholesky.cpp compiled where they lie).  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

import oracle_lib as ol

STATE = np.dtype([("prev_zmp", "f8", 32), ("prev_active", "i4", 32), ("n_prev", "i4"), ("pad_", "i4")])
_sig = False


def lib():
    global _sig
    L = ol.oracle()
    if not _sig:
        L.oracle_pldp_solve.restype = C.c_int
        L.oracle_pldp_solve.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_pldp_solve_sim.restype = C.c_int
        L.oracle_pldp_solve_sim.argtypes = L.oracle_pldp_solve.argtypes + [C.c_void_p]
        L.oracle_optcholesky_add_rows.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_optcholesky_full.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_optcholesky_inverse.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _sig = True
    return L


def oracle_solve(K, pb, b, hot=None, hot_start=True, starting=True, n_removed=0, max_iter=128, similar=None):
    """One SolveProblem by the oracle port.  -> (X[32], info[4], active[32])."""
    sim = None if similar is None else np.ascontiguousarray(similar, dtype=np.int32)
    X = np.zeros(32); info = np.zeros(4, dtype=np.int32); act = np.zeros(32, dtype=np.int32)
    m = int(pb["m"][b])
    A = np.ascontiguousarray(pb["DPu"][b]); bb = np.ascontiguousarray(pb["DPx"][b])
    lib().oracle_pldp_solve_sim(16, K.iPu.ctypes.data, K.Px.ctypes.data, K.Pu.ctypes.data, pb["D"][b].ctypes.data, m,
                                A.ctypes.data, bb.ctypes.data, pb["ZMPRef"][b].ctypes.data, pb["XkYk"][b].ctypes.data,
                                X.ctypes.data, int(n_removed), int(starting), None if hot is None else hot.ctypes.data,
                                int(hot_start), max_iter, info.ctypes.data, act.ctypes.data,
                                None if sim is None else sim.ctypes.data)
    return X, info, act


class RefPLDP:
    """The reference's PLDPSolver object (hot-start memory lives inside it, as in the reference)."""

    def __init__(self, K):
        self.L = ol.ref()
        if self.L is None:
            raise RuntimeError("oracle/_ref not built")
        self.L.ref_pldp_new.restype = C.c_void_p
        self.L.ref_pldp_new.argtypes = [C.c_uint] + [C.c_void_p] * 4
        self.L.ref_pldp_solve.restype = C.c_int
        self.L.ref_pldp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_uint] + [C.c_void_p] * 5 + [C.c_void_p, C.c_uint, C.c_uint, C.c_int]
        self.L.ref_pldp_delete.argtypes = [C.c_void_p]
        self.K = K
        self.iLQ2 = np.zeros((32, 32)); self.iLQ2[:16, :16] = K.iLQ; self.iLQ2[16:, 16:] = K.iLQ
        self.h = self.L.ref_pldp_new(16, K.iPu.ctypes.data, K.Px.ctypes.data, K.Pu.ctypes.data, self.iLQ2.ctypes.data)

    def solve(self, pb, b, starting=True, n_removed=0, similar=None):
        m = int(pb["m"][b])
        X = np.zeros(32)
        sim = np.zeros(8 * 16, dtype=np.int32) if similar is None else np.ascontiguousarray(similar, dtype=np.int32)
        D = np.ascontiguousarray(pb["D"][b]).copy(); A = np.ascontiguousarray(pb["DPu"][b]).copy()
        bb = np.ascontiguousarray(pb["DPx"][b]).copy(); z = pb["ZMPRef"][b].copy(); xk = pb["XkYk"][b].copy()
        rc = self.L.ref_pldp_solve(self.h, D.ctypes.data, m, A.ctypes.data, bb.ctypes.data, z.ctypes.data,
                                   xk.ctypes.data, X.ctypes.data, sim.ctypes.data, len(sim), int(n_removed), int(starting))
        return rc, X

    def close(self):
        """Deliberately leaks the reference object: ~PLDPSolver deletes m_iL after ~OptCholesky already did
        (PLDPSolver.cpp:172-175 / OptCholesky.cpp:52-57) - a double free in the reference."""
        self.h = None
