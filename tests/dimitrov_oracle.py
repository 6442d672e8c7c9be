"""ctypes access to the Dimitrov2008 oracle (oracle/oracle_dimitrov.cpp) and to the reference's own ConvexHull.cpp /
FootConstraintsAsLinearSystem.cpp object code (oracle/_ref).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

import oracle_lib as ol

LCI = np.dtype([("A", "f8", (8, 2)), ("B", "f8", 8), ("center", "f8", 2), ("t_start", "f8"), ("t_end", "f8"),
                ("rows", "i4"), ("first_sample", "i4"), ("state", "i4"), ("rc", "i4"), ("similar", "i4", 8)])
assert LCI.itemsize == 272
PERIOD = np.dtype([("t_start", "f8"), ("xk", "f8", 6), ("jerk_x", "f8"), ("jerk_y", "f8"), ("m", "i4"), ("n_first", "i4"),
                   ("rc", "i4"), ("status", "i4"), ("iterations", "i4"), ("n_active", "i4"), ("active", "i4", 32)])
assert PERIOD.itemsize == 224


class Params(C.Structure):
    _fields_ = [("T", C.c_double), ("sampling_period", C.c_double), ("com_height", C.c_double), ("alpha", C.c_double),
                ("beta", C.c_double), ("constraint_x", C.c_double), ("constraint_y", C.c_double),
                ("sole_length", C.c_double), ("sole_width", C.c_double), ("max_iterations", C.c_int32),
                ("cold_restart", C.c_int32), ("merge_duplicate_rows", C.c_int32), ("reserved", C.c_int32)]


def default_params():
    """ZMPConstrainedQPFastFormulation.cpp:79-96 + the sole of the HRP-2 test robot (0.25 x 0.14)."""
    return Params(0.1, 0.005, 0.80, 200.0, 1000.0, 0.04, 0.04, 0.25, 0.14, 0, 0, 0, 0)


def clock(n, Ts=0.005):
    """The accumulated m_CurrentTime of ZMPDiscretization: sample i carries Ts added i times."""
    return np.concatenate([[0.0], np.cumsum(np.full(n - 1, Ts))]) if n > 1 else np.zeros(n)


def hull(points):
    p = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros((16, 2))
    n = ol.oracle().oracle_convex_hull(len(p), C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data), 16)
    return out[:n]


def ref_hull(points):
    p = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros((16, 2))
    n = ol.ref().ref_convex_hull(len(p), C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data), 16)
    return out[:n]


def _feet4(f):
    """(x, y, z, theta) columns of a [n][6] foot buffer."""
    return np.ascontiguousarray(f[:, :4], dtype=np.float64)


def fcals(left, right, left_type, par=None, cap=512):
    par = par or default_params()
    L, R = _feet4(left), _feet4(right)
    st = np.ascontiguousarray(left_type, dtype=np.int32)
    t = clock(len(L), par.sampling_period)
    out = np.zeros(cap, dtype=LCI)
    n = ol.oracle().oracle_fcals_build(len(L), C.c_void_p(L.ctypes.data), C.c_void_p(R.ctypes.data),
                                       C.c_void_p(st.ctypes.data), C.c_void_p(t.ctypes.data),
                                       C.c_double(par.sole_length), C.c_double(par.sole_width),
                                       C.c_double(par.constraint_x), C.c_double(par.constraint_y), cap,
                                       C.c_void_p(out.ctypes.data), int(par.merge_duplicate_rows))
    assert 0 <= n <= cap, n
    return out[:n]


def ref_fcals(left, right, left_type, par=None, cap=512):
    """The reference's own BuildLinearConstraintInequalities (oracle/ref_glue.cc:ref_fcals_build) -> LCI records
    (first_sample / state / rc are not reported by the reference and stay 0)."""
    par = par or default_params()
    L, R = _feet4(left), _feet4(right)
    st = np.ascontiguousarray(left_type, dtype=np.int32)
    t = clock(len(L), par.sampling_period)
    rows = np.zeros(cap, dtype=np.int32); A = np.zeros((cap, 8, 2)); B = np.zeros((cap, 8)); cen = np.zeros((cap, 2))
    sim = np.zeros((cap, 8), dtype=np.int32); tt = np.zeros((cap, 2))
    f = ol.ref().ref_fcals_build
    f.restype = C.c_int
    n = f(len(L), C.c_void_p(L.ctypes.data), C.c_void_p(R.ctypes.data), C.c_void_p(st.ctypes.data),
          C.c_void_p(t.ctypes.data), C.c_double(par.sole_length), C.c_double(par.sole_width),
          C.c_double(par.constraint_x), C.c_double(par.constraint_y), cap, C.c_void_p(rows.ctypes.data),
          C.c_void_p(A.ctypes.data), C.c_void_p(B.ctypes.data), C.c_void_p(cen.ctypes.data),
          C.c_void_p(sim.ctypes.data), C.c_void_p(tt.ctypes.data))
    assert 0 <= n <= cap, n
    out = np.zeros(n, dtype=LCI)
    out["rows"], out["A"], out["B"], out["center"], out["similar"] = rows[:n], A[:n], B[:n], cen[:n], sim[:n]
    out["t_start"], out["t_end"] = tt[:n, 0], tt[:n, 1]
    return out


class Constants:
    def __init__(self, par=None, N=16):
        par = par or default_params()
        self.N, self.T, self.zc = N, par.T, par.com_height
        self.iPu = np.zeros((N, N)); self.Px = np.zeros((N, 3)); self.Pu = np.zeros((N, N)); self.iLQ = np.zeros((N, N))
        self.OptB = np.zeros((N, 3)); self.OptC = np.zeros((N, N))
        ol.oracle().oracle_dimitrov_constants(N, C.c_double(par.T), C.c_double(par.com_height), C.c_double(par.alpha),
                                              C.c_double(par.beta), *[C.c_void_p(a.ctypes.data) for a in
                                                                      (self.iPu, self.Px, self.Pu, self.iLQ, self.OptB, self.OptC)])


def build_constraints(K, lci, t_start, xk):
    """oracle_dimitrov_build_constraints -> dict in the layout of workloads.pldp_problem_from (+ first polygon)."""
    N = K.N
    DPu = np.zeros((8 * N + 1) * 2 * N); DPx = np.zeros(8 * N + 1); zref = np.zeros(2 * N); D = np.zeros(2 * N)
    first = np.zeros(2, dtype=np.int32)
    xk = np.ascontiguousarray(xk, dtype=np.float64)
    lci = np.ascontiguousarray(lci)
    m = ol.oracle().oracle_dimitrov_build_constraints(N, C.c_double(K.T), C.c_double(t_start), len(lci),
                                                      C.c_void_p(lci.ctypes.data), *[C.c_void_p(a.ctypes.data) for a in
                                                      (K.Px, K.Pu, K.OptB, K.OptC, xk, DPu, DPx, zref, D, first)])
    if m < 0:
        raise RuntimeError(f"build_constraints: {m}")
    return {"D": D, "m": m, "DPu": DPu[:(m + 1) * 2 * N].copy(), "DPx": DPx[:m].copy(), "ZMPRef": zref, "XkYk": xk,
            "n_first": int(first[1]), "first": int(first[0])}


def similar_flags(K, lci, t_start):
    """m_SimilarConstraints for the period starting at t_start (ZMPConstrainedQPFastFormulation.cpp:889)."""
    sim = np.zeros(8 * K.N, dtype=np.int32)
    lci = np.ascontiguousarray(lci)
    m = ol.oracle().oracle_dimitrov_similar(K.N, C.c_double(K.T), C.c_double(t_start), len(lci), C.c_void_p(lci.ctypes.data),
                                            C.c_void_p(sim.ctypes.data))
    if m < 0:
        raise RuntimeError(f"similar_flags: {m}")
    return sim, m


def period_count(n, par=None):
    par = par or default_params()
    f = ol.oracle().oracle_dimitrov_period_count
    f.restype = C.c_long
    return f(16, C.c_double(par.T), C.c_double(par.sampling_period), C.c_long(n))


def run(left, right, left_type, par=None, hot_start=True):
    """oracle_dimitrov_run -> dict(com [n][6], zmp [n][2], periods)."""
    par = par or default_params()
    L, R = _feet4(left), _feet4(right)
    st = np.ascontiguousarray(left_type, dtype=np.int32)
    n = len(L)
    cap = period_count(n, par) + 4
    com = np.zeros((n, 6)); zmp = np.zeros((n, 2)); per = np.zeros(cap, dtype=PERIOD)
    f = ol.oracle().oracle_dimitrov_run
    f.restype = C.c_long
    k = f(C.byref(par), C.c_long(n), C.c_void_p(L.ctypes.data), C.c_void_p(R.ctypes.data), C.c_void_p(st.ctypes.data),
          C.c_void_p(com.ctypes.data), C.c_void_p(zmp.ctypes.data), C.c_long(cap), C.c_void_p(per.ctypes.data),
          int(hot_start))
    if k == -2 or k == -4:
        raise RuntimeError(f"oracle_dimitrov_run failed: {k}")
    failed_at = None
    if k <= -100:          # the walk stopped at period -(100+k): the reference's exit(0) / IFAIL path
        failed_at = -k - 100
        k = failed_at + 1
    return {"com": com, "zmp": zmp, "periods": per[:k], "failed_at": failed_at}
