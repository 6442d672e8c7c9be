"""The reference's OWN ZMPConstrainedQPFastFormulation object (oracle/_ref, built from the reference source by oracle/Makefile)
against the oracle: the constants InitConstants() computes, and BuildZMPTrajectoryFromFootTrajectory (PLDP mode, the only
reachable one) on the feet buffers of one TestKajita2003 profile up to the period at which the reference stops (exit(0) on an
infeasible hot start - caught by oracle/ref_glue_dimitrov.cc -, or IFAIL on a NaN solution).
Run as a subprocess by tests/test_dimitrov.py.  TEST INFRASTRUCTURE ONLY.
    python tests/dimitrov_ref_object.py <profile> [<rows.npz> [<iPu.npy>]]  -> one JSON line ending the output."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import dimitrov_oracle as do  # noqa: E402
import oracle_lib as ol       # noqa: E402
import preview_ref as pr      # noqa: E402
import zmpdisc_oracle as zo   # noqa: E402

N = 16


class RefDimitrov:
    def __init__(self, par):
        self.r = ol.ref()
        self.r.ref_dimitrov_new.restype = C.c_void_p
        self.h = C.c_void_p(self.r.ref_dimitrov_new(C.c_double(par.sole_length), C.c_double(par.sole_width)))

    def constants(self):
        out = {"Px": np.zeros((N, 3)), "iPu": np.zeros((N, N)), "iLQ": np.zeros((2 * N, 2 * N)), "OptB": np.zeros((2 * N, 6)),
               "OptC": np.zeros((2 * N, 2 * N)), "Pu": np.zeros((N, N)), "PPu": np.zeros((N, N)), "VPu": np.zeros((N, N))}
        n = self.r.ref_dimitrov_constants(self.h, *[C.c_void_p(out[k].ctypes.data) for k in
                                                    ("Px", "iPu", "iLQ", "OptB", "OptC", "Pu", "PPu", "VPu")])
        assert n == N
        return out

    def set_ipu(self, iPu):
        a = np.ascontiguousarray(iPu, dtype=np.float64)
        self.r.ref_dimitrov_set_ipu(self.h, C.c_void_p(a.ctypes.data))

    def run(self, left, right, left_type, par):
        L, R = do._feet4(left), do._feet4(right)
        st = np.ascontiguousarray(left_type, dtype=np.int32)
        t = do.clock(len(L), par.sampling_period)
        zmp = np.zeros((len(L), 3)); com = np.zeros((len(L), 7))
        f = self.r.ref_dimitrov_run
        f.restype = C.c_int
        rc = f(self.h, C.c_long(len(L)), C.c_void_p(L.ctypes.data), C.c_void_p(R.ctypes.data), C.c_void_p(st.ctypes.data),
               C.c_void_p(t.ctypes.data), C.c_void_p(zmp.ctypes.data), C.c_void_p(com.ctypes.data),
               C.c_double(par.constraint_x), C.c_double(par.constraint_y), C.c_double(par.T), C.c_uint(N))
        return rc, com, zmp

    def close(self):
        self.r.ref_dimitrov_delete(self.h)


def main(name, dump=None, ipu_file=None):
    assert pr.lapack_available(), "no LAPACK (dgetrf_/dgetri_) for the reference's MAL_INVERSE"
    par = do.default_params()
    ref = RefDimitrov(par)
    res = {"profile": name}
    res_name = name
    # ---- InitConstants(): bitwise, except iPu which the reference takes from LAPACK's LU inverse
    Kr = ref.constants()
    K = do.Constants(par)
    res["constants_bitwise"] = bool(
        (Kr["Px"] == K.Px).all() and (Kr["Pu"] == K.Pu).all() and
        (Kr["iLQ"][:N, :N] == K.iLQ).all() and (Kr["iLQ"][N:, N:] == K.iLQ).all() and
        not Kr["iLQ"][:N, N:].any() and not Kr["iLQ"][N:, :N].any() and
        (Kr["OptB"][:N, :3] == K.OptB).all() and (Kr["OptB"][N:, 3:] == K.OptB).all() and
        (Kr["OptC"][:N, :N] == K.OptC).all() and (Kr["OptC"][N:, N:] == K.OptC).all())
    res["iPu_rel"] = float(np.abs(Kr["iPu"] - K.iPu).max() / np.abs(K.iPu).max())
    res["iPu_Pu_identity"] = float(np.abs(Kr["iPu"] @ Kr["Pu"] - np.eye(N)).max())
    # ---- the generator on the profile's feet
    # "<profile>@<k>": the generator is handed the feet buffers from sample k on (clock restarted at 0, CoM at rest at the
    # origin): every walk of the reference stops at its 18th period (below), so later parts of a walk - single support, rotated
    # feet, the duplicated half-planes of arcs - only reach the reference's loop this way
    name, _, skip = name.partition("@")
    skip = int(skip or 0)
    o = zo.run(zo.default_params(), zo.profile_steps(name))
    left, right, lt = o["left"][skip:].copy(), o["right"][skip:].copy(), o["types"][skip:, 1].copy()
    out = do.run(left, right, lt, par)                 # the oracle with the reference's semantics (no cold restart)
    rc_l, com_l, zmp_l = ref.run(left, right, lt, par)  # the reference as built: iPu from LAPACK's LU inverse
    ref.close()
    ref = RefDimitrov(par)
    ref.set_ipu(np.load(ipu_file) if ipu_file else K.iPu)   # the same inverse on both sides
    rc, com, zmp = ref.run(left, right, lt, par)
    ref.close()
    res["ref_rc_lapack_inverse"] = int(rc_l)
    per = out["periods"]
    stop = out["failed_at"]
    res["oracle_stop"] = None if stop is None else int(stop)
    res["oracle_stop_status"] = None if stop is None else int(per["status"][stop])
    res["ref_rc"] = int(rc)
    nper = len(per) if stop is None else stop           # periods both sides completed
    interval = int(round(par.T / par.sampling_period))
    rows = nper * interval                              # samples written by the completed periods
    res["periods_compared"] = int(nper)
    res["rows_compared"] = int(rows)
    if rows:
        res["com_err"] = float(np.abs(com[:rows, :6] - out["com"][:rows]).max())
        res["zmp_err"] = float(np.abs(zmp[:rows, :2] - out["zmp"][:rows]).max())
        res["com_bitwise"] = bool((com[:rows, :6] == out["com"][:rows]).all())
        res["com_span"] = float(np.abs(out["com"][:rows, 0]).max())
        res["com_err_lapack_inverse"] = float(np.abs(com_l[:rows, :6] - out["com"][:rows]).max())
        res["zmp_err_lapack_inverse"] = float(np.abs(zmp_l[:rows, :2] - out["zmp"][:rows]).max())
    if dump:      # the reference object's rows (same-inverse run) for a caller that compares something else with them
        np.savez(dump, com=com[:rows, :6], zmp=zmp[:rows, :2], com_lapack=com_l[:rows, :6], zmp_lapack=zmp_l[:rows, :2])
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
