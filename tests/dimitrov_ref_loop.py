"""Drives the receding-horizon loop of ZMPConstrainedQPFastFormulation::BuildZMPTrajectoryFromFootTrajectory (PLDP branch,
:1190-1400) through the reference's OWN PLDPSolver object with the REAL m_SimilarConstraints of the reference's own
FootConstraintsAsLinearSystem object code, and compares every period with the oracle loop (cold_restart on both sides).
Run as a subprocess by tests/test_dimitrov.py: the reference calls exit(0) on an infeasible hot start
(PLDPSolver.cpp:822-828), which must not be able to end the pytest process with a success code.
TEST INFRASTRUCTURE ONLY.   python tests/dimitrov_ref_loop.py <profile>   -> one JSON line ending the output."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import dimitrov_oracle as do  # noqa: E402
import pldp_oracle as po      # noqa: E402
import zmpdisc_oracle as zo   # noqa: E402


def main(name):
    o = zo.run(zo.default_params(), zo.profile_steps(name))
    left, right, lt = o["left"], o["right"], o["types"][:, 1].copy()
    par = do.default_params()
    par.cold_restart = 1
    out = do.run(left, right, lt, par)
    per = out["periods"]
    K = do.Constants(par)
    P_ref = do.ref_fcals(left, right, lt, par)        # polygons AND SimilarConstraints from the reference object code
    P_ora = do.fcals(left, right, lt, par)
    assert (P_ref["similar"] == P_ora["similar"]).all() and P_ref["A"].tobytes() == P_ora["A"].tobytes()
    ref = po.RefPLDP(K)
    removed, starting, st = 0, True, 0.0
    compared = restarts = flagged_rows = 0
    last = len(per) if out["failed_at"] is None else out["failed_at"]
    for li in range(last):
        xk = per["xk"][li].copy()
        pb = do.build_constraints(K, P_ref, st, xk)
        sim, m = do.similar_flags(K, P_ref, st)
        assert m == pb["m"] == per["m"][li]
        flagged_rows += int((sim[:m] != 0).sum())
        if per["status"][li] == 5:
            # the oracle solved this period again from the cold start point; the reference side does the same with a
            # fresh solver object (no kept constraints, no previous ZMP solution) - calling the old object on the
            # infeasible hot start would end this process (exit(0))
            ref.close()
            ref = po.RefPLDP(K)
            starting, removed = True, 0
            restarts += 1
        packed = {"m": np.array([pb["m"]]), "DPu": pb["DPu"][None], "DPx": pb["DPx"][None], "D": pb["D"][None],
                  "ZMPRef": pb["ZMPRef"][None], "XkYk": pb["XkYk"][None]}
        rc, X = ref.solve(packed, 0, starting=starting, n_removed=removed, similar=sim)
        assert rc == 0, (li, rc)
        jx = jy = 0.0
        for j in range(16):
            jx += K.iLQ[j, 0] * X[j]
            jy += K.iLQ[j, 0] * X[j + 16]
        assert jx == per["jerk_x"][li] and jy == per["jerk_y"][li], (name, li, jx, per["jerk_x"][li])
        compared += 1
        starting = False
        removed = pb["n_first"]
        st += par.T
    # the period at which the oracle stops for good (NaN on a duplicated half-plane): the reference returns -1 there too
    nan_stop = None
    if out["failed_at"] is not None and per["rc"][last] != 0:
        xk = per["xk"][last].copy()
        pb = do.build_constraints(K, P_ref, st, xk)
        sim, m = do.similar_flags(K, P_ref, st)
        packed = {"m": np.array([pb["m"]]), "DPu": pb["DPu"][None], "DPx": pb["DPx"][None], "D": pb["D"][None],
                  "ZMPRef": pb["ZMPRef"][None], "XkYk": pb["XkYk"][None]}
        if per["status"][last] in (0, 5):
            if per["status"][last] == 5:
                ref = po.RefPLDP(K); starting, removed = True, 0
            rc, X = ref.solve(packed, 0, starting=starting, n_removed=removed, similar=sim)
            nan_stop = int(rc)
    print(json.dumps({"profile": name, "periods": int(len(per)), "compared": compared, "restarts": restarts,
                      "failed_at": out["failed_at"], "flagged_rows": flagged_rows, "ref_rc_at_stop": nan_stop,
                      "done": True}))


if __name__ == "__main__":
    main(sys.argv[1])
