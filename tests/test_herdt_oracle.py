"""Pins for the Herdt2010 oracle (oracle/oracle_herdt.cpp, oracle/oracle_qp.cpp) - all CPU.

Golden vectors: the reference's own tests/TestHerdt2010OnLineTestFGPI.datref.cmake (ALL 22 348 rows: translations,
turns on the spot, curved walking, stop) and tests/TestHerdt2010EmergencyStopTestFGPI.datref.cmake (all 4 508 rows),
committed as tests/golden/herdt_{online,emergency}_full.npz by tests/golden/make_golden.py.  The reference compares
every column with |delta| < 1e-6 on 7-decimal-truncated values (tests/TestObject.cpp:475-495); so do we.

Finding (documented in DESIGN.md): the committed datrefs are older than the source (ChangeLog 3.1.8).  They are
reproduced, every row and column, by the surveyed code path with three datref-era settings (herdt_oracle.Sim.datref_era):
initial double-support frame (0, 0.1, 0) instead of the left-foot position (ZMPVelocityReferencedQP.cpp:277-279,
"Fix PG initialization"); hip-yaw velocity bound 0 (what OrientationsPreview.cpp:67 reads from the test robot) with
the default -30/+45 deg angle limits; and no return-to-centre jerk at the end of the walk (:410-421, "Put the CoM at
the center of the feet when stopping").  Each of the three is shown to matter below.
"""
import ctypes as C
import os

import numpy as np
import pytest

import herdt_oracle as ho

EVENTS = ho.ONLINE_EVENTS


@pytest.fixture(scope="module")
def online_run():
    sim, rows = ho.run_online_script(5000, EVENTS, initial_support=(0.0, 0.1, 0.0), logging=True)
    yield sim, rows
    sim.close()


def test_reference_qld_object_code_is_the_solver_here():
    if not os.path.exists(os.path.join(ho.ROOT, "oracle", "_ref", "libwalkgen_ref.so")):
        pytest.skip("oracle/_ref not built on this machine")
    assert ho.lib().oracle_herdt_have_ref_qld() == 1


def _check_full(name, events):
    gold = ho.load_golden(name)
    sim, rows = ho.run_online_script(len(gold), events, datref_era=True)
    sim.close()
    assert np.allclose(gold[:, 0], 0.005 * np.arange(1, len(gold) + 1), atol=1e-9)
    err = np.abs(rows[:, :36] - gold[:, 1:37])
    assert err.max() < 1e-6, (err.max(), np.unravel_index(err.argmax(), err.shape))
    return gold, rows, err


def test_oracle_reproduces_whole_online_datref():
    """All 22 348 rows x 36 columns of TestHerdt2010OnLine: 1117 QPs through the reference's own ql0001_."""
    gold, rows, err = _check_full("online", ho.ONLINE_EVENTS)
    assert gold.shape == (22348, 38)
    assert err.max() < 3.5e-7                                  # 1e-7 truncation + QLD's 1e-8 tolerance through the loop
    assert np.abs(gold[:, 4]).max() > 0.8 and np.abs(gold[:, 19]).max() > 40.0   # trunk yaw (rad), foot yaw (deg) exercised
    assert np.abs(rows[-1, :2] - gold[-1, 1:3]).max() < 1e-6


def test_oracle_reproduces_whole_emergency_stop_datref():
    gold, rows, err = _check_full("emergency", ho.EMERGENCY_EVENTS)
    assert gold.shape == (4508, 38)
    assert err.max() < 5e-7


def test_each_datref_era_setting_matters():
    """Surveyed-HEAD behaviour differs from the committed datref exactly where the three settings act."""
    gold = ho.load_golden("online")
    d = np.pi / 180.0
    # (2) a non-zero hip-yaw velocity bound (HRP-2's 3.54108 rad/s): first difference at the first QP after the
    # first rotation command (tick 5000 -> row 5028)
    sim = ho.Sim(); sim.steps_before_stop(2); sim.datref_era()
    ho.lib().oracle_herdt_sim_set_robot(sim.h, -30 * d, 45 * d, -30 * d, 45 * d, 3.54108)
    rows = np.zeros((5600, 37))
    for it in range(5600):
        rows[it] = sim.tick()[1]
        if it in ho.ONLINE_EVENTS:
            ho.ONLINE_EVENTS[it](sim)
    sim.close()
    bad = np.nonzero(np.abs(rows[:, :36] - gold[:5600, 1:37]).max(axis=1) > 1e-6)[0]
    assert bad[0] == 5028
    # (3) with the return-to-centre branch the emergency-stop trace leaves the datref only once the last step is taken
    golde = ho.load_golden("emergency")
    sim = ho.Sim(); sim.steps_before_stop(2); sim.datref_era()
    ho.lib().oracle_herdt_sim_set_return_to_centre(sim.h, 1)
    rows = np.zeros((len(golde), 37))
    for it in range(len(golde)):
        rows[it] = sim.tick()[1]
        if it in ho.EMERGENCY_EVENTS:
            ho.EMERGENCY_EVENTS[it](sim)
    sim.close()
    bad = np.nonzero(np.abs(rows[:, :36] - golde[:, 1:37]).max(axis=1) > 1e-6)[0]
    assert bad[0] > 3040 + 600


def test_oracle_reproduces_reference_datref_prefix(online_run):
    sim, rows = online_run
    gold = ho.load_golden("online")[:5000]
    assert np.allclose(gold[:, 0], 0.005 * np.arange(1, 5001), atol=1e-9)
    err = np.abs(rows[:, :36] - gold[:, 1:37])
    # datref is truncated to 7 decimals -> up to 1e-7 of quantisation; reference tolerance is 1e-6
    assert err.max() < 1e-6, (err.max(), np.unravel_index(err.argmax(), err.shape))
    assert err.max() < 2.5e-7
    assert ho.lib().oracle_herdt_sim_num_qp(sim.h) in (250, 251)  # one QP every 20 ticks (clock round-off decides the last)


def test_oracle_without_datref_era_override_differs_only_in_first_ds(online_run):
    """With the surveyed InitOnLine (support frame = left foot) only the QPs solved while the robot is still in
    its initial double support differ, and only laterally."""
    sim2, rows2 = ho.run_online_script(1300, EVENTS, logging=False)
    gold = ho.load_golden("online")
    err = np.abs(rows2[:, :36] - gold[:1300, 1:37])
    xcols = [0, 4, 7, 9, 21]  # CoM x, dx, ZMP x, LF x, RF x
    assert err[:, xcols].max() < 1e-6
    assert err[:1028].max() < 1e-6
    assert err[:, 1].max() > 1e-4  # CoM y does differ
    sim2.close()


def test_emergency_stop_rest_prefix():
    gold = ho.load_golden("emergency")[:1028]
    sim, rows = ho.run_online_script(1028, {}, initial_support=(0.0, 0.1, 0.0))
    assert np.abs(rows[:, :36] - gold[:, 1:37]).max() < 1e-6
    sim.close()


def _dense(p, inp):
    n = C.c_int(); m = C.c_int()
    Q = np.zeros(36 * 36); D = np.zeros(36); DU = np.zeros(77 * 36); DS = np.zeros(77)
    ho.lib().oracle_herdt_build_qp(C.byref(p), inp.ctypes.data, C.byref(n), C.byref(m), Q.ctypes.data, D.ctypes.data,
                                   DU.ctypes.data, DS.ctypes.data)
    n = n.value; m = m.value
    Qm = Q[:n * n].reshape(n, n).T
    A = DU[:(m + 1) * n].reshape(n, m + 1).T[:m]
    return n, m, Qm, D[:n], A, DS[:m]


def test_qp_shapes_and_dummy_row(online_run):
    sim, _ = online_run
    ins, X, U, meta = sim.log()
    p = ho.default_params()
    seen = set()
    for k in range(len(ins)):
        ns = int(ins[k]["sup_step"][16])
        n, m, Q, d, A, b = _dense(p, ins[k:k + 1])
        assert n == 32 + 2 * ns and m == 1 + 64 + 5 * ns           # qp-problem.cpp:249 m_ = NbConstraints_+1
        assert (meta[k, 0], meta[k, 1]) == (n, m)
        assert not A[0].any() and b[0] == 0.0                      # dummy row (qp-problem.cpp:428,440,512)
        assert np.allclose(Q, Q.T, atol=1e-15)
        assert np.linalg.eigvalsh(Q).min() > 0
        seen.add(ns)
    assert seen == {0, 1, 2}


def test_reference_qld_solutions_satisfy_kkt_and_match_textbook_solver(online_run):
    """Pins oracle_qp_solve (textbook Goldfarb-Idnani) against the reference's ql0001_ on the 250 QPs of the
    TestHerdt2010 prefix, and both against the KKT conditions."""
    if ho.lib().oracle_herdt_have_ref_qld() != 1:
        pytest.skip("oracle/_ref not built on this machine")
    sim, _ = online_run
    ins, X, U, meta = sim.log()
    p = ho.default_params()
    out = np.zeros(1, dtype=ho.QP_OUTPUT_DTYPE)
    for k in range(len(ins)):
        n, m, Q, d, A, b = _dense(p, ins[k:k + 1])
        assert meta[k, 2] == 0
        x = X[k, :n]; u = U[k, :m]
        # QLD: feasibility to its own 1e-8 normalised tolerance, stationarity
        rown = np.maximum(np.linalg.norm(A, axis=1), 1e-30)
        assert ((A @ x + b) / rown).min() > -2e-8
        assert (u >= 0).all()
        assert np.abs(Q @ x + d - A.T @ u).max() < 1e-7 * max(1.0, np.abs(d).max())
        rc = ho.lib().oracle_herdt_solve_qp(C.byref(p), ins[k:k + 1].ctypes.data, out.ctypes.data, 1)
        assert rc == 0
        xt = out["x"][0, :n]; ut = out["lagr"][0, :m]
        assert np.abs(xt - x).max() < 1e-6 * max(1.0, np.abs(x).max())
        # identical optimal active sets (positive multipliers), ignoring multipliers at round-off level
        big = max(u.max(), 1e-12)
        assert set(np.nonzero(u > 1e-7 * big)[0]) == set(np.nonzero(ut > 1e-7 * big)[0])
        # the textbook solver is exact to round-off
        assert ((A @ xt + b) / rown).min() > -1e-11
        assert np.abs(Q @ xt + d - A.T @ ut).max() < 1e-9 * max(1.0, np.abs(d).max())
