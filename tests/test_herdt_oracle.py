"""Pins for the Herdt2010 oracle (oracle/oracle_herdt.cpp, oracle/oracle_qp.cpp) - all CPU.

Golden vector: the reference's own tests/TestHerdt2010OnLineTestFGPI.datref.cmake (rows 0..4999, t < 25 s),
committed as tests/golden/herdt_online_prefix.npz by tests/golden/make_golden.py.  The reference compares every
column with |delta| < 1e-6 on 7-decimal-truncated values (tests/TestObject.cpp:475-495); so do we.

Finding (documented in DESIGN.md): the committed datref was generated before ChangeLog 3.1.8 "Fix PG
initialization": its first velocity event is only reproduced when the initial double-support frame is
(0, 0.1, 0) instead of the left-foot position the surveyed InitOnLine uses
(ZMPVelocityReferencedQP.cpp:277-279).  Everything else is the surveyed code path, restated.
"""
import ctypes as C
import os

import numpy as np
import pytest

import herdt_oracle as ho

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EVENTS = {5 * 200: lambda s: s.vel_ref(0.2, 0.0, 0.0),      # walkForward      tests/TestHerdt2010.cpp:232
          10 * 200: lambda s: s.vel_ref(0.0, 0.2, 0.0)}     # walkSidewards    tests/TestHerdt2010.cpp:233


@pytest.fixture(scope="module")
def online_run():
    sim, rows = ho.run_online_script(5000, EVENTS, initial_support=(0.0, 0.1, 0.0), logging=True)
    yield sim, rows
    sim.close()


def test_reference_qld_object_code_is_the_solver_here():
    if not os.path.exists(os.path.join(ho.ROOT, "oracle", "_ref", "libwalkgen_ref.so")):
        pytest.skip("oracle/_ref not built on this machine")
    assert ho.lib().oracle_herdt_have_ref_qld() == 1


def test_oracle_reproduces_reference_datref_prefix(online_run):
    sim, rows = online_run
    gold = np.load(os.path.join(GOLD, "herdt_online_prefix.npz"))["q"] / 1e7
    assert gold.shape == (5000, 38)
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
    gold = np.load(os.path.join(GOLD, "herdt_online_prefix.npz"))["q"] / 1e7
    err = np.abs(rows2[:, :36] - gold[:1300, 1:37])
    xcols = [0, 4, 7, 9, 21]  # CoM x, dx, ZMP x, LF x, RF x
    assert err[:, xcols].max() < 1e-6
    assert err[:1028].max() < 1e-6
    assert err[:, 1].max() > 1e-4  # CoM y does differ
    sim2.close()


def test_emergency_stop_rest_prefix():
    gold = np.load(os.path.join(GOLD, "herdt_emergency_prefix.npz"))["q"] / 1e7
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
