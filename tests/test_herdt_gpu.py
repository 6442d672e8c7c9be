"""Herdt2010 QP: CUDA (wg_herdt_qp_solve_batch through the C ABI) against the oracle.

Oracle = oracle/oracle_herdt.cpp (assembly restated from generator-vel-ref.cpp / qp-problem.cpp, pinned to the
reference datref by tests/test_herdt_oracle.py) + the reference's own ql0001_ (oracle/_ref) or the textbook
Goldfarb-Idnani solver (oracle/oracle_qp.cpp).  Tolerances (north_star): identical optimal active sets and
foot placements, solution within 1e-6, KKT residuals <= 1e-9.
"""
import ctypes as C

import numpy as np
import pytest

import herdt_oracle as ho

N = 16


def logged_qps(nticks, events, **kw):
    sim, rows = ho.run_online_script(nticks, events, logging=True, **kw)
    ins, X, U, meta = sim.log()
    sim.close()
    return ins.copy(), X.copy(), U.copy(), meta.copy()


def random_walk_qps(seed, nsims=12, nticks=2400):
    """QP inputs logged from oracle closed-loop runs under random velocity references, yaw rate included
    (rotated hulls, trunk-yaw dependent global reference), with reference changes every 2-4 s."""
    rng = np.random.default_rng(seed)
    all_ins = []
    for _ in range(nsims):
        ev = {}
        t = 20
        while t < nticks:
            v = (rng.uniform(-0.2, 0.3), rng.uniform(-0.15, 0.15), rng.uniform(-0.2, 0.2))
            if rng.random() < 0.15:
                v = (0.0, 0.0, 0.0)
            ev[t] = (lambda vv: (lambda s: s.vel_ref(*vv)))(v)
            t += int(rng.integers(400, 800))
        ins, X, U, meta = logged_qps(nticks, ev)
        all_ins.append(ins)
    return np.concatenate(all_ins)


def oracle_solve(ins, solver):
    p = ho.default_params()
    out = np.zeros(len(ins), dtype=ho.QP_OUTPUT_DTYPE)
    ho.lib().oracle_herdt_solve_qp_batch(C.byref(p), len(ins), ins.ctypes.data, out.ctypes.data, solver)
    return out


def dense(ins_k):
    p = ho.default_params()
    n = C.c_int(); m = C.c_int()
    Q = np.zeros(36 * 36); D = np.zeros(36); DU = np.zeros(77 * 36); DS = np.zeros(77)
    ho.lib().oracle_herdt_build_qp(C.byref(p), ins_k.ctypes.data, C.byref(n), C.byref(m), Q.ctypes.data, D.ctypes.data,
                                   DU.ctypes.data, DS.ctypes.data)
    n = n.value; m = m.value
    return n, m, Q[:n * n].reshape(n, n).T, D[:n], DU[:(m + 1) * n].reshape(n, m + 1).T[:m], DS[:m]


def check_against(ins, gpu, ref, xtol, check_kkt=True):
    assert (gpu["fail"] == 0).all(), np.nonzero(gpu["fail"])[0][:10]
    assert np.array_equal(gpu["n_vars"], ref["n_vars"]) and np.array_equal(gpu["n_rows"], ref["n_rows"])
    worst_x = 0.0; worst_kkt = 0.0; worst_feas = 0.0
    for k in range(len(ins)):
        n, m = int(ref["n_vars"][k]), int(ref["n_rows"][k])
        x, xr = gpu["x"][k, :n], ref["x"][k, :n]
        scale = max(1.0, np.abs(xr).max())
        worst_x = max(worst_x, np.abs(x - xr).max() / scale)
        assert np.abs(x - xr).max() < xtol * scale, (k, np.abs(x - xr).max())
        # foot placements
        assert np.abs(x[2 * N:] - xr[2 * N:]).max() < 1e-6 if n > 2 * N else True
        u, ur = gpu["lagr"][k, :m], ref["lagr"][k, :m]
        big = max(ur.max(), 1e-12)
        assert set(np.nonzero(u > 1e-7 * big)[0]) == set(np.nonzero(ur > 1e-7 * big)[0]), k
        assert (u >= 0).all()
        if check_kkt:
            n_, m_, Q, d, A, b = dense(ins[k:k + 1])
            rown = np.maximum(np.linalg.norm(A, axis=1), 1e-30)
            feas = ((A @ x + b) / rown).min()
            stat = np.abs(Q @ x + d - A.T @ u).max() / max(1.0, np.abs(d).max())
            comp = np.abs(u * (A @ x + b)).max() / max(1.0, np.abs(d).max())
            worst_kkt = max(worst_kkt, stat, comp); worst_feas = min(worst_feas, feas)
            assert feas > -1e-9 and stat < 1e-9 and comp < 1e-9, (k, feas, stat, comp)
    return worst_x, worst_kkt, worst_feas


@pytest.fixture(scope="module")
def hctx(ctx):
    ctx.herdt_set_params()
    return ctx


@pytest.mark.gpu
def test_gpu_qp_matches_reference_qld_on_testherdt2010_prefix(hctx):
    """The 250 QPs of the TestHerdt2010 OnLine prefix (the run pinned to the reference datref)."""
    ev = {1000: lambda s: s.vel_ref(0.2, 0.0, 0.0), 2000: lambda s: s.vel_ref(0.0, 0.2, 0.0)}
    ins, X, U, meta = logged_qps(5000, ev, initial_support=(0.0, 0.1, 0.0))
    gpu = hctx.herdt_qp_solve(ins)
    # against the solver the oracle used in that run (reference ql0001_ when oracle/_ref is present)
    ref = np.zeros(len(ins), dtype=ho.QP_OUTPUT_DTYPE)
    ref["x"] = X; ref["lagr"] = U; ref["n_vars"] = meta[:, 0]; ref["n_rows"] = meta[:, 1]
    wx, wk, wf = check_against(ins, gpu, ref, xtol=1e-6)
    # and, tighter, against the textbook solver
    tb = oracle_solve(ins, 1)
    wx2, _, _ = check_against(ins, gpu, tb, xtol=1e-8, check_kkt=False)
    assert np.abs(gpu["com_next_x"] - tb["com_next_x"]).max() < 1e-8
    assert np.abs(gpu["com_next_y"] - tb["com_next_y"]).max() < 1e-8
    print(f"prefix: max rel |x-x_qld| {wx:.2e}, |x-x_textbook| {wx2:.2e}, kkt {wk:.2e}, feas {wf:.2e}, "
          f"iterations mean {gpu['iterations'].mean():.1f} max {gpu['iterations'].max()}")


@pytest.mark.gpu
def test_gpu_qp_random_velocity_references_with_rotation(hctx):
    ins = random_walk_qps(7)
    assert len(ins) > 1000
    assert {0, 1, 2} <= set(int(v) for v in ins["sup_step"][:, N])
    assert np.abs(ins["sup_yaw"]).max() > 0.05          # rotated hulls are exercised
    tb = oracle_solve(ins, 1)
    ok = tb["fail"] == 0
    assert ok.mean() > 0.99
    gpu = hctx.herdt_qp_solve(ins)
    assert (gpu["fail"][ok] == 0).all()
    wx, wk, wf = check_against(ins[ok], gpu[ok], tb[ok], xtol=1e-7)
    print(f"random: {ok.sum()} QPs, max rel |x-x_textbook| {wx:.2e}, kkt {wk:.2e}, feas {wf:.2e}, "
          f"iterations mean {gpu['iterations'].mean():.1f} max {gpu['iterations'].max()}")


@pytest.mark.gpu
def test_gpu_qp_device_memory_and_batch_edges(hctx):
    import jrl_walkgen_b200 as wg
    ev = {100: lambda s: s.vel_ref(0.2, 0.05, 0.0)}
    ins, X, U, meta = logged_qps(1200, ev)
    host = hctx.herdt_qp_solve(ins)
    for B in (1, 3, len(ins)):
        d_in = hctx.to_device(ins[:B]); d_out = hctx.alloc(B * wg.QP_OUTPUT_DTYPE.itemsize)
        hctx.herdt_qp_solve(d_in, d_out, mem=wg.WG_MEM_DEVICE, count=B)
        hctx.sync()
        dev = d_out.download(wg.QP_OUTPUT_DTYPE, (B,))
        assert np.array_equal(dev["x"], host["x"][:B]) and np.array_equal(dev["lagr"], host["lagr"][:B])
        d_in.free(); d_out.free()
    # empty batch is a no-op
    assert len(hctx.herdt_qp_solve(ins[:0])) == 0


@pytest.mark.gpu
def test_gpu_qp_full_size_idempotence(hctx):
    """BASELINE config 3 size (16 384 instances): tiling the logged problems must give identical answers
    for identical inputs, whatever warp/block solved them."""
    ev = {100: lambda s: s.vel_ref(0.25, -0.05, 0.1)}
    ins, X, U, meta = logged_qps(2400, ev)
    reps = -(-16384 // len(ins))
    big = np.tile(ins, reps)[:16384]
    out = hctx.herdt_qp_solve(big)
    assert (out["fail"] == 0).all()
    base = out[:len(ins)]
    for r in range(1, reps):
        seg = out[r * len(ins):(r + 1) * len(ins)]
        assert np.array_equal(seg["x"], base["x"][:len(seg)])
        assert np.array_equal(seg["lagr"], base["lagr"][:len(seg)])


@pytest.mark.gpu
def test_gpu_qp_warm_start_reaches_the_cold_optimum(hctx):
    """wg_herdt_qp_solve_batch_warm: the optimum does not depend on the guess.  (i) the QP's own optimal set as the guess
    (age 0) is accepted in about one change per active row; (ii) consecutive closed-loop QPs, each warm started from its
    predecessor's optimal set shifted by one sample (age 1: the closed loop's use), take far fewer iterations than cold;
    (iii) a garbage guess (all rows of a random subset) still ends on the same optimum.  Reference optimum: the textbook
    solver; x within 1e-8, identical active sets, KKT 1e-9 (check_against)."""
    import jrl_walkgen_b200 as wg
    ev = {200: lambda s: s.vel_ref(0.2, 0.05, 0.1), 1400: lambda s: s.vel_ref(-0.1, 0.1, -0.15),
          2600: lambda s: s.vel_ref(0.0, 0.0, 0.0)}
    ins, X, U, meta = logged_qps(3600, ev)
    tb = oracle_solve(ins, 1)
    ok = tb["fail"] == 0
    assert ok.all()
    cold, act = hctx.herdt_qp_solve_warm(ins)                      # no guess: cold start, returns the active sets
    check_against(ins, cold, tb, xtol=1e-8)
    plain = hctx.herdt_qp_solve(ins)
    assert np.array_equal(plain["x"], cold["x"]) and np.array_equal(plain["iterations"], cold["iterations"])
    for k in range(len(ins)):
        m = int(tb["n_rows"][k])
        rows = set(int(r) for r in act["rows"][k, :act["n"][k]])
        assert set(np.nonzero(tb["lagr"][k, 1:m] > 1e-7 * max(tb["lagr"][k].max(), 1e-12))[0]) <= rows, k
    # (i) own optimal set
    own, act2 = hctx.herdt_qp_solve_warm(ins, guess=act, age=0)
    check_against(ins, own, tb, xtol=1e-8)
    assert (own["iterations"] <= act["n"] + 4).all()
    # (ii) predecessor's set, one period older
    guess = np.zeros(len(ins), dtype=wg.ACTIVE_SET_DTYPE)
    guess[1:] = act[:-1]
    warm, _ = hctx.herdt_qp_solve_warm(ins, guess=guess, age=1)
    check_against(ins, warm, tb, xtol=1e-8)
    # (iii) garbage
    rng = np.random.default_rng(3)
    bad = np.zeros(len(ins), dtype=wg.ACTIVE_SET_DTYPE)
    bad["rows"][:] = -1
    for k in range(len(ins)):
        n = int(rng.integers(1, 30))
        bad["rows"][k, :n] = rng.choice(74, size=n, replace=False)
        bad["n"][k] = n
        bad["step_pi"][k] = rng.integers(0, 17, size=2)
    junk, _ = hctx.herdt_qp_solve_warm(ins, guess=bad, age=int(rng.integers(0, 3)))
    check_against(ins, junk, tb, xtol=1e-8)
    print(f"warm start: iterations cold {cold['iterations'].mean():.1f}, own set {own['iterations'].mean():.1f}, "
          f"previous period {warm['iterations'].mean():.1f}, garbage {junk['iterations'].mean():.1f}")
    assert warm["iterations"].mean() < cold["iterations"].mean()
