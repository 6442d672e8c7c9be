"""Dimitrov PLDP solver and OptCholesky on the GPU (wg_pldp_solve_batch / wg_optcholesky_* through the C ABI) against
the oracle port (oracle/oracle_pldp.cpp, bitwise-pinned to the reference's own PLDPSolver/OptCholesky object code by
tests/test_pldp_oracle.py) and, when oracle/_ref travelled to this machine, against that object code directly.

Bar: bitwise equal X, identical sequences of activated constraints, iteration counts and status (the kernel keeps the
reference's operation order with non-fused arithmetic, see pldp.cu); OptCholesky rows bitwise equal.
"""
import numpy as np
import pytest

import oracle_lib as ol
import pldp_oracle as po
from jrl_walkgen_b200 import workloads as W


@pytest.fixture(scope="module")
def pctx(ctx):
    K = W.DimitrovConstants()
    ctx.pldp_set_constants(K.iPu, K.Px, K.Pu)
    return ctx, K


@pytest.mark.gpu
def test_gpu_pldp_cold_batch_matches_oracle_and_reference(pctx):
    ctx, K = pctx
    K2, pb = W.pldp_batch(512, seed=11, K=K)
    X, info = ctx.pldp_solve(pb)
    assert (info["rc"] == 0).all() and (info["status"] == 0).all()
    worst = 0.0
    for b in range(512):
        Xo, io, act = po.oracle_solve(K, pb, b, starting=True)
        assert info["iterations"][b] == io[2] and info["n_active"][b] == io[3]
        assert np.array_equal(info["active"][b], act), b
        worst = max(worst, np.abs(X[b] - Xo).max())
    assert worst == 0.0
    assert info["n_active"].max() >= 10
    if ol.ref() is not None:
        for b in range(0, 512, 16):
            ref = po.RefPLDP(K)
            rc, Xr = ref.solve(pb, b, starting=True)
            assert rc == 0 and np.array_equal(X[b], Xr)
    print(f"cold: max |X_gpu - X_oracle| {worst:.2e}, iterations mean {info['iterations'].mean():.1f} "
          f"max {info['iterations'].max()}, active max {info['n_active'].max()}")


@pytest.mark.gpu
def test_gpu_pldp_hot_started_receding_horizon(pctx):
    """64 walks x 20 receding-horizon steps with the hot start of the reference (previous active set shifted by the
    rows that left the horizon, previous ZMP plan as the start point); every step against the oracle."""
    import jrl_walkgen_b200 as wg
    ctx, K = pctx
    B, T = 64, 20
    polys = [W._support_polygons(np.random.default_rng([21, b]), K.N, K.T, count=K.N + T) for b in range(B)]
    xk = np.zeros((B, 6))
    for b in range(B):
        xk[b, 0], xk[b, 3] = polys[b][0][0]
    xo = xk.copy()
    hot = np.zeros(B, dtype=wg.PLDP_STATE_DTYPE)
    hot_o = np.zeros(B, dtype=po.STATE)
    n_removed = np.zeros(B, dtype=np.int32)
    alive = np.ones(B, bool)
    kept = 0
    for t in range(T):
        probs = [W.pldp_problem_from(K, polys[b][t:t + K.N], xk[b]) for b in range(B)]
        pb = W.pldp_pack(K, probs)
        X, info = ctx.pldp_solve(pb, hot=hot, starting=np.full(B, int(t == 0), dtype=np.int32), n_removed=n_removed)
        for b in range(B):
            if not alive[b]:
                continue
            p_o = W.pldp_problem_from(K, polys[b][t:t + K.N], xo[b])
            Xo, io, act = po.oracle_solve(K, W.pldp_pack(K, [p_o]), 0, hot=hot_o[b:b + 1], starting=(t == 0),
                                          n_removed=int(n_removed[b]))
            assert info["status"][b] == io[1], (t, b)
            if io[1] != 0:                     # the reference would exit(0) here (overshoot by m_tol, see the oracle test)
                alive[b] = False
                continue
            assert info["iterations"][b] == io[2] and np.array_equal(info["active"][b], act), (t, b)
            assert np.array_equal(X[b], Xo), (t, b, np.abs(X[b] - Xo).max())
            assert hot["n_prev"][b] == hot_o["n_prev"][b]
            assert np.array_equal(hot["prev_active"][b][:hot["n_prev"][b]], hot_o["prev_active"][b][:hot_o["n_prev"][b]])
            kept += int(hot["n_prev"][b])
            xk[b] = W.pldp_advance(K, xk[b], X[b]); xo[b] = W.pldp_advance(K, xo[b], Xo)
        n_removed = np.array([p["n_first"] for p in probs], dtype=np.int32)
    assert alive.sum() >= B // 2 and kept > 0


@pytest.mark.gpu
def test_gpu_optcholesky_matches_oracle_bitwise(pctx):
    ctx, K = pctx
    rng = np.random.default_rng(0)
    # tests/TestOptCholesky.cpp: 12 x 15 uniform matrix, MODE_NORMAL, rows added one by one
    B = 7
    A = rng.uniform(0.0, 1.0, size=(B, 12, 15))
    rows = np.tile(np.arange(12, dtype=np.int32), (B, 1))
    L = np.zeros((B, 12, 12))
    for k in range(12):
        ctx.optcholesky_add_rows(A.reshape(B, -1), rows, L.reshape(B, -1), 0, 12, 15, 12, k, k + 1)
    Lo = np.zeros((B, 12, 12))
    for b in range(B):
        po.lib().oracle_optcholesky_add_rows(0, 12, 15, 12, A[b].ctypes.data, rows[b].ctypes.data, 0, 12, Lo[b].ctypes.data)
        assert np.linalg.norm(A[b] @ A[b].T - L[b] @ L[b].T) <= 1e-6          # the reference test's own check
    assert np.array_equal(L, Lo)
    # MODE_FORTRAN: column-major constraint matrix with leading dimension nb + 1, a random subset of rows at once
    nb, cu = 40, 32
    A2 = rng.uniform(-1.0, 1.0, size=(B, cu, nb + 1))
    rows2 = np.stack([rng.permutation(nb)[:20] for _ in range(B)]).astype(np.int32)
    L2 = np.zeros((B, nb, nb)); L2o = np.zeros((B, nb, nb))
    ctx.optcholesky_add_rows(A2.reshape(B, -1), rows2, L2.reshape(B, -1), 1, nb, cu, nb, 0, 20)
    for b in range(B):
        po.lib().oracle_optcholesky_add_rows(1, nb, cu, nb, A2[b].ctypes.data, rows2[b].ctypes.data, 0, 20, L2o[b].ctypes.data)
    assert np.array_equal(L2, L2o)
    # ComputeNormalCholeskyOnANormal + ComputeInverseCholeskyNormal(1) (TestOptCholesky.cpp:155-177)
    M = np.stack([a @ a.T for a in A])
    Lf, iL = ctx.optcholesky_full(M)
    for b in range(B):
        Lb = np.zeros((12, 12)); iLb = np.zeros((12, 12))
        po.lib().oracle_optcholesky_full(12, M[b].ctypes.data, Lb.ctypes.data)
        po.lib().oracle_optcholesky_inverse(12, 12, Lb.ctypes.data, iLb.ctypes.data)
        assert np.array_equal(Lf[b], Lb) and np.array_equal(iL[b], iLb)
        assert np.abs(iL[b] @ Lf[b] - np.eye(12)).max() < 1e-9


@pytest.mark.gpu
def test_gpu_pldp_full_size_properties_and_device_mode(pctx):
    """BASELINE config 4 size (16 384 problems): feasibility and optimality hold for every instance, duplicated
    instances give bit-identical answers, and the device-resident call equals the host-buffer call."""
    import jrl_walkgen_b200 as wg
    ctx, K = pctx
    _, small = W.pldp_batch(1024, seed=31, K=K)
    reps = 16
    pb = {k: (np.tile(v, (reps,) + (1,) * (v.ndim - 1)) if isinstance(v, np.ndarray) else v) for k, v in small.items()}
    B = 1024 * reps
    X, info = ctx.pldp_solve(pb)
    assert (info["rc"] == 0).all() and (info["status"] == 0).all()
    for r in range(1, reps):
        assert np.array_equal(X[r * 1024:(r + 1) * 1024], X[:1024])
    for b in range(0, 1024, 5):
        m = int(pb["m"][b])
        A = pb["DPu"][b, :(m + 1) * 32].reshape(32, m + 1).T[:m]
        s = A @ X[b] + pb["DPx"][b, :m]
        assert s.min() > -1e-7
        k = info["n_active"][b]
        g = X[b] + pb["D"][b]
        if k:
            E = A[info["active"][b][:k]]
            lam, *_ = np.linalg.lstsq(E.T, g, rcond=None)
            assert np.abs(E.T @ lam - g).max() < 1e-6
    # device-resident
    dev = {k: (ctx.to_device(v) if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    dX = ctx.alloc(B * 32 * 8); dinfo = ctx.alloc(B * wg.PLDP_INFO_DTYPE.itemsize)
    ctx.pldp_solve(dev, mem=wg.WG_MEM_DEVICE, B=B, X=dX, info=dinfo)
    ctx.sync()
    assert np.array_equal(dX.download(np.float64, (B, 32)), X)
    assert dinfo.download(wg.PLDP_INFO_DTYPE, (B,)).tobytes() == info.tobytes()
    for v in list(dev.values()) + [dX, dinfo]:
        if hasattr(v, "free"):
            v.free()
    # empty batch
    empty = {k: (v[:0] if isinstance(v, np.ndarray) else v) for k, v in small.items()}
    Xe, ie = ctx.pldp_solve(empty)
    assert len(Xe) == 0


@pytest.mark.gpu
def test_gpu_pldp_ranked_entry_is_bitwise_the_dense_entry(pctx):
    """wg_pldp_solve_batch_ranked (rows as (A_r(0), A_r(1), i_r), products formed on the fly) against wg_pldp_solve_batch
    on the materialised matrix: X, activation order, iteration counts and hot-start memory bit for bit - cold batch,
    device-resident call and a hot-started receding-horizon sequence."""
    import jrl_walkgen_b200 as wg
    ctx, K = pctx
    _, pb = W.pldp_batch(2048, seed=41, K=K)
    Xd, infod = ctx.pldp_solve(pb)
    Xr, infor = ctx.pldp_solve(pb, ranked=True)
    assert (infod["status"] == 0).all()
    assert np.array_equal(Xd, Xr) and infod.tobytes() == infor.tobytes()
    # device-resident ranked call
    B = 2048
    dev = {k: (ctx.to_device(v) if isinstance(v, np.ndarray) else v) for k, v in pb.items() if k != "DPu"}
    dX = ctx.alloc(B * 32 * 8); dinfo = ctx.alloc(B * wg.PLDP_INFO_DTYPE.itemsize)
    ctx.pldp_solve(dev, mem=wg.WG_MEM_DEVICE, B=B, X=dX, info=dinfo, ranked=True)
    ctx.sync()
    assert np.array_equal(dX.download(np.float64, (B, 32)), Xd)
    for v in list(dev.values()) + [dX, dinfo]:
        if hasattr(v, "free"):
            v.free()
    # hot-started sequence, both entries side by side
    Bh, T = 32, 12
    polys = [W._support_polygons(np.random.default_rng([43, b]), K.N, K.T, count=K.N + T) for b in range(Bh)]
    xk = np.zeros((Bh, 6))
    for b in range(Bh):
        xk[b, 0], xk[b, 3] = polys[b][0][0]
    hot_d = np.zeros(Bh, dtype=wg.PLDP_STATE_DTYPE); hot_r = hot_d.copy()
    n_removed = np.zeros(Bh, dtype=np.int32)
    for t in range(T):
        probs = [W.pldp_problem_from(K, polys[b][t:t + K.N], xk[b]) for b in range(Bh)]
        p2 = W.pldp_pack(K, probs)
        st = np.full(Bh, int(t == 0), dtype=np.int32)
        X1, i1 = ctx.pldp_solve(p2, hot=hot_d, starting=st, n_removed=n_removed)
        X2, i2 = ctx.pldp_solve(p2, hot=hot_r, starting=st, n_removed=n_removed, ranked=True)
        assert np.array_equal(X1, X2, equal_nan=True) and i1.tobytes() == i2.tobytes() and hot_d.tobytes() == hot_r.tobytes()
        for b in range(Bh):
            if i1["status"][b] == 0:
                xk[b] = W.pldp_advance(K, xk[b], X1[b])
        n_removed = np.array([p["n_first"] for p in probs], dtype=np.int32)


@pytest.mark.gpu
@pytest.mark.parametrize("ranked", [False, True])
def test_gpu_pldp_similar_constraints_semantics(pctx, ranked):
    """SimilarConstraints through the C ABI: flags that match the matrix are bit-neutral; backward flags that do NOT match
    it reproduce the reference's reuse of -tmp1 exactly (against the oracle port, itself bitwise the reference object on
    such flags, tests/test_pldp_oracle.py); forward / out-of-range flags are refused."""
    import jrl_walkgen_b200 as wg
    ctx, K = pctx
    _, pb = W.pldp_batch(96, seed=11, K=K)
    X0, i0 = ctx.pldp_solve(pb, ranked=ranked)
    X1, i1 = ctx.pldp_solve(pb, similar=pb["similar"], ranked=ranked)
    assert np.array_equal(X0, X1) and i0.tobytes() == i1.tobytes()
    rng = np.random.default_rng(5)
    bad = np.zeros((96, 128), dtype=np.int32)
    for b in range(96):
        for li in range(1, int(pb["m"][b])):
            if rng.random() < 0.3:
                bad[b, li] = -int(rng.integers(1, min(li, 5) + 1))
    X2, i2 = ctx.pldp_solve(pb, similar=bad, ranked=ranked)
    differ = 0
    for b in range(96):
        Xo, io, act = po.oracle_solve(K, pb, b, starting=True, similar=bad[b])
        assert i2["status"][b] == io[1] and i2["iterations"][b] == io[2], b
        assert np.array_equal(X2[b], Xo, equal_nan=True) and np.array_equal(i2["active"][b], act), b
        differ += int(not np.array_equal(X2[b], X0[b], equal_nan=True))
    assert differ >= 10
    fwd = np.zeros((96, 128), dtype=np.int32); fwd[7, 0] = 2
    with pytest.raises(wg.WalkgenError) as ei:
        ctx.pldp_solve(pb, similar=fwd, ranked=ranked)
    assert ei.value.code == -3
    # device-side batches are not validated on the host: the kernel reports status 6 for that instance only
    keys = [k for k in pb if not (ranked and k == "DPu")]
    dev = {k: (ctx.to_device(pb[k]) if isinstance(pb[k], np.ndarray) else pb[k]) for k in keys}
    dfwd = ctx.to_device(fwd)
    dX = ctx.alloc(96 * 32 * 8); dinfo = ctx.alloc(96 * wg.PLDP_INFO_DTYPE.itemsize)
    ctx.pldp_solve(dev, mem=wg.WG_MEM_DEVICE, B=96, X=dX, info=dinfo, similar=dfwd, similar_stride=128, ranked=ranked)
    ctx.sync()
    info = dinfo.download(wg.PLDP_INFO_DTYPE, (96,))
    assert info["status"][7] == 6 and (np.delete(info["status"], 7) == 0).all()
    for v in list(dev.values()) + [dX, dinfo, dfwd]:
        if hasattr(v, "free"):
            v.free()


@pytest.mark.gpu
def test_gpu_pldp_refuses_out_of_range_problem_sizes(pctx):
    """m > WG_PLDP_MAX_ROWS (= 128 = 8 rows x 16 samples), m < 0 and strides shorter than the problem: WG_ERR_INVALID for
    host batches; for device batches the kernel writes status 7 / NaN for the offending instance and solves the others."""
    import jrl_walkgen_b200 as wg
    ctx, K = pctx
    _, pb = W.pldp_batch(8, seed=3, K=K)
    for bad_m in (129, -1):
        p2 = dict(pb); p2["m"] = pb["m"].copy(); p2["m"][3] = bad_m
        with pytest.raises(wg.WalkgenError) as ei:
            ctx.pldp_solve(p2)
        assert ei.value.code == -3
    p3 = dict(pb); p3["dpu_stride"] = 32 * int(pb["m"].max())        # one row short of (m+1)*32
    with pytest.raises(wg.WalkgenError):
        ctx.pldp_solve(p3)
    p4 = dict(pb); p4["dpx_stride"] = int(pb["m"].max()) - 1
    with pytest.raises(wg.WalkgenError):
        ctx.pldp_solve(p4)
    X0, i0 = ctx.pldp_solve(pb)
    m_bad = pb["m"].copy(); m_bad[3] = 200
    dev = {k: (ctx.to_device(m_bad if k == "m" else v) if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
    dX = ctx.alloc(8 * 32 * 8); dinfo = ctx.alloc(8 * wg.PLDP_INFO_DTYPE.itemsize)
    ctx.pldp_solve(dev, mem=wg.WG_MEM_DEVICE, B=8, X=dX, info=dinfo)
    ctx.sync()
    X = dX.download(np.float64, (8, 32)); info = dinfo.download(wg.PLDP_INFO_DTYPE, (8,))
    assert info["status"][3] == 7 and info["rc"][3] == -1 and np.isnan(X[3]).all()
    keep = np.arange(8) != 3
    assert np.array_equal(X[keep], X0[keep]) and (info["status"][keep] == 0).all()
    for v in list(dev.values()) + [dX, dinfo]:
        if hasattr(v, "free"):
            v.free()
