"""Kajita2003 preview control: oracle pins (CPU) and CUDA-vs-oracle parity (GPU).

Reference: PreviewControl::ComputeOptimalWeights / OneIterationOfPreview
(src/PreviewControl/PreviewControl.cpp:198-374), OptimalControllerSolver::ComputeWeights
(src/PreviewControl/OptimalControllerSolver.cpp:200-352).
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

# Gains shipped by the reference in src/data/PreviewControlParameters.ini (zc 0.814, T 5 ms, 1.6 s),
# 5 significant digits: a sanity pin at ~1e-4 relative.
INI_KX = (72719.0, 21550.0, 177.01)
INI_KS = 618.7
INI_F012 = (618.7, 777.51, 952.26)

TOL_COM = 1e-9   # m; north_star asks for 1e-6 m - we test three orders tighter


def synth_walk(rng, L):
    """Piecewise-constant ZMP reference resembling a footstep sequence."""
    z = np.zeros((L, 2))
    k = 0
    x = 0.0
    side = 1.0
    while k < L:
        n = int(rng.integers(120, 200))
        z[k:k + n, 0] = x
        z[k:k + n, 1] = side * 0.095
        x += rng.uniform(0.05, 0.25)
        side = -side
        k += n
    return z


def test_oracle_gains_match_reference_ini():
    g = ol.OracleGains(0.005, 1.6, 0.814, 1)
    assert g.NL == 320
    assert np.allclose(g.Kx, INI_KX, rtol=2e-4)
    assert abs(g.Ks - INI_KS) / INI_KS < 2e-4
    assert np.allclose(g.F[:3], INI_F012, rtol=2e-4)


@pytest.mark.parametrize("T,Tp,zc,mode,R", [(0.005, 1.6, 0.814, 1, 1e-6), (0.005, 1.6, 0.807709, 1, 1e-6),
                                              (0.01, 1.6, 0.814, 0, 1e-5)])
def test_oracle_gains_match_scipy_dare(T, Tp, zc, mode, R):
    """TestRiccatiEquation's two configurations (tests/TestRiccatiEquation.cpp) against an independent
    DARE solver."""
    from scipy.linalg import solve_discrete_are
    g = ol.OracleGains(T, Tp, zc, mode)
    A = np.array([[1, T, T * T / 2], [0, 1, T], [0, 0, 1.0]])
    B = np.array([[T ** 3 / 6], [T * T / 2], [T]])
    Cm = np.array([[1, 0, -zc / 9.81]])
    if mode == 1:
        Ax = np.zeros((4, 4)); Ax[0, 0] = 1; Ax[0, 1:] = Cm @ A; Ax[1:, 1:] = A
        bx = np.vstack([Cm @ B, B]); cx = np.array([[1, 0, 0, 0.0]])
    else:
        Ax, bx, cx = A, B, Cm
    P = solve_discrete_are(Ax, bx, cx.T @ cx, np.array([[R]]))
    la = 1.0 / (R + bx.T @ P @ bx)
    K = (la * bx.T @ P @ Ax).ravel()
    rec = P @ cx.T if mode == 1 else cx.T.copy()
    Ac = (Ax - bx @ K[None, :]).T
    F = []
    for _ in range(g.NL):
        F.append((la * bx.T @ rec).item())
        rec = Ac @ rec
    F = np.array(F)
    assert np.abs(F - g.F).max() <= 1e-9 * np.abs(F).max()
    assert abs(K[0] - g.Ks) <= 1e-9 * abs(K[0])
    kx = K[1:] if mode == 1 else K
    assert np.allclose(kx, g.Kx, rtol=1e-9)


def test_product_gains_match_oracle():
    """wg_preview_gains is host code (SDA Riccati) - runs without a GPU."""
    import jrl_walkgen_b200 as wg
    for (T, Tp, zc, mode) in [(0.005, 1.6, 0.814, 1), (0.005, 1.6, 0.807709, 1), (0.01, 1.6, 0.814, 0),
                              (0.005, 0.8, 0.75, 1)]:
        g = wg.preview_gains(T, Tp, zc, mode)
        o = ol.OracleGains(T, Tp, zc, mode)
        assert g.NL == o.NL
        F = np.array(g.F[:g.NL])
        assert np.abs(F - o.F).max() <= 1e-9 * np.abs(o.F).max()
        assert abs(g.Ks - o.Ks) <= 1e-9 * abs(o.Ks)
        assert np.allclose(np.array(g.Kx[:]), o.Kx, rtol=1e-9)
        assert np.allclose(np.array(g.A[:]), o.A) and np.allclose(np.array(g.B[:]), o.B)


def test_oracle_step_equals_run():
    """oracle_preview_run is oracle_preview_step iterated (FIFO popped once per tick)."""
    g = ol.OracleGains()
    rng = np.random.default_rng(0)
    z = synth_walk(rng, 400)
    st = np.zeros(8)
    com, zmp, steps = ol.oracle_preview_batch(g, [0, 400], z, st)
    assert steps == 81
    x = np.zeros(3); y = np.zeros(3); sx = C.c_double(0); sy = C.c_double(0)
    zx = C.c_double(); zy = C.c_double()
    for k in range(81):
        rc = ol.oracle().oracle_preview_step(ol.dptr(g.A), ol.dptr(g.B), ol.dptr(g.C), ol.dptr(g.Kx), g.Ks,
                                             ol.dptr(g.F), g.NL, ol.dptr(x), ol.dptr(y), C.byref(sx), C.byref(sy),
                                             ol.dptr(z[k:]), 400 - k, C.byref(zx), C.byref(zy), 1)
        assert rc == 0
        assert np.array_equal(com[k, :3], x) and np.array_equal(com[k, 3:], y)
        assert zmp[k, 0] == zx.value and zmp[k, 1] == zy.value
    # window under-filled -> error, as the reference LTHROWs (PreviewControl.cpp:341-344)
    rc = ol.oracle().oracle_preview_step(ol.dptr(g.A), ol.dptr(g.B), ol.dptr(g.C), ol.dptr(g.Kx), g.Ks,
                                         ol.dptr(g.F), g.NL, ol.dptr(x), ol.dptr(y), C.byref(sx), C.byref(sy),
                                         ol.dptr(z), 319, C.byref(zx), C.byref(zy), 1)
    assert rc == -1


def test_oracle_tracks_constant_reference():
    """Property: with a constant ZMP reference the controller converges to CoM = ZMP = reference."""
    g = ol.OracleGains()
    z = np.tile(np.array([[0.3, -0.1]]), (4000, 1))
    st = np.zeros(8)
    com, zmp, steps = ol.oracle_preview_batch(g, [0, 4000], z, st)
    assert abs(com[steps - 1, 0] - 0.3) < 1e-6 and abs(com[steps - 1, 3] + 0.1) < 1e-6
    assert abs(zmp[steps - 1, 0] - 0.3) < 1e-6 and abs(zmp[steps - 1, 1] + 0.1) < 1e-6


# ------------------------------------------------------------------------------------------------
# Recursive evaluation of the preview sum (preview_rec_kernel): host logic
# ------------------------------------------------------------------------------------------------
def test_window_weights_are_matrix_geometric_and_tables_that_are_not_fall_back():
    """OptimalControllerSolver.cpp:323-345 builds F[i] = la b' ((A - bK)')^i P c'Q: wg_preview_sum_fit recovers that structure
    from a gain table to rounding (both modes, several T / preview times / heights), and refuses tables without it: one tap
    off by 1e-9 relative, or the 5-digit precision of the reference's PreviewControlParameters.ini."""
    import jrl_walkgen_b200 as wg
    from jrl_walkgen_b200 import _capi
    lib = _capi.load()
    tol = 1e-12                                            # WG_PREVIEW_REC_TOL
    for mode in (wg.MODE_WITHOUT_INITIALPOS, wg.MODE_WITH_INITIALPOS):
        for T, Tp, zc in [(0.005, 1.6, 0.814), (0.005, 0.8, 0.6), (0.01, 1.6, 1.0), (0.002, 2.0, 0.814), (0.005, 0.04, 0.814)]:
            g = wg.preview_gains(T, Tp, zc, mode)
            r = lib.wg_preview_sum_fit(C.byref(g))
            assert 0.0 <= r < 2e-14, (mode, T, Tp, zc, r)
    o = ol.OracleGains(0.005, 1.6, 0.814, 1)               # another solver's weights (the oracle's Riccati fixed point)
    g = wg.preview_gains()
    for i in range(3):
        g.Kx[i] = o.Kx[i]
    g.Ks = o.Ks
    for i in range(g.NL):
        g.F[i] = o.F[i]
    assert 0.0 <= lib.wg_preview_sum_fit(C.byref(g)) < 2e-14
    g = wg.preview_gains()
    g.F[10] *= 1 + 1e-9
    assert lib.wg_preview_sum_fit(C.byref(g)) > tol
    g = wg.preview_gains()
    for i in range(g.NL):
        g.F[i] = float(f"{g.F[i]:.5g}")
    assert lib.wg_preview_sum_fit(C.byref(g)) > 1e-7


@pytest.mark.parametrize("threads,NLcfg", [(64, (0.005, 1.6)), (128, (0.005, 1.6)), (32, (0.005, 0.8)), (64, (0.005, 1.595))])
def test_recursive_sum_constants_reproduce_the_direct_sum(threads, NLcfg):
    """The decomposition preview_rec_kernel runs per tile - W at the tile's end from the table E, the 8-tap triangular sums of a
    thread, its local total, the scan DOWN the tile with L^(8d), the per-lane power of the state entering a warp, the
    correction w' L^(8-r) W_in - executed here in numpy with the library's own constants (wgi_preview_rec_dump), thread by
    thread, against the direct 320-tap sum in extended precision: 1e-13 relative."""
    import jrl_walkgen_b200 as wg
    from jrl_walkgen_b200 import _capi
    lib = _capi.load()
    lib.wgi_preview_rec_dump.restype = C.c_longlong
    lib.wgi_preview_rec_dump.argtypes = [C.POINTER(_capi.PreviewGains), C.c_void_p, C.c_longlong]
    g = wg.preview_gains(NLcfg[0], NLcfg[1], 0.814, 1)
    NL = g.NL; R = 8; NLpad = (NL + 7) // 8 * 8
    n = lib.wgi_preview_rec_dump(C.byref(g), None, 0)
    assert n == 8 + 8 + 32 + 32 + 36 + 96 + NLpad * 4 + 512
    flat = np.zeros(n)
    assert lib.wgi_preview_rec_dump(C.byref(g), flat.ctypes.data, n) == n
    o = 0
    def take(k, shape):
        nonlocal o
        a = flat[o:o + k].reshape(shape); o += k
        return a
    RF0 = take(8, (8,)); RFN = take(8, (8,)); RV = take(32, (8, 4)); RVN = take(32, (8, 4)); RW = take(36, (9, 4))
    RP = take(96, (6, 4, 4)); E = take(NLpad * 4, (NLpad, 4)); LP = take(512, (32, 4, 4))
    F = np.array(g.F[:NL])
    assert np.allclose(RF0, F[:8], rtol=1e-13) and np.array_equal(LP[0], np.eye(4)) and np.allclose(LP[1], RP[0], rtol=1e-15)
    assert (E[NL:] == 0).all()
    rng = np.random.default_rng(5)
    TILE = R * threads
    L = 2 * TILE + NL + 77                                  # three tiles, the last one ragged
    p = synth_walk(rng, L)[:, 0] + rng.normal(scale=1e-3, size=L)
    nsteps = L - NL + 1
    fex = np.array([float((F.astype(np.longdouble) * p[k:k + NL].astype(np.longdouble)).sum()) for k in range(nsteps)])
    f = np.zeros(nsteps)
    for start in range(0, nsteps, TILE):
        span = TILE + NLpad
        sp = np.zeros(span); m = min(span, L - start); sp[:m] = p[start:start + m]
        Wend = (E[:NL] * sp[TILE:TILE + NL, None]).sum(axis=0)
        floc = np.zeros((threads, R)); c = np.zeros((threads, 4))
        for t in range(threads):
            own = sp[R * t:R * t + R]; far = sp[R * t + NL:R * t + NL + R]
            for j in range(R):
                for r in range(j + 1):
                    floc[t, r] += RF0[j - r] * own[j] + RFN[j - r] * far[j]
                c[t] += RV[j] * own[j] + RVN[j] * far[j]
        for l in range(5):
            d = 1 << l; new = c.copy()
            for t in range(threads):
                if (t & 31) + d < 32:
                    new[t] = c[t] + RP[l] @ c[t + d]
            c = new
        nw = threads // 32
        V = [None] * nw; V[nw - 1] = Wend
        for w in range(nw - 2, -1, -1):
            V[w] = c[32 * (w + 1)] + RP[5] @ V[w + 1]
        for t in range(threads):
            lane = t & 31
            Win = (c[t + 1] if lane < 31 else np.zeros(4)) + LP[31 - lane] @ V[t >> 5]
            for r in range(R):
                k = start + R * t + r
                if k < nsteps:
                    f[k] = floc[t, r] + RW[R - r] @ Win
    assert np.abs(f - fex).max() < 1e-13 * np.abs(fex).max(), np.abs(f - fex).max() / np.abs(fex).max()


# ------------------------------------------------------------------------------------------------
# GPU parity
# ------------------------------------------------------------------------------------------------
def _run_gpu(ctx, gains, offsets, z, st0, simulation=True, mem_device=False):
    import jrl_walkgen_b200 as wg
    ctx.preview_set_gains(gains)
    plan = ctx.preview_plan(offsets)
    n = int(offsets[-1])
    st = st0.copy()
    com = np.full((max(n, 1), 6), np.nan)
    zmp = np.full((max(n, 1), 2), np.nan)
    if mem_device:
        dz = ctx.to_device(z if n else np.zeros(2)); ds = ctx.to_device(st)
        dc = ctx.to_device(com); dzo = ctx.to_device(zmp)
        plan.run(dz, ds, dc, dzo, simulation, mem=wg.WG_MEM_DEVICE)
        ctx.sync()
        st = ds.download(np.float64, st.shape); com = dc.download(np.float64, com.shape)
        zmp = dzo.download(np.float64, zmp.shape)
        for b in (dz, ds, dc, dzo):
            b.free()
    else:
        plan.run(z, st, com, zmp, simulation)
    steps = plan.total_steps
    plan.destroy()
    return com, zmp, st, steps


def _valid_rows(offsets, NL):
    rows = []
    for b in range(len(offsets) - 1):
        L = offsets[b + 1] - offsets[b]
        rows.extend(range(offsets[b], offsets[b] + max(0, L - NL + 1)))
    return np.array(rows, dtype=np.int64)


@pytest.mark.gpu
@pytest.mark.parametrize("mem_device", [False, True])
def test_gpu_preview_matches_oracle_ragged(ctx, mem_device):
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(1)
    gains = wg.preview_gains(0.005, 1.6, 0.807709, 1)
    og = ol.OracleGains(0.005, 1.6, 0.807709, 1)
    # ragged lengths incl. edge cases: shorter than the window (0 steps), exactly NL (1 step),
    # tile boundaries of the FIR kernel (1024 outputs per block)
    lens = [3362 + 640, 319, 320, 321, 320 + 1023, 320 + 1024, 320 + 1025, 0, 2500, 5000, 640]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    z = np.concatenate([synth_walk(rng, L) for L in lens if L > 0])
    st0 = rng.normal(scale=0.01, size=(len(lens), 8))
    com, zmp, st, steps = _run_gpu(ctx, gains, offsets, z, st0, True, mem_device)
    st_o = st0.copy()
    com_o, zmp_o, steps_o = ol.oracle_preview_batch(og, offsets, z, st_o)
    assert steps == steps_o
    rows = _valid_rows(offsets, 320)
    assert np.abs(com[rows] - com_o[rows])[:, [0, 3]].max() < TOL_COM
    assert np.abs(zmp[rows] - zmp_o[rows]).max() < 1e-8
    assert np.allclose(st[:, :6], st_o[:, :6], atol=1e-8)
    assert np.allclose(st, st_o, rtol=1e-7, atol=1e-8)
    # rows past a trajectory's last step are untouched
    mask = np.ones(len(com), bool); mask[rows] = False
    assert np.isnan(com[mask]).all() if mem_device else (com[mask] == 0).all()


@pytest.mark.gpu
def test_gpu_preview_no_simulation_flag(ctx):
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(2)
    gains = wg.preview_gains(); og = ol.OracleGains()
    offsets = np.array([0, 900, 2000], dtype=np.int64)
    z = np.concatenate([synth_walk(rng, 900), synth_walk(rng, 1100)])
    st0 = rng.normal(scale=0.01, size=(2, 8))
    com, zmp, st, _ = _run_gpu(ctx, gains, offsets, z, st0, simulation=False)
    st_o = st0.copy()
    com_o, zmp_o, _ = ol.oracle_preview_batch(og, offsets, z, st_o, simulation=False)
    rows = _valid_rows(offsets, 320)
    assert np.abs(com[rows] - com_o[rows]).max() < 1e-7
    assert np.array_equal(st[:, 6:], st0[:, 6:])  # integrators untouched when Simulation=false


@pytest.mark.gpu
def test_gpu_preview_one_iteration_and_window_error(ctx):
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(3)
    gains = wg.preview_gains(); og = ol.OracleGains()
    ctx.preview_set_gains(gains)
    z = synth_walk(rng, 320)
    x0 = rng.normal(scale=0.01, size=3); y0 = rng.normal(scale=0.01, size=3)
    x, y, sx, sy, zx, zy = ctx.preview_one_iteration(x0, y0, 0.01, -0.02, z)
    st = np.concatenate([x0, y0, [0.01, -0.02]])
    com_o, zmp_o, _ = ol.oracle_preview_batch(og, [0, 320], z, st)
    assert np.allclose(x, st[:3], atol=1e-10) and np.allclose(y, st[3:6], atol=1e-10)
    assert abs(sx - st[6]) < 1e-9 and abs(sy - st[7]) < 1e-9
    assert abs(zx - zmp_o[0, 0]) < 1e-9 and abs(zy - zmp_o[0, 1]) < 1e-9
    with pytest.raises(wg.WalkgenError) as ei:
        ctx.preview_one_iteration(x0, y0, 0.0, 0.0, z[:319])
    assert ei.value.code == -6  # WG_ERR_WINDOW <-> LTHROW at PreviewControl.cpp:341-344


@pytest.mark.gpu
def test_gpu_preview_linearity_full_size(ctx):
    """Size-independent property at BASELINE config-2 scale (4096 walks): the map
    (zmpref, state0) -> outputs is linear, so run(a)+run(b) == run(a+b)."""
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(4)
    gains = wg.preview_gains()
    B = 4096
    lens = rng.integers(2800, 4600, size=B)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(offsets[-1])
    za = rng.normal(scale=0.1, size=(n, 2)); zb = rng.normal(scale=0.1, size=(n, 2))
    sa = rng.normal(scale=0.01, size=(B, 8)); sb = rng.normal(scale=0.01, size=(B, 8))
    ca, _, fa, _ = _run_gpu(ctx, gains, offsets, za, sa)
    cb, _, fb, _ = _run_gpu(ctx, gains, offsets, zb, sb)
    cs, _, fs, _ = _run_gpu(ctx, gains, offsets, za + zb, sa + sb)
    rows = _valid_rows(offsets, 320)[::97]
    assert np.abs(ca[rows] + cb[rows] - cs[rows])[:, [0, 3]].max() < 1e-9
    assert np.abs(fa + fb - fs)[:, :6].max() < 1e-8


@pytest.mark.gpu
def test_gpu_preview_full_size_device_resident_vs_oracle_and_host_path(ctx):
    """BASELINE config-2 scale through the kernel the bench times: 4096 ragged walks in ONE device-resident launch (the launch
    heuristic picks preview_rec_warp_kernel from 1792 trajectories up) against (i) the oracle on 32 of the walks, every valid row,
    and (ii) the host-buffer path on ALL walks (16 chunks of 256 walks: preview_rec_kernel at 128 x 4) - two kernels, two store
    paths (bulk-copy engine / shared-memory staging), same numbers."""
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(8)
    gains = wg.preview_gains()
    og = ol.OracleGains(0.005, 1.6, 0.814, 1)
    B = 4096
    lens = rng.integers(2800, 4600, size=B)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(offsets[-1])
    z = np.cumsum(rng.normal(scale=2e-3, size=(n, 2)), axis=0)
    for b in range(B):                               # every walk starts at the origin
        z[offsets[b]:offsets[b + 1]] -= z[offsets[b]]
    st0 = rng.normal(scale=0.01, size=(B, 8))
    com_d, zmp_d, st_d, steps = _run_gpu(ctx, gains, offsets, z, st0, True, mem_device=True)
    com_h, zmp_h, st_h, _ = _run_gpu(ctx, gains, offsets, z, st0, True, mem_device=False)
    rows = _valid_rows(offsets, 320)
    assert steps == len(rows)
    assert np.abs(com_d[rows] - com_h[rows])[:, [0, 3]].max() < 1e-11
    assert np.abs(com_d[rows] - com_h[rows]).max() < 1e-8 and np.abs(zmp_d[rows] - zmp_h[rows]).max() < 1e-9
    assert np.abs(st_d - st_h).max() < 1e-9
    pick = rng.choice(B, size=32, replace=False)
    for b in pick:
        o, L = int(offsets[b]), int(lens[b])
        so = st0[b:b + 1].copy()
        com_o, zmp_o, _ = ol.oracle_preview_batch(og, np.array([0, L], dtype=np.int64), z[o:o + L], so)
        k = L - 320 + 1
        assert np.abs(com_d[o:o + k] - com_o[:k])[:, [0, 3]].max() < TOL_COM, b
        assert np.abs(zmp_d[o:o + k] - zmp_o[:k]).max() < 1e-8, b
        assert np.allclose(st_d[b], so[0], rtol=1e-7, atol=1e-8), b


@pytest.mark.gpu
def test_gpu_batched_gains_match_host_solve_and_oracle(ctx):
    """wg_preview_gains_batch (OptimalControllerSolver::ComputeWeights for per-instance (T, preview time, zc),
    one thread per parameter set) against the host solve of wg_preview_gains on every set, and against the oracle's Riccati
    fixed point (an independent algorithm) on a sample: Ks, Kx and all NL window weights at 1e-9 relative.  Both modes;
    a refused set (T <= 0) reports NL = 0 / NaN without disturbing its neighbours."""
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(17)
    B = 513
    par = np.column_stack([rng.choice([0.005, 0.01, 0.002], B), rng.uniform(0.8, 2.0, B), rng.uniform(0.6, 1.0, B)])
    par[0] = (0.005, 1.6, 0.814)
    for mode in (wg.MODE_WITHOUT_INITIALPOS, wg.MODE_WITH_INITIALPOS):
        p = par.copy()
        p[7, 0] = -1.0
        heads, F = ctx.preview_gains_batch(p, mode)
        assert heads["NL"][7] == 0 and np.isnan(heads["Ks"][7])
        worst = 0.0
        for b in range(B):
            if b == 7:
                continue
            g = wg.preview_gains(p[b, 0], p[b, 1], p[b, 2], mode)
            assert heads["NL"][b] == g.NL
            Fh = np.array(g.F[:g.NL])
            e = max(abs(heads["Ks"][b] - g.Ks) / abs(g.Ks), np.abs(heads["Kx"][b] - np.array(g.Kx[:])).max() / np.abs(g.Kx[:]).max(),
                    np.abs(F[b, :g.NL] - Fh).max() / np.abs(Fh).max())
            worst = max(worst, e)
            assert np.array_equal(heads["A"][b], np.array(g.A[:])) and np.array_equal(heads["C"][b], np.array(g.C[:]))
            assert (F[b, g.NL:] == 0).all()
        assert worst < 1e-9, worst
        for b in (0, 1, 2, 100):
            og = ol.OracleGains(p[b, 0], p[b, 1], p[b, 2], mode)
            assert abs(heads["Ks"][b] - og.Ks) < 1e-8 * abs(og.Ks)
            assert np.abs(F[b, :og.NL] - og.F).max() < 1e-8 * np.abs(og.F).max()
        print(f"batched gains mode {mode}: max rel deviation from the host solve {worst:.2e}")


@pytest.mark.gpu
def test_gpu_position_only_output_equals_the_full_run(ctx):
    """wg_preview_run_batch_pos (output selection: 16 B per step instead of 64) writes exactly the x / y of the full run, host
    and device memory, with and without the integrated error."""
    import jrl_walkgen_b200 as wg
    gains = wg.preview_gains(0.005, 1.6, 0.814, wg.MODE_WITHOUT_INITIALPOS)
    ctx.preview_set_gains(gains)
    rng = np.random.default_rng(31)
    lens = [1500, 330, 4000, 320, 900]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(offsets[-1])
    z = np.cumsum(rng.normal(scale=1e-3, size=(n, 2)), axis=0)
    plan = ctx.preview_plan(offsets)
    for sim in (True, False):
        st_a = rng.normal(scale=0.01, size=(len(lens), 8)); st_b = st_a.copy(); st_c = st_a.copy()
        com = np.zeros((n, 6)); zmp = np.zeros((n, 2)); pos = np.zeros((n, 2))
        plan.run(z, st_a, com, zmp, sim)
        plan.run_pos(z, st_b, pos, sim)
        assert np.array_equal(pos, com[:, [0, 3]]) and np.array_equal(st_a, st_b)
        dz = ctx.to_device(z); ds = ctx.to_device(st_c); dp = ctx.alloc(pos.nbytes)
        plan.run_pos(dz, ds, dp, sim, mem=wg.WG_MEM_DEVICE)
        ctx.sync()
        got = dp.download(np.float64, (n, 2))
        for b, L in enumerate(lens):
            o = int(offsets[b]); steps = L - 320 + 1
            assert np.array_equal(got[o:o + steps], pos[o:o + steps])
        assert np.array_equal(ds.download(np.float64, (len(lens), 8)), st_a)
        for d in (dz, ds, dp):
            d.free()
    plan.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(0.005, 1.6, 0.807709, 1), (0.01, 1.6, 0.814, 0), (0.005, 0.8, 0.75, 1), (0.005, 1.595, 0.814, 1),
                                 (0.002, 2.0, 0.814, 1), (0.005, 0.04, 0.814, 1), (0.001, 2.048, 0.814, 1)])
def test_gpu_recursive_and_direct_preview_sums_agree_with_the_oracle(ctx, cfg):
    """Both evaluations of the preview sum - preview_rec_kernel (the default for weights of the reference's structure) and the
    direct 320-tap sum of preview_fused_kernel - against the oracle on a ragged batch (tile boundaries of every CTA shape,
    windows that are not a multiple of 8, windows of 8, 1000 and WG_PREVIEW_MAX_NL = 2048 samples, both solver modes, with and
    without the integrated error), and against each other."""
    import jrl_walkgen_b200 as wg
    T, Tp, zc, mode = cfg
    rng = np.random.default_rng(11)
    gains = wg.preview_gains(T, Tp, zc, mode)
    og = ol.OracleGains(T, Tp, zc, mode)
    NL = gains.NL
    lens = [NL + 4000, NL - 1, NL, NL + 1, NL + 511, NL + 512, NL + 513, NL + 1023, NL + 1024, NL + 1025, 0, 2500, 5000, 2 * NL,
            NL + 255, NL + 256, NL + 257]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    z = np.concatenate([synth_walk(rng, L) for L in lens if L > 0])
    st0 = rng.normal(scale=0.01, size=(len(lens), 8))
    rows = _valid_rows(offsets, NL)
    out = {}
    try:
        # MODE_WITH_INITIALPOS has no error integrator in its design (Ks is K(0,0) again): with Simulation = true the loop of
        # the reference itself diverges (the oracle reaches 1e145 within 1000 ticks), so that mode runs without it
        for sim in ((True, False) if mode == 1 else (False,)):
            st_o = st0.copy()
            com_o, zmp_o, _ = ol.oracle_preview_batch(og, offsets, z, st_o, simulation=sim)
            for name, m, shape in (("recursive", wg.PREVIEW_SUM_RECURSIVE, -1), ("recursive 64x8", wg.PREVIEW_SUM_RECURSIVE, 0),
                                   ("recursive 128x4", wg.PREVIEW_SUM_RECURSIVE, 1), ("recursive one warp", wg.PREVIEW_SUM_RECURSIVE, 2),
                                   ("recursive 256x2", wg.PREVIEW_SUM_RECURSIVE, 3),
                                   ("direct", wg.PREVIEW_SUM_DIRECT, -1)):
                ctx.preview_set_gains(gains)
                ctx.preview_set_sum_mode(m)
                ctx.preview_set_cta_shape(shape)
                used, resid = ctx.preview_sum_info()
                assert used == m and 0 <= resid < 1e-13
                com, zmp, st, _ = _run_gpu(ctx, gains, offsets, z, st0, sim, mem_device=name.startswith("recursive"))
                out[name] = (com, zmp, st)
                assert np.abs(com[rows] - com_o[rows])[:, [0, 3]].max() < TOL_COM, (name, sim)
                assert np.abs(com[rows] - com_o[rows]).max() < 1e-7, (name, sim)
                assert np.abs(zmp[rows] - zmp_o[rows]).max() < 1e-8, (name, sim)
                assert np.allclose(st, st_o, rtol=1e-7, atol=1e-8), (name, sim)
            for name in ("recursive 64x8", "recursive 128x4", "recursive one warp", "recursive 256x2"):
                assert np.abs(out[name][0][rows] - out["direct"][0][rows])[:, [0, 3]].max() < 1e-11, name
            d = np.abs(out["recursive"][0][rows] - out["direct"][0][rows])
            print(f"cfg {cfg} sim {sim}: recursive vs direct sum, max |dCoM| = {d[:, [0, 3]].max():.2e} m, "
                  f"|d ddCoM| = {d[:, [2, 5]].max():.2e} m/s^2")
            assert d[:, [0, 3]].max() < 1e-11
    finally:
        ctx.preview_set_sum_mode(wg.PREVIEW_SUM_AUTO)
        ctx.preview_set_cta_shape(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["recursive", "direct"])
def test_gpu_two_contexts_on_one_device_keep_their_own_gains(ctx, mode):
    """Gains belong to a context, the kernels' constant block to the device: two contexts on cuda:0 with different gain sets
    (different CoM height AND different window length), run alternately on plans created once, must each reproduce the oracle
    with THEIR gains - the block is re-bound before a launch whenever it holds the other context's image."""
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(5)
    cfgs = [(0.005, 1.6, 0.814, 1), (0.005, 1.2, 0.60, 1)]
    lens = [2100, 1400, 900]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    z = np.concatenate([synth_walk(rng, L) for L in lens])
    st0 = rng.normal(scale=0.01, size=(len(lens), 8))
    other = wg.Context(0)
    try:
        ctxs = [ctx, other]
        plans, refs = [], []
        for c, cfg in zip(ctxs, cfgs):
            c.preview_set_gains(wg.preview_gains(*cfg))
            c.preview_set_sum_mode(wg.PREVIEW_SUM_RECURSIVE if mode == "recursive" else wg.PREVIEW_SUM_DIRECT)
            plans.append(c.preview_plan(offsets))
            st_o = st0.copy()
            com_o, zmp_o, _ = ol.oracle_preview_batch(ol.OracleGains(*cfg), offsets, z, st_o)
            refs.append((com_o, zmp_o, st_o, _valid_rows(offsets, c.gains.NL)))
        assert ctxs[0].gains.NL != ctxs[1].gains.NL
        for _ in range(3):                       # A, B, A, B, ...: every launch finds the other context's image bound
            for k in (0, 1):
                st = st0.copy()
                com = np.zeros((len(z), 6)); zmp = np.zeros((len(z), 2))
                plans[k].run(z, st, com, zmp, True)
                com_o, zmp_o, st_o, rows = refs[k]
                assert np.abs(com[rows] - com_o[rows])[:, [0, 3]].max() < TOL_COM, (mode, k)
                assert np.abs(zmp[rows] - zmp_o[rows]).max() < 1e-8, (mode, k)
                assert np.allclose(st, st_o, rtol=1e-7, atol=1e-8), (mode, k)
        for p in plans:
            p.destroy()
    finally:
        other.close()
        ctx.preview_set_sum_mode(wg.PREVIEW_SUM_AUTO)


@pytest.mark.gpu
def test_gpu_preview_refuses_device_arrays_that_are_not_16_byte_aligned(ctx):
    """The batch kernels move rows with 128-bit accesses and bulk copies: a device array at an odd multiple of 8 bytes is refused
    (WG_ERR_INVALID) instead of faulting on the device."""
    import ctypes as C
    import jrl_walkgen_b200 as wg
    from jrl_walkgen_b200 import _capi
    gains = wg.preview_gains(0.005, 1.6, 0.814, 1)
    ctx.preview_set_gains(gains)
    offsets = np.array([0, 1000], dtype=np.int64)
    plan = ctx.preview_plan(offsets)
    dz = ctx.to_device(np.zeros((1001, 2))); ds = ctx.to_device(np.zeros((1, 8)))
    dc = ctx.to_device(np.zeros((1001, 6))); dzo = ctx.to_device(np.zeros((1001, 2)))
    try:
        for which in range(3):
            ptrs = [dz.ptr, dc.ptr, dzo.ptr]
            ptrs[which] += 8
            rc = ctx.lib.wg_preview_run_batch(ctx.h, plan.h, wg.WG_MEM_DEVICE, C.c_void_p(ptrs[0]), C.c_void_p(ds.ptr),
                                              C.c_void_p(ptrs[1]), C.c_void_p(ptrs[2]), 1)
            assert rc == _capi.WG_ERR_INVALID, (which, rc)
        rc = ctx.lib.wg_preview_run_batch(ctx.h, plan.h, wg.WG_MEM_DEVICE, C.c_void_p(dz.ptr), C.c_void_p(ds.ptr),
                                          C.c_void_p(dc.ptr), C.c_void_p(dzo.ptr), 1)
        assert rc == 0
        ctx.sync()
    finally:
        for d in (dz, ds, dc, dzo):
            d.free()
        plan.destroy()


@pytest.mark.gpu
def test_gpu_weights_without_the_structure_keep_the_direct_sum(ctx):
    """A gain table that is not matrix-geometric (here: rounded to the 5 digits of the reference's PreviewControlParameters.ini)
    runs through the direct sum under AUTO, RECURSIVE is refused, and the result is the oracle's for THAT table."""
    import jrl_walkgen_b200 as wg
    gains = wg.preview_gains()
    og = ol.OracleGains()
    for i in range(gains.NL):
        gains.F[i] = float(f"{gains.F[i]:.5g}")
        og.F[i] = gains.F[i]
    rng = np.random.default_rng(12)
    offsets = np.array([0, 1700, 1700 + 900], dtype=np.int64)
    z = np.concatenate([synth_walk(rng, 1700), synth_walk(rng, 900)])
    st0 = np.zeros((2, 8))
    try:
        ctx.preview_set_gains(gains)
        used, resid = ctx.preview_sum_info()
        assert used == wg.PREVIEW_SUM_DIRECT and resid > 1e-7
        with pytest.raises(wg.WalkgenError):
            ctx.preview_set_sum_mode(wg.PREVIEW_SUM_RECURSIVE)
        com, zmp, st, _ = _run_gpu(ctx, gains, offsets, z, st0)
        st_o = st0.copy()
        com_o, zmp_o, _ = ol.oracle_preview_batch(og, offsets, z, st_o)
        rows = _valid_rows(offsets, 320)
        assert np.abs(com[rows] - com_o[rows])[:, [0, 3]].max() < TOL_COM
    finally:
        ctx.preview_set_sum_mode(wg.PREVIEW_SUM_AUTO)
        ctx.preview_set_gains(wg.preview_gains())
