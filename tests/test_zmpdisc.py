"""Kajita2003 front end (footsteps -> 5 ms ZMP reference + feet): oracle pins (CPU) and CUDA-vs-oracle parity (GPU).

Pins: the four TestKajita2003 datrefs of the reference (tests/golden/kajita_*.npz = columns 11-13, 20-25, 32-36: feet
and world ZMP reference, which do not depend on the proprietary HRP-2 model).  The datref is truncated to 7 decimals
(tests/TestObject.cpp:48-56) and the reference's own comparison accepts 1e-6 (:475-495); the restatement reproduces all
18 088 rows to 1.01e-7 (the truncation), so the tolerance here is 1.2e-7.
"""
import ctypes as C

import numpy as np
import pytest

import zmpdisc_oracle as zo

PROFILES = ("StraightWalking", "Circle", "PbFlorentSeq1", "PbFlorentSeq2")
DATREF_TOL = 1.2e-7


def stacked(o, n):
    return np.hstack([o["left"][:n], o["right"][:n], o["zmp"][:n, :2]])


# ------------------------------------------------------------------------------------------------
# CPU: the oracle against the reference's golden vectors; host-only entry points of the product
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", PROFILES)
def test_oracle_reproduces_the_kajita_datref(name):
    p = zo.default_params()
    steps = zo.profile_steps(name)
    o = zo.run(p, steps)
    g = zo.golden(name)
    L = len(o["zmp"])
    # GetZMPDiscretization emits 2*NL more samples than the test logs: the two preview stages consume one window each
    assert L == len(g) + 640 == zo.sample_count(p, steps)
    err = np.abs(stacked(o, len(g)) - g)
    assert err.max() < DATREF_TOL, (name, err.max(axis=0))
    # the reference's own tolerance, on the un-truncated side
    assert (err < 1e-6).all()


def test_oracle_segment_structure():
    """Lead-in 640 samples at the start ZMP, 160 per step (4 double support + 156 single support), 2 + 960 at the end."""
    p = zo.default_params()
    o = zo.run(p, zo.profile_steps("StraightWalking"))
    t = o["types"]
    assert (t[:640, 0] == 0).all() and (t[:640, 1:] == 10).all()
    assert (t[640:644] == 11).all()                       # double support: stepType + 10
    assert (t[644:800, 0] == -1).all()                    # right foot is the first support (sy < 0)
    assert (t[644:800, 1] == 1).all() and (t[644:800, 2] == -1).all()   # left swings
    assert (t[-962:] == 0).all()
    assert np.abs(o["zmp"][:640, :2]).max() == 0.0
    assert abs(o["left"][644:800, 2].max() - 0.07) < 1e-4     # step height
    assert o["left"][800, 2] == 0.0


def test_host_step_stack_matches_oracle():
    """wg_steps_* (product, host side) against the oracle's restatement of StepStackHandler: bitwise."""
    from jrl_walkgen_b200 import _capi
    lib = _capi.load()
    a = zo.steps_circle()
    b = zo.steps_circle(builder=lib)
    assert len(a) == len(b) == 5
    for f in ("sx", "sy", "theta", "ss_time", "ds_time", "step_type"):
        assert np.array_equal(a[f], b[f]), f
    rng = np.random.default_rng(5)
    for _ in range(50):
        x, y, arc = rng.uniform(-1, 1), rng.uniform(-2, 2), rng.uniform(10, 200)
        sf = int(rng.choice([-1, 1]))
        out = []
        for pre, L in (("oracle_", zo.lib()), ("wg_", lib)):
            s = np.zeros(256, dtype=zo.REL_STEP_DTYPE)
            n = C.c_int(0); keep = C.c_int(0)
            rc = getattr(L, pre + "steps_arc")(C.c_void_p(s.ctypes.data), 256, C.byref(n), C.c_double(x), C.c_double(y),
                                               C.c_double(arc), sf, C.c_double(0.7), C.c_double(0.1), C.byref(keep))
            assert rc == 0
            out.append((s[:n.value].copy(), keep.value))
        assert out[0][1] == out[1][1] and len(out[0][0]) == len(out[1][0])
        assert out[0][0].tobytes() == out[1][0].tobytes()
    # capacity is an error, not an overflow
    s = np.zeros(2, dtype=zo.REL_STEP_DTYPE); n = C.c_int(0); keep = C.c_int(0)
    assert lib.wg_steps_arc(s.ctypes.data, 2, C.byref(n), 0.0, 0.75, 90.0, -1, 0.78, 0.02, C.byref(keep)) == _capi.WG_ERR_INVALID


def test_sample_count_matches_oracle():
    import jrl_walkgen_b200 as wg
    from jrl_walkgen_b200 import _capi
    lib = _capi.load()
    p = wg.zmpdisc_default_params()
    po = zo.default_params()
    assert bytes(p) == bytes(po)
    for name in PROFILES:
        st = zo.profile_steps(name)
        assert lib.wg_zmpdisc_sample_count(C.byref(p), len(st), st.ctypes.data) == len(zo.run(po, st)["zmp"])
    st = zo.profile_steps("StraightWalking").copy()
    st["ss_time"][3], st["ds_time"][3] = 0.6, 0.1
    st["ss_time"][7], st["ds_time"][7] = 1.0, 0.2
    assert lib.wg_zmpdisc_sample_count(C.byref(p), len(st), st.ctypes.data) == len(zo.run(po, st)["zmp"]) == 4002 - 20 + 80


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA path against the oracle and the golden vectors, through the C ABI
# ------------------------------------------------------------------------------------------------
def gpu_discretize(ctx, step_lists, feet=None, params=None, mem=None):
    import jrl_walkgen_b200 as wg
    lens = np.array([len(s) for s in step_lists], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    steps = np.concatenate(step_lists).astype(wg.REL_STEP_DTYPE)
    B = len(step_lists)
    feet = np.tile(zo.INIT_FEET, (B, 1)) if feet is None else feet
    plan = ctx.kajita_plan(off, steps, feet, params)
    n = plan.total_samples
    z = np.zeros((n, 2)); th = np.zeros(n)
    left = np.zeros(n, dtype=wg.KAJITA_FOOT_DTYPE); right = np.zeros(n, dtype=wg.KAJITA_FOOT_DTYPE)
    types = np.zeros((n, 3), dtype=np.int32)
    plan.discretize(z, th, left, right, types)
    return plan, z, th, left.view(np.float64).reshape(n, 6), right.view(np.float64).reshape(n, 6), types


def check_against_oracle(po, step_lists, plan, z, th, left, right, types, feet=None, tol=1e-12):
    so = plan.sample_offsets
    for b, st in enumerate(step_lists):
        o = zo.run(po, st, zo.INIT_FEET if feet is None else feet[b])
        a, e = int(so[b]), int(so[b + 1])
        assert e - a == len(o["zmp"]), (b, e - a, len(o["zmp"]))
        assert np.array_equal(types[a:e], o["types"]), b
        for name, got, want in (("zmp", z[a:e], o["zmp"][:, :2]), ("theta", th[a:e], o["zmp"][:, 2]),
                                ("left", left[a:e], o["left"]), ("right", right[a:e], o["right"])):
            err = np.abs(got - want).max()
            assert err <= tol, (b, name, err)


@pytest.mark.gpu
def test_cuda_reproduces_the_kajita_datrefs_and_the_oracle(ctx):
    lists = [zo.profile_steps(n) for n in PROFILES]
    plan, z, th, left, right, types = gpu_discretize(ctx, lists)
    so = plan.sample_offsets
    for b, name in enumerate(PROFILES):
        g = zo.golden(name)
        a = int(so[b])
        got = np.hstack([left[a:a + len(g)], right[a:a + len(g)], z[a:a + len(g)]])
        err = np.abs(got - g)
        assert err.max() < DATREF_TOL, (name, err.max(axis=0))
    check_against_oracle(zo.default_params(), lists, plan, z, th, left, right, types)
    plan.destroy()


@pytest.mark.gpu
def test_cuda_random_walks_match_oracle(ctx):
    from jrl_walkgen_b200 import workloads as W
    off, steps, feet = W.kajita_steps_batch(96, seed=3)
    lists = [steps[off[b]:off[b + 1]] for b in range(96)]
    plan, z, th, left, right, types = gpu_discretize(ctx, lists, feet)
    check_against_oracle(zo.default_params(), lists, plan, z, th, left, right, types, feet)
    # support foot placement = the ZMP reference plateau of each single support: identical footstep placements
    plan.destroy()


@pytest.mark.gpu
def test_cuda_edge_cases_match_oracle(ctx):
    """One-step walks, per-step timings, step-over ZMP shifts (types 3/4/5), toe/heel rotation (omega != 0), a neutral
    ZMP off the ankle, rotated initial feet, a very short step segment."""
    import jrl_walkgen_b200 as wg
    rng = np.random.default_rng(11)
    base = zo.profile_steps("PbFlorentSeq1")
    one = base[:1].copy()
    timed = base.copy()
    timed["ss_time"] = rng.uniform(0.4, 1.2, len(base)).round(2)
    timed["ds_time"] = rng.uniform(0.02, 0.3, len(base)).round(2)
    over = zo.profile_steps("StraightWalking").copy()
    over["step_type"][4:9] = [3, 4, 5, 3, 4]
    short = zo.profile_steps("StraightWalking").copy()
    short["ss_time"][5], short["ds_time"][5] = 0.03, 0.01        # 8 samples: shorter than the filter history
    lists = [one, timed, over, short, zo.profile_steps("Circle")]
    feet = np.tile(zo.INIT_FEET, (len(lists), 1))
    feet[1] = [0.01, 0.1, 5.0, -0.01, -0.09, -3.0]
    for params in (wg.zmpdisc_default_params(), None):
        if params is None:
            params = wg.zmpdisc_default_params()
            params.omega = 4.0
            params.zmp_neutral[0], params.zmp_neutral[1] = 0.01, -0.005
            params.zmp_shift[0], params.zmp_shift[1], params.zmp_shift[2], params.zmp_shift[3] = 0.015, 0.02, 0.01, 0.005
            params.step_height = 0.05
            params.t_single, params.t_double = 0.7, 0.1
        po = zo.ZmpDiscParams.from_buffer_copy(bytes(params))
        plan, z, th, left, right, types = gpu_discretize(ctx, lists, feet, params)
        check_against_oracle(po, lists, plan, z, th, left, right, types, feet, tol=2e-12)
        plan.destroy()


@pytest.mark.gpu
def test_invalid_step_timing_is_refused(ctx):
    import jrl_walkgen_b200 as wg
    st = zo.profile_steps("StraightWalking").copy()
    st["ds_time"][3], st["ss_time"][3] = 0.001, 0.5          # rounds to zero double-support samples
    with pytest.raises(wg.WalkgenError) as ei:
        ctx.kajita_plan(np.array([0, len(st)]), st, zo.INIT_FEET)
    assert ei.value.code == -3


@pytest.mark.gpu
def test_footsteps_to_com_pipeline(ctx):
    """wg_kajita_run_batch = GetZMPDiscretization + the preview loop, on the device; host (chunked, pipelined copies) and
    device modes agree bitwise and match oracle(ZMPDiscretization) -> oracle(OneIterationOfPreview)."""
    import jrl_walkgen_b200 as wg
    import oracle_lib as ol
    from jrl_walkgen_b200 import workloads as W
    B = 40
    off, steps, feet = W.kajita_steps_batch(B, seed=9)
    gains = wg.preview_gains(0.005, 1.6, 0.8078, wg.MODE_WITHOUT_INITIALPOS)
    ctx.preview_set_gains(gains)
    plan = ctx.kajita_plan(off, steps, feet)
    n = plan.total_samples
    st = np.zeros((B, 8)); com = np.zeros((n, 6)); zmp = np.zeros((n, 2)); zref = np.zeros((n, 2))
    left = np.zeros(n, dtype=wg.KAJITA_FOOT_DTYPE); right = np.zeros(n, dtype=wg.KAJITA_FOOT_DTYPE)
    plan.run(st, com, zmp, zref, left, right)
    # oracle chain
    po = zo.default_params()
    zo_all = np.concatenate([zo.run(po, steps[off[b]:off[b + 1]], feet[b])["zmp"][:, :2] for b in range(B)])
    assert np.abs(zref - zo_all).max() < 1e-12
    og = ol.OracleGains(0.005, 1.6, 0.8078, 1)
    st_o = np.zeros((B, 8))
    com_o, zmp_o, nsteps = ol.oracle_preview_batch(og, plan.sample_offsets, np.ascontiguousarray(zo_all), st_o)
    assert nsteps == plan.total_steps
    assert np.abs(com - com_o).max() < 1e-9 and np.abs(zmp - zmp_o).max() < 1e-9 and np.abs(st - st_o).max() < 1e-9
    # device mode, single launch
    d_st = ctx.to_device(np.zeros((B, 8))); d_com = ctx.alloc(n * 48); d_zmp = ctx.alloc(n * 16)
    ctx.lib.wg_memset_device(ctx.h, d_com.ptr, 0, n * 48); ctx.lib.wg_memset_device(ctx.h, d_zmp.ptr, 0, n * 16)
    plan.run(d_st, d_com, d_zmp, mem=wg.WG_MEM_DEVICE)
    ctx.sync()
    assert np.array_equal(d_com.download(np.float64, (n, 6)), com)
    assert np.array_equal(d_zmp.download(np.float64, (n, 2)), zmp)
    assert np.array_equal(d_st.download(np.float64, (B, 8)), st)
    # new footsteps into the same plan
    off2, steps2, feet2 = W.kajita_steps_batch(B, seed=9)
    steps2["sx"] *= 0.5
    plan.set_steps(steps2, feet2)
    st2 = np.zeros((B, 8)); com2 = np.zeros((n, 6))
    plan.run(st2, com2)
    assert np.abs(com2 - com).max() > 1e-3
    for b in (d_st, d_com, d_zmp):
        b.free()
    plan.destroy()
