"""The host-side C++ class mirror (jrl_walkgen_b200/host: PatternGeneratorInterface::ParseCmd, ZMPVelocityReferencedQP,
PreviewControl, OptCholesky, PLDPSolver - same names and signatures as the reference) driven by tests/cpp/host_api_test.cpp
the way the reference's own tests drive jrl-walkgen; outputs compared here against the golden datref and the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
import pldp_oracle as po
from jrl_walkgen_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_api_test")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def exe():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")], check=True)
    return EXE


@pytest.mark.gpu
def test_cpp_testoptcholesky(exe):
    """tests/TestOptCholesky.cpp through the OptCholesky class: exit code 0 = its own 1e-6 criterion holds."""
    r = subprocess.run([exe, "optcholesky"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("profile", ["online", "emergency"])
def test_cpp_testherdt2010_matches_whole_datref(exe, tmp_path, profile):
    """tests/TestHerdt2010.cpp (both profiles, every tick until the deques run empty) through ParseCmd strings and
    RunOneStepOfTheControlLoop: the 38-column trace against the reference's datref with the reference's tolerance
    (TestObject.cpp:475-495)."""
    import herdt_oracle as ho
    gold = ho.load_golden(profile)
    out = tmp_path / "TestHerdt2010.dat"
    r = subprocess.run([exe, "herdt2010", str(out), str(len(gold))] + (["emergency"] if profile == "emergency" else []),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(out)
    assert rows.shape == gold.shape
    err = np.abs(rows[:, :37] - gold[:, :37])
    acc = np.zeros(37, bool); acc[[16, 17, 28, 29]] = True      # swing-foot accelerations: see test_herdt_mpc_gpu.py
    # CoM velocity over the runaway ending of the datref-era code (t > 111 s): see test_herdt_mpc_gpu._gpu_full_datref
    tail = gold[:, 0] > 111.0
    strict = err.copy(); strict[np.ix_(tail, [5, 6])] = 0.0
    assert strict[:, ~acc].max() < 1e-6, (strict[:, ~acc].max(), np.unravel_index((strict * ~acc).argmax(), err.shape))
    assert err[:, [5, 6]].max() < 2e-6 and err[:, [1, 2, 8, 9]].max() < 5e-7
    assert err[:, acc].max() < 1e-4


@pytest.mark.gpu
def test_cpp_preview_control_matches_oracle(exe, tmp_path):
    out = tmp_path / "preview.bin"
    r = subprocess.run([exe, "preview", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out).reshape(200, 8)
    g = ol.OracleGains(0.005, 1.6, 0.814, 1)
    z = np.zeros((520, 2))
    i = np.arange(520)
    z[:, 0] = 0.2 * (i // 160); z[:, 1] = np.where((i // 160) % 2 == 1, -0.095, 0.095)
    st = np.zeros((1, 8))
    com, zmp, steps = ol.oracle_preview_batch(g, np.array([0, 519]), z[:519], st)
    assert steps == 200
    assert np.abs(got[:, :6] - com[:200]).max() < 1e-9 and np.abs(got[:, 6:] - zmp[:200]).max() < 1e-8


@pytest.mark.gpu
def test_cpp_pldpsolver_hot_sequence_matches_oracle(exe, tmp_path):
    """A hot-started receding-horizon sequence through PLDPSolver::SolveProblem (one solver object, as
    ZMPConstrainedQPFastFormulation uses it): X bitwise equal to the oracle port."""
    K = W.DimitrovConstants()
    rng = np.random.default_rng([9, 1])
    polys = W._support_polygons(rng, K.N, K.T, count=K.N + 12)
    xk = np.zeros(6); xk[0], xk[3] = polys[0][0]
    hot = np.zeros(1, dtype=po.STATE)
    n_removed = 0
    blobs, expect = [], []
    for t in range(12):
        p = W.pldp_problem_from(K, polys[t:t + K.N], xk)
        X, info, act = po.oracle_solve(K, W.pldp_pack(K, [p]), 0, hot=hot, starting=(t == 0), n_removed=n_removed)
        if info[1] != 0:
            break
        blobs.append(struct.pack("4i", p["m"], n_removed, int(t == 0), 0) + p["D"].tobytes() + p["DPu"].tobytes()
                     + p["DPx"].tobytes() + p["ZMPRef"].tobytes() + p["XkYk"].tobytes())
        expect.append(X)
        n_removed = p["n_first"]
        xk = W.pldp_advance(K, xk, X)
    assert len(expect) >= 6
    fin = tmp_path / "pldp_in.bin"; fout = tmp_path / "pldp_out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("2i", 0, len(expect)) + K.iPu.tobytes() + K.Px.tobytes() + K.Pu.tobytes())
        for b in blobs:
            f.write(b)
    r = subprocess.run([exe, "pldp", str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(fout).reshape(len(expect), 32)
    assert np.array_equal(got, np.array(expect))


@pytest.mark.gpu
@pytest.mark.parametrize("robust", (False, True))
def test_cpp_dimitrov_generator_matches_oracle(exe, tmp_path, robust):
    """ZMPConstrainedQPFastFormulation::GetZMPDiscretization (command strings + RelativeFootPosition deque in, CoM deque
    out) and FootConstraintsAsLinearSystem::BuildLinearConstraintInequalities through the class mirror, on the straight
    walk of TestKajita2003: bitwise the oracle chain, in the reference-faithful mode (stops at period 17, where the
    reference calls exit(0)) and in the robust mode (all 185 periods)."""
    import dimitrov_oracle as do
    import zmpdisc_oracle as zo
    out = tmp_path / "dim.bin"
    r = subprocess.run([exe, "dimitrov", str(out)] + (["robust"] if robust else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = out.read_bytes()
    n = struct.unpack_from("<q", raw, 0)[0]
    status, done, npoly, _ = struct.unpack_from("<4i", raw, 8)
    com = np.frombuffer(raw, dtype=np.float64, count=6 * n, offset=24).reshape(n, 6)
    rows = np.frombuffer(raw, dtype=np.int32, count=npoly, offset=24 + 48 * n)
    o = zo.run(zo.default_params(), zo.profile_steps("StraightWalking"))
    par = do.default_params()
    par.cold_restart = par.merge_duplicate_rows = int(robust)
    ref = do.run(o["left"], o["right"], o["types"][:, 1].copy(), par)
    assert n == len(o["left"])
    assert done == len(ref["periods"]) and status == (0 if ref["failed_at"] is None else 1)
    assert (done, status) == ((185, 0) if robust else (18, 1))
    good = done - status
    assert com[:20 * good].tobytes() == ref["com"][:20 * good].tobytes()
    P = do.fcals(o["left"], o["right"], o["types"][:, 1].copy())
    assert npoly == len(P) and (rows == P["rows"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("profile", ["StraightWalking", "Circle"])
def test_cpp_testkajita2003_through_parsecmd(exe, tmp_path, profile):
    """tests/TestKajita2003.cpp (StraightWalking: `:stepseq`; Circle: `:supportfoot`, `:arc`, `:lastsupport`, `:finish`)
    through the PGI mirror built by patternGeneratorInterfaceFactory over the stand-in robot, ticked with the 7-argument
    RunOneStepOfTheControlLoop overload: feet and ZMP reference (datref columns 11-36) at the datref's 1e-7 truncation on
    every row, the number of ticks of the reference (3362 for the straight walk), and the CoM of the first preview stage
    against the oracle's OneIterationOfPreview on the same ZMP reference (the datref CoM holds the multibody second stage
    of the proprietary HRP-2 model)."""
    import zmpdisc_oracle as zo
    out = tmp_path / "kajita.dat"
    r = subprocess.run([exe, "kajita2003", profile, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(out)
    gold = zo.golden(profile)                       # columns 11-13, 20-22 (left), 23-25, 32-34 (right), 35-36 (ZMP ref)
    assert rows.shape[0] == gold.shape[0]
    got = np.column_stack([rows[:, 9:15], rows[:, 15:21], rows[:, 7:9]])
    assert np.abs(got - gold).max() < 1.2e-7
    o = zo.run(zo.default_params(), zo.profile_steps(profile))
    z = np.ascontiguousarray(o["zmp"][:, :2])
    g = ol.OracleGains(0.005, 1.6, 0.8078, 1)
    st = np.zeros((1, 8))
    com, _, steps = ol.oracle_preview_batch(g, np.array([0, len(z)]), z, st)
    n = rows.shape[0]
    assert steps >= n
    assert np.abs(rows[:, 1] - com[:n, 0]).max() < 1e-9 and np.abs(rows[:, 2] - com[:n, 3]).max() < 1e-9
    assert np.abs(rows[:, 5] - com[:n, 1]).max() < 1e-8 and np.abs(rows[:, 6] - com[:n, 4]).max() < 1e-8
    assert np.abs(rows[:, 3] - 0.8078).max() == 0.0


@pytest.mark.gpu
def test_cpp_kajita_online_step_sequencing(exe, tmp_path):
    """`:StartOnLineStepSequencing` with three steps, default steps taken from the stack handler while the queues run low,
    `:StopOnLineStepSequencing` after 5 s: the walk goes on past the given steps, ends with the feet side by side, and the
    ZMP reference stays inside the box spanned by the feet (margin of half a sole)."""
    out = tmp_path / "online.dat"
    r = subprocess.run([exe, "kajita2003", "OnLine", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = np.loadtxt(out)
    assert rows.shape[0] > 1500
    lf, rf, z = rows[:, 9:12], rows[:, 15:18], rows[:, 7:9]
    assert lf[:, 2].min() >= 0.0 and rf[:, 2].min() >= 0.0 and max(lf[:, 2].max(), rf[:, 2].max()) > 0.06
    assert lf[-1, 0] > 0.5 and abs(lf[-1, 0] - rf[-1, 0]) < 1e-9 and abs(lf[-1, 1] - rf[-1, 1] - 0.19) < 1e-9
    lo = np.minimum(lf[:, :2], rf[:, :2]) - 0.13; hi = np.maximum(lf[:, :2], rf[:, :2]) + 0.13
    assert ((z >= lo) & (z <= hi)).all()
    assert np.abs(rows[-1, 1:3] - z[-1]).max() < 5e-3          # the CoM has converged onto the final ZMP reference
