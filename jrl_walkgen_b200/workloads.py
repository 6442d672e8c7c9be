"""Synthetic workloads of the shapes BASELINE.json names (host-side numpy; no algorithmic product code).

Config 2 - "Kajita2003 preview control batched: 4096 random straight/circle footstep walks, 320-tap preview":
each walk is a footstep list (straight: 8-20 steps of 0.05-0.25 m; circle: arc of radius 0.5-2 m walked with
0.15 m steps, in the spirit of StepStackHandler::CreateArcInStepStack, src/StepStackHandler.cpp:299-457) turned
into a 5 ms ZMP reference with the timing of the reference's ZMPDiscretization (lead-in of 2 preview windows =
640 samples, 160 samples per step = 0.78 s single support + 0.02 s double support, tail of 2 + 960 samples;
src/ZMPRefTrajectoryGeneration/ZMPDiscretization.cpp:319-513, 573-1020, 1129-1300).  The ZMP sits under the
support foot during single support and moves linearly between the feet during double support.  This generator
produces the *shape* of those references for throughput measurements; it is not a parity restatement of
ZMPDiscretization.
"""
from __future__ import annotations

import numpy as np

def shard_instances(B, rank, world):
    """Instances are independent: rank r of `world` owns the contiguous index range [lo, hi)."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


NL = 320
LEAD_IN = 2 * NL
PER_STEP = 160
SS_SAMPLES = 156
TAIL = 2 + 3 * NL
HALF_FEET = 0.095


def _footsteps(rng):
    """-> (k, 2) support-foot positions of one walk, starting with the right foot."""
    if rng.random() < 0.5:
        n = int(rng.integers(8, 21))
        adv = rng.uniform(0.05, 0.25, size=n)
        heading = np.zeros(n)
    else:
        R = rng.uniform(0.5, 2.0)
        arc = np.deg2rad(rng.uniform(30.0, 180.0))
        n = int(np.clip(np.ceil(arc * R / 0.15), 8, 20))
        adv = np.full(n, arc * R / n)
        heading = np.cumsum(np.full(n, arc / n))
    cx = np.cumsum(adv * np.cos(heading))
    cy = np.cumsum(adv * np.sin(heading))
    side = np.where(np.arange(n) % 2 == 0, -1.0, 1.0) * HALF_FEET
    px = cx - side * np.sin(heading)
    py = cy + side * np.cos(heading)
    return np.stack([px, py], axis=1), (cx[-1], cy[-1])


def preview_walk(rng):
    """One ZMP reference, shape (L, 2) with L = 640 + 160*steps + 962."""
    feet, last = _footsteps(rng)
    n = len(feet)
    L = LEAD_IN + PER_STEP * n + TAIL
    z = np.empty((L, 2))
    z[:LEAD_IN] = 0.0
    prev = np.zeros(2)
    ramp = (np.arange(1, PER_STEP - SS_SAMPLES + 1) / (PER_STEP - SS_SAMPLES + 1))[:, None]
    k = LEAD_IN
    for s in range(n):
        nd = PER_STEP - SS_SAMPLES
        z[k:k + nd] = prev + ramp * (feet[s] - prev)
        z[k + nd:k + PER_STEP] = feet[s]
        prev = feet[s]
        k += PER_STEP
    z[k:] = np.array(last)
    return z


def preview_batch(B, seed=0):
    """Config 2: -> (offsets int64[B+1], zmpref float64[total, 2]); walk b uses stream seed + b."""
    walks = [preview_walk(np.random.default_rng([seed, b])) for b in range(B)]
    lens = np.array([len(w) for w in walks], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return offsets, np.ascontiguousarray(np.concatenate(walks, axis=0))


def kajita_steps_batch(B, seed=0, ss=0.78, ds=0.02, straight_only=False):
    """Config 2 as FOOTSTEPS (the input of the on-GPU front end, wg_kajita_plan_create): walk b is, with equal odds,
    a straight walk (:stepseq of 8-20 steps, sx in U[0.05, 0.25], sy = -/+ U[0.19, 0.21] alternating, theta = 0, closed by
    a half step that brings the feet together) or an arc (:supportfoot 1, :arc 0 R a -1, :lastsupport with R in U[0.5, 2],
    a in U[30, 180] degrees: StepStackHandler::CreateArcInStepStack through the product's host builder wg_steps_arc).
    -> (step_offsets int64[B+1], steps REL_STEP_DTYPE[total], init_feet float64[B][6])."""
    import ctypes as C
    from . import _capi
    from . import REL_STEP_DTYPE
    lib = _capi.load()
    walks = []
    for b in range(B):
        rng = np.random.default_rng([seed, 2, b])
        if rng.random() < 0.5 or straight_only:
            n = int(rng.integers(8, 21))
            st = np.zeros(n + 2, dtype=REL_STEP_DTYPE)
            side = -1.0
            st[0]["sy"] = side * rng.uniform(0.095, 0.105)
            for i in range(1, n + 1):
                side = -side
                st[i]["sx"] = rng.uniform(0.05, 0.25)
                st[i]["sy"] = side * rng.uniform(0.19, 0.21)
            st[n + 1]["sy"] = -side * rng.uniform(0.19, 0.21)
            st["ss_time"], st["ds_time"], st["step_type"] = ss, ds, 1
        else:
            st = np.zeros(128, dtype=REL_STEP_DTYPE)
            n = C.c_int(0); keep = C.c_int(0)
            R = rng.uniform(0.5, 2.0); arc = rng.uniform(30.0, 180.0)
            assert lib.wg_steps_support_foot(st.ctypes.data, 128, C.byref(n), 1, ss, ds) == 0
            assert lib.wg_steps_arc(st.ctypes.data, 128, C.byref(n), 0.0, R, arc, -1, ss, ds, C.byref(keep)) == 0
            assert lib.wg_steps_last_support(st.ctypes.data, 128, C.byref(n), keep.value, ss, ds) == 0
            st = st[:n.value].copy()
        walks.append(st)
    lens = np.array([len(w) for w in walks], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    feet = np.tile(np.array([0.00949035, 0.095, 0.0, 0.00949035, -0.095, 0.0]), (B, 1))
    return offsets, np.ascontiguousarray(np.concatenate(walks)), np.ascontiguousarray(feet)


# ------------------------------------------------------------------------------------------------
# Config 4 - "Dimitrov ZMPQPWithConstraint PLDPSolver/OptCholesky batched: 16,384 constrained CoP QPs"
# ------------------------------------------------------------------------------------------------
class DimitrovConstants:
    """Constants of ZMPConstrainedQPFastFormulation in PLDP mode (N = 16, T = 0.1 s, zc = 0.80, alpha = 200,
    beta = 1000: src/ZMPRefTrajectoryGeneration/ZMPConstrainedQPFastFormulation.cpp:87-96), following
    InitializeMatrixPbConstants (:158-246), BuildingConstantPartOfTheObjectiveFunction (:512-614, including the
    `lterm2 = alpha * VPu^T` term of which only the diagonal reaches the Cholesky factor, :525-533 and :390-418)
    and BuildingConstantPartOfConstraintMatrices (:616-680).  Workload data, not product code."""

    def __init__(self, N=16, T=0.1, zc=0.80, alpha=200.0, beta=1000.0):
        self.N, self.T, self.zc = N, T, zc
        i = np.arange(N)[:, None]; j = np.arange(N)[None, :]
        d = i - j
        low = j <= i
        PPu = np.where(low, (1 + 3 * d + 3 * d * d) * T ** 3 / 6.0, 0.0)
        VPu = np.where(low, (2 * d + 1) * T * T * 0.5, 0.0)
        PPx = np.column_stack([np.ones(N), (i[:, 0] + 1) * T, (i[:, 0] + 1) ** 2 * T * T * 0.5])
        VPx = np.column_stack([np.zeros(N), np.ones(N), (i[:, 0] + 1) * T])
        Q = np.eye(N) + beta * PPu.T @ PPu + alpha * np.diag(np.diag(VPu))
        LQ = np.linalg.cholesky(Q)
        iLQ = np.linalg.inv(LQ)
        self.iLQ = iLQ
        self.OptB = iLQ @ (alpha * VPu.T @ VPx + beta * PPu.T @ PPx)        # N x 3 per axis
        self.OptC = iLQ @ (beta * PPu.T)                                   # N x N per axis
        PuT = np.where(j >= i, (1 + 3 * (j - i) + 3 * (j - i) ** 2) * T ** 3 / 6.0 - T * zc / 9.81, 0.0)  # ptPu[k*N+i], k<=i
        self.Pu = np.ascontiguousarray(iLQ @ PuT)                           # m_Pu = iLQ * Pu'
        self.iPu = np.ascontiguousarray(np.linalg.inv(self.Pu))
        self.Px = np.ascontiguousarray(np.column_stack([np.ones(N), (i[:, 0] + 1) * T,
                                                        (i[:, 0] + 1) ** 2 * T * T * 0.5 - zc / 9.81]))


def _support_polygons(rng, N, T, count=None):
    """Per previewed sample: (centre, rows (a0, a1, b) with a.p + b >= 0 inside).  Single support: the rectangle of
    the support foot (4 rows); double support: a hexagon around both feet (6 rows)."""
    hx, hy = 0.085, 0.03                       # ConstraintOnX/Y = 0.04 inside a 0.25 x 0.14 sole
    t0 = rng.uniform(0.0, 1.6)
    step = rng.uniform(0.05, 0.25)
    yaw = rng.uniform(-0.3, 0.3)
    c, s = np.cos(yaw), np.sin(yaw)
    out = []
    for i in range(count or N):
        t = t0 + i * T
        k = int(t // 0.8)                       # step index
        ph = t - 0.8 * k
        side = 1.0 if k % 2 == 0 else -1.0
        fx, fy = k * step, side * 0.095
        if ph < 0.1 and k > 0:                  # double support between step k-1 and k
            px, py = (k - 0.5) * step, 0.0
            ex, ey = hx + 0.5 * step, hy + 0.095
            normals = [(1, 0, ex), (-1, 0, ex), (0, 1, ey), (0, -1, ey), (0.6, 0.8, 0.9 * (0.6 * ex + 0.8 * ey)),
                       (-0.6, -0.8, 0.9 * (0.6 * ex + 0.8 * ey))]
        else:
            px, py = fx, fy
            normals = [(1, 0, hx), (-1, 0, hx), (0, 1, hy), (0, -1, hy)]
        cx, cy = c * px - s * py, s * px + c * py
        rows = []
        for (nx, ny, h) in normals:             # inward normal -n rotated by yaw: -n.(p - centre) + h >= 0
            ax, ay = -(c * nx - s * ny), -(s * nx + c * ny)
            rows.append((ax, ay, h - ax * cx - ay * cy))
        out.append(((cx, cy), rows))
    return out


def pldp_problem_from(K: DimitrovConstants, polys, xk):
    """One call of PLDPSolver::SolveProblem as ZMPConstrainedQPFastFormulation prepares it (:759-1022, :1250-1258)
    for the N support polygons `polys` and the LIPM state xk:
    -> dict(D[32], m, DPu[(m+1)*32] column-major, DPx[m], ZMPRef[32], XkYk[6], n_first = rows of the first sample)."""
    N = K.N
    m = sum(len(r) for _, r in polys)
    zref = np.zeros(2 * N)
    DPu = np.zeros((2 * N, m + 1))              # column-major storage: DPu[c, r] is element (r, c)
    DPx = np.zeros(m)
    a01 = np.zeros((m, 2)); ri = np.zeros(m, dtype=np.uint8); sim = np.zeros(m, dtype=np.int32)
    r = 0
    for i, (cen, rows) in enumerate(polys):
        zref[i], zref[i + N] = cen
        zx = xk[0] * K.Px[i, 0] + xk[1] * K.Px[i, 1] + xk[2] * K.Px[i, 2]
        zy = xk[3] * K.Px[i, 0] + xk[4] * K.Px[i, 1] + xk[5] * K.Px[i, 2]
        for (a0, a1, b) in rows:
            DPx[r] = zx * a0 + zy * a1 + b
            DPu[:N, r] = a0 * K.Pu[:, i]
            DPu[N:, r] = a1 * K.Pu[:, i]
            a01[r] = (a0, a1); ri[r] = i
            # rows come in exactly negated pairs (n, -n): the odd row of a pair is flagged, SimilarConstraints-style
            if r > 0 and ri[r - 1] == i and a01[r - 1, 0] == -a0 and a01[r - 1, 1] == -a1 and sim[r - 1] == 0:
                sim[r] = -1
            r += 1
    D = np.concatenate([K.OptB @ xk[:3] - K.OptC @ zref[:N], K.OptB @ xk[3:] - K.OptC @ zref[N:]])
    return {"D": D, "m": m, "DPu": DPu.ravel(), "DPx": DPx, "ZMPRef": zref, "XkYk": np.array(xk, dtype=np.float64),
            "n_first": len(polys[0][1]), "a01": a01, "ri": ri, "similar": sim}


def pldp_problem(K: DimitrovConstants, rng):
    """A random stand-alone problem: random walk phase and a LIPM state near the first support centre."""
    polys = _support_polygons(rng, K.N, K.T)
    xk = np.zeros(6)
    c0 = polys[0][0]
    xk[0] = c0[0] + rng.uniform(-0.02, 0.02); xk[3] = c0[1] + rng.uniform(-0.02, 0.02)
    xk[1] = rng.uniform(-0.1, 0.3); xk[4] = rng.uniform(-0.2, 0.2)
    xk[2] = rng.uniform(-0.5, 0.5); xk[5] = rng.uniform(-0.5, 0.5)
    return pldp_problem_from(K, polys, xk)


def pldp_advance(K: DimitrovConstants, xk, X):
    """The receding-horizon step of ZMPConstrainedQPFastFormulation (:1363-1397): jerk = (iLQ^T X)[0], [N]; LIPM
    OneIteration with T = 0.1 (LinearizedInvertedPendulum2D.cpp:230-264)."""
    N, T = K.N, K.T
    jx = float(K.iLQ[:, 0] @ X[:N]); jy = float(K.iLQ[:, 0] @ X[N:])
    A = np.array([[1.0, T, T * T / 2.0], [0.0, 1.0, T], [0.0, 0.0, 1.0]]); B = np.array([T ** 3 / 6.0, T * T / 2.0, T])
    return np.concatenate([A @ xk[:3] + B * jx, A @ xk[3:] + B * jy])


def pldp_pack(K, probs):
    """Pack a list of problems for wg_pldp_solve_batch (common strides = the maxima)."""
    B = len(probs)
    mmax = max(p["m"] for p in probs)
    dpu_stride = (mmax + 1) * 2 * K.N
    out = {"D": np.stack([p["D"] for p in probs]), "m": np.array([p["m"] for p in probs], dtype=np.int32),
           "DPu": np.zeros((B, dpu_stride)), "DPx": np.zeros((B, mmax)),
           "ZMPRef": np.stack([p["ZMPRef"] for p in probs]), "XkYk": np.stack([p["XkYk"] for p in probs]),
           "dpu_stride": dpu_stride, "dpx_stride": mmax,
           # the same matrices in rank-structured form (wg_pldp_solve_batch_ranked): row r = (A_r(0), A_r(1), sample i_r),
           # element (r, k + 16 ax) = A_r(ax) * Pu[k][i_r]; and SimilarConstraints-style flags
           "a0": np.zeros((B, 128)), "a1": np.zeros((B, 128)), "ri": np.zeros((B, 128), dtype=np.uint8), "row_stride": 128,
           "similar": np.zeros((B, 128), dtype=np.int32)}
    for b, p in enumerate(probs):
        out["DPu"][b, :len(p["DPu"])] = p["DPu"]
        out["DPx"][b, :p["m"]] = p["DPx"]
        out["a0"][b, :p["m"]] = p["a01"][:, 0]; out["a1"][b, :p["m"]] = p["a01"][:, 1]
        out["ri"][b, :p["m"]] = p["ri"]; out["similar"][b, :p["m"]] = p["similar"]
    return out


def pldp_batch(B, seed=0, K=None):
    """Config 4: B independent PLDP problems packed for wg_pldp_solve_batch."""
    K = K or DimitrovConstants()
    probs = [pldp_problem(K, np.random.default_rng([seed, 4, b])) for b in range(B)]
    return K, pldp_pack(K, probs)
