"""Probe: the first Wieber QPs through wg_qld_solve_batch (shared Hessian) vs the reference ql0001_."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jrl_walkgen_b200 as wg
import oracle_lib as ol, wieber_oracle as wo, zmpdisc_oracle as zo, dimitrov_oracle as do
from test_qld_gpu import ref_qld
from test_wieber import short_walk, walk_steps

ctx = wg.Context(0)
N, T = 75, 0.02
Cm, OptB, OptC = wo.constants()
print("cond(C) =", np.linalg.cond(Cm))
w = short_walk(4)
L, R, st, t, z = wo.inputs(w)
lci = do.fcals(w["left"], w["right"], w["types"][:, 1])
o = ol.oracle()
o.oracle_wieber_build.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_double, ol.D, ol.D, ol.D, C.POINTER(C.c_int)]
ctx.qld_set_shared_hessian(Cm)
xk = np.zeros(6)
for li in (0, 50, 120, 200):
    Px = np.zeros(8 * N + 1); Pu = np.zeros((8 * N + 1) * 2 * N); nb = C.c_int(0)
    rc = o.oracle_wieber_build(N, T, li * T, len(lci), lci.ctypes.data, 0.80, ol.dptr(xk), ol.dptr(Px), ol.dptr(Pu), C.byref(nb))
    m = nb.value
    A = Pu[:(m + 1) * 2 * N].reshape(2 * N, m + 1).T[:m].copy()
    b = Px[:m].copy()
    ZMPRef = np.concatenate([z[li * 4 + 4 * np.arange(N), 0], z[li * 4 + 4 * np.arange(N), 1]])
    D = OptB @ xk - OptC @ ZMPRef
    xr, ur, fr = ref_qld(Cm, D, A, b)
    As = np.zeros((1, m + 1, 2 * N)); As[0, :m] = A
    bs = np.zeros((1, m + 1)); bs[0, :m] = b
    x, u, ifail, it = ctx.qld_solve(D[None], As, bs, np.array([m]))
    print(f"li={li}: m={m} ref ifail {fr} active {int((ur[:m] != 0).sum())}; gpu ifail {ifail[0]} iterations {it[0]} "
          f"max|x-xr| {np.abs(x[0] - xr).max():.2e} (|xr| {np.abs(xr).max():.2e}) min slack gpu {(A @ x[0] + b).min():.2e} ref {(A @ xr + b).min():.2e}")
out = ctx.wieber_run([walk_steps(4, 0.2)], np.array([zo.INIT_FEET]))
print("run:", out["status"], out["periods_done"], out["period_counts"], out["qp_iterations"])

def exact_kkt(Cm, D, A, b, act):
    """Equality-constrained QP on the active rows in extended precision (Gaussian elimination with partial pivoting)."""
    LD = np.longdouble
    n = len(D); k = len(act)
    K = np.zeros((n + k, n + k), dtype=LD); rhs = np.zeros(n + k, dtype=LD)
    K[:n, :n] = Cm.astype(LD); K[:n, n:] = -A[act].T.astype(LD); K[n:, :n] = A[act].astype(LD)
    rhs[:n] = -D.astype(LD); rhs[n:] = -b[act].astype(LD)
    M = np.concatenate([K, rhs[:, None]], axis=1)
    N_ = n + k
    for c in range(N_):
        piv = c + int(np.argmax(np.abs(M[c:, c])))
        if piv != c:
            M[[c, piv]] = M[[piv, c]]
        M[c] = M[c] / M[c, c]
        f = M[:, c].copy(); f[c] = 0
        M -= f[:, None] * M[c][None, :]
    return M[:n, -1].astype(np.float64), M[n:, -1].astype(np.float64)

li = 120
xk = np.zeros(6)
Px = np.zeros(8 * N + 1); Pu = np.zeros((8 * N + 1) * 2 * N); nb = C.c_int(0)
o.oracle_wieber_build(N, T, li * T, len(lci), lci.ctypes.data, 0.80, ol.dptr(xk), ol.dptr(Px), ol.dptr(Pu), C.byref(nb))
m = nb.value
A = Pu[:(m + 1) * 2 * N].reshape(2 * N, m + 1).T[:m].copy(); b = Px[:m].copy()
ZMPRef = np.concatenate([z[li * 4 + 4 * np.arange(N), 0], z[li * 4 + 4 * np.arange(N), 1]])
D = OptB @ xk - OptC @ ZMPRef
xr, ur, fr = ref_qld(Cm, D, A, b)
As = np.zeros((1, m + 1, 2 * N)); As[0, :m] = A
bs = np.zeros((1, m + 1)); bs[0, :m] = b
x, u, ifail, it = ctx.qld_solve(D[None], As, bs, np.array([m]))
act_r = np.nonzero(ur[:m] != 0)[0]; act_g = np.nonzero(u[0, :m] != 0)[0]
print("active ref", act_r, "gpu", act_g)
xe, ue = exact_kkt(Cm, D, A, b, act_r)
f = lambda xx: float(0.5 * xx.astype(np.longdouble) @ (Cm.astype(np.longdouble) @ xx.astype(np.longdouble)) + D.astype(np.longdouble) @ xx.astype(np.longdouble))
print("exact-on-ref-active-set: |xe-xr|", np.abs(xe - xr).max(), "|xe-xgpu|", np.abs(xe - x[0]).max(), "ue", ue, "min slack exact", (A @ xe + b).min())
print("objective ref", f(xr), "gpu", f(x[0]), "exact", f(xe))
print("x[0], x[N]: ref", xr[0], xr[N], "gpu", x[0, 0], x[0, N], "exact", xe[0], xe[N])
