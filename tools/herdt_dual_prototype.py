"""Numpy prototype of the dual-space active-set algorithm implemented by jrl_walkgen_b200/csrc/herdt_qp.cu.

DESIGN ARTEFACT (not product, not oracle): it mirrors the kernel's data flow step by step so that the
mathematics could be validated on the CPU against the oracle before the CUDA was written.  See DESIGN.md
"Herdt QP kernel" for the derivation:

  x = [jx(16) jy(16) fx(ns) fy(ns)],  Hessian = diag(Qx, Qx) with Qx = [[Qc, Cx],[Cx', E]]
  every constraint k is (a_k, b_k, c_k, point kappa_k):  s_k = a_k*PX[kappa] + b_k*PY[kappa] + c_k
  Gram(k,l) = (a_k a_l + b_k b_l) * Gamma(kappa_k, kappa_l),  Gamma = G + theta' S^-1 theta
"""
import numpy as np

N = 16
g = 9.81


def constants(T=0.1, h=0.814, wj=1e-5, wv=1.0, wc=1e-6):
    i = np.arange(N)[:, None]; j = np.arange(N)[None, :]
    low = (j <= i)
    Uv = np.where(low, (2 * (i - j) + 1) * T * T * 0.5, 0.0)
    Uz = np.where(low, (1 + 3 * (i - j) + 3 * (i - j) ** 2) * T ** 3 / 6.0 - T * h / g, 0.0)
    Sv = np.stack([np.zeros(N), np.ones(N), (np.arange(N) + 1) * T], axis=1)
    Sz = np.stack([np.ones(N), (np.arange(N) + 1) * T, ((np.arange(N) + 1) * T) ** 2 * 0.5 - h / g], axis=1)
    Qc = wj * np.eye(N) + wv * Uv.T @ Uv + wc * Uz.T @ Uz
    G0 = np.linalg.inv(Qc)
    K1 = G0 @ Uz.T          # 16x16
    G = Uz @ K1             # 16x16 symmetric
    K3 = G0 @ Uv.T          # 16x16
    K4 = K3 @ Sv            # 16x3
    return dict(Uv=Uv, Uz=Uz, Sv=Sv, Sz=Sz, Qc=Qc, G0=G0, K1=K1, G=G, K3=K3, K4=K4, wj=wj, wv=wv, wc=wc, T=T)


def hull(P, foot, phase, yaw, cop):
    if cop:
        hw, hh, dsd = P["cop_half_x"], P["cop_half_y"], P["ds_feet_distance"]
        lx = np.array([1, 1, -1, -1.0])
        ly = np.array([1, -1, -1, 1.0]) if foot == 0 else np.array([-1, 1, 1, -1.0])
        X = lx * hw
        if phase == 1:
            Y = ly * (hh + dsd / 2) + (-dsd / 2 if foot == 0 else dsd / 2)
        else:
            Y = ly * hh
    else:
        X = np.array(P["foot_hull_x"]); Y = np.array(P["foot_hull_y"]) * (1 if foot == 0 else -1)
    c, s = np.cos(yaw), np.sin(yaw)
    return X * c - Y * s, X * s + Y * c


def halfplanes(X, Y, foot):
    sign = 1.0 if foot == 0 else -1.0
    X2 = np.roll(X, -1); Y2 = np.roll(Y, -1)
    A = Y - Y2; B = X2 - X; D = A * X + B * Y
    return sign * A, sign * B, sign * D


def solve(C, P, inp, maxit=400, tol=1e-10):
    """inp: one record of QP_INPUT_DTYPE.  Returns x, lagr (with dummy row 0), iterations."""
    ns = int(inp["sup_step"][N])
    step = inp["sup_step"][1:].astype(int)          # per sample
    V = np.zeros((N, ns))
    for i in range(N):
        if step[i] > 0: V[i, step[i] - 1] = 1.0
    Vf = np.zeros((ns, ns)); Vcf = np.zeros((ns, 2))
    Vc = np.zeros((N, 2))
    for i in range(N):
        k = i + 1
        if step[i] > 0:
            if step[i] == 1 and inp["sup_changed"][k] and inp["sup_phase"][k] == 0:
                Vcf[0] = [inp["sup_x"][k - 1], inp["sup_y"][k - 1]]; Vf[0, 0] = 1.0
            elif step[i] > 1:
                Vf[step[i] - 1, step[i] - 2] = -1.0; Vf[step[i] - 1, step[i] - 1] = 1.0
        else:
            Vc[i] = [inp["sup_x"][k], inp["sup_y"][k]]
    wc, wv = C["wc"], C["wv"]
    cnt = V.sum(axis=0)
    Y = -wc * C["K1"] @ V                       # 16 x ns
    S = wc * np.diag(cnt) - wc * wc * V.T @ C["G"] @ V
    Si = np.linalg.inv(S) if ns else np.zeros((0, 0))
    npts = N + ns
    theta = np.zeros((npts, ns))
    theta[:N] = V - wc * (C["G"] @ V)           # theta_i = V_i - wc (G V)_i
    theta[N:] = -Vf
    sigma = theta @ Si                          # rows sigma_k'
    is_cop = np.arange(npts) < N

    def Gamma(k, l):
        v = theta[k] @ Si @ theta[l] if ns else 0.0
        if k < N and l < N: v += C["G"][k, l]
        return v

    # unconstrained optimum per axis
    xi0 = []; P0 = []
    for ax, (com, ref) in enumerate(((inp["com_x"], inp["ref_x"]), (inp["com_y"], inp["ref_y"]))):
        G0pj = wv * (C["K4"] @ com - C["K3"] @ ref)        # G0 pj
        pj = wv * (C["Uv"].T @ (C["Sv"] @ com - ref))
        Z = C["Sz"] @ com
        pf = -wc * V.T @ (Z - Vc[:, ax])
        tau = Si @ (pf - Y.T @ pj) if ns else np.zeros(0)
        j0 = -G0pj + Y @ tau if ns else -G0pj
        f0 = -tau
        xi0.append((j0, f0))
        pts = np.zeros(npts)
        pts[:N] = -C["Uz"] @ j0 + (V @ f0 if ns else 0.0) - Z + Vc[:, ax]   # includes the constant shift
        pts[N:] = -Vf @ f0 + Vcf[:, ax] if ns else 0.0
        P0.append(pts)
    PX, PY = P0[0].copy(), P0[1].copy()
    # constraints: (a, b, d, point); s = a*PX + b*PY + d
    a = []; b = []; d = []; pt = []
    X, Yh = hull(P, inp["sup_foot"][0], inp["sup_phase"][0], inp["sup_yaw"][0], True)
    for i in range(N):
        k = i + 1
        if inp["sup_changed"][k]:
            X, Yh = hull(P, inp["sup_foot"][k], inp["sup_phase"][k], inp["sup_yaw"][k], True)
        A_, B_, D_ = halfplanes(X, Yh, inp["sup_foot"][k])
        for e in range(4):
            a.append(A_[e]); b.append(B_[e]); d.append(D_[e]); pt.append(i)
    feet = {}
    for i in range(N):
        k = i + 1
        if inp["sup_changed"][k] and step[i] > 0 and inp["sup_phase"][k] != 1:
            Xf, Yf = hull(P, inp["sup_foot"][k - 1], inp["sup_phase"][k - 1], inp["sup_yaw"][k - 1], False)
            feet[step[i] - 1] = halfplanes(Xf, Yf, inp["sup_foot"][k])
    for s_ in range(ns):
        if s_ in feet:
            A_, B_, D_ = feet[s_]
        else:
            A_ = B_ = D_ = np.zeros(5)
        for e in range(5):
            a.append(A_[e]); b.append(B_[e]); d.append(D_[e]); pt.append(N + s_)
    a = np.array(a); b = np.array(b); d = np.array(d); pt = np.array(pt)
    m = len(a)
    gdiag = np.array([Gamma(k, k) for k in range(npts)])
    nrm = np.sqrt((a * a + b * b) * gdiag[pt])          # norm in the metric of H (dual norm)
    nrm[nrm == 0] = 1.0

    W = []; u = []; Tm = np.zeros((40, 40)); q = 0
    iters = 0
    active = np.zeros(m, bool)
    while True:
        s = a * PX[pt] + b * PY[pt] + d
        sc = np.where(active, np.inf, s / nrm)
        p = int(np.argmin(sc))
        if sc[p] >= -tol: break
        up = 0.0
        while True:
            iters += 1
            if iters > maxit: raise RuntimeError("maxit")
            gv = np.array([(a[j] * a[p] + b[j] * b[p]) * Gamma(pt[j], pt[p]) for j in W])
            Mpp = (a[p] ** 2 + b[p] ** 2) * Gamma(pt[p], pt[p])
            w = Tm[:q, :q] @ gv
            r = Tm[:q, :q].T @ w
            delta = Mpp - w @ w
            t1 = np.inf; l = -1
            for j in range(q):
                if r[j] > 0 and u[j] / r[j] < t1: t1 = u[j] / r[j]; l = j
            sp = a[p] * PX[pt[p]] + b[p] * PY[pt[p]] + d[p]
            t2 = -sp / delta if delta > 1e-13 * Mpp else np.inf
            t = min(t1, t2)
            if t == np.inf: raise RuntimeError("infeasible")
            # direction in point space
            dPX = np.zeros(npts); dPY = np.zeros(npts)
            coef = [(-r[j], W[j]) for j in range(q)] + [(1.0, p)]
            for kap in range(npts):
                for cf, j in coef:
                    gm = Gamma(kap, pt[j])
                    dPX[kap] += cf * a[j] * gm; dPY[kap] += cf * b[j] * gm
            if t2 < np.inf or True:
                PX += t * dPX; PY += t * dPY
            for j in range(q): u[j] -= t * r[j]
            up += t
            if t == t2:
                dd = np.sqrt(delta)
                Tm[q, :q] = -r / dd; Tm[q, q] = 1.0 / dd
                W.append(p); u.append(up); active[p] = True; q += 1
                break
            # drop l: rotate rows (l, r) for r = l+1..q-1 to annihilate column l of T in rows > l
            for rr in range(l + 1, q):
                x1, x2 = Tm[l, l], Tm[rr, l]
                hyp = np.hypot(x1, x2); c_, s_ = x1 / hyp, x2 / hyp
                rowl = Tm[l, :q].copy(); rowr = Tm[rr, :q].copy()
                Tm[l, :q] = c_ * rowl + s_ * rowr
                Tm[rr, :q] = -s_ * rowl + c_ * rowr
            # delete row l and column l
            Tm[:q - 1, :q] = np.delete(Tm[:q, :q], l, axis=0)
            Tm[:q - 1, :q - 1] = np.delete(Tm[:q - 1, :q], l, axis=1)
            Tm[:, q - 1] = 0.0; Tm[q - 1, :] = 0.0   # (the kernel only ever reads the lower triangle)
            active[W[l]] = False
            del W[l]; del u[l]; q -= 1
    # recover solution
    wx = np.zeros(npts); wy = np.zeros(npts)
    for j in range(q):
        wx[pt[W[j]]] += u[j] * a[W[j]]; wy[pt[W[j]]] += u[j] * b[W[j]]
    x = np.zeros(2 * N + 2 * ns)
    for ax, wv_ in enumerate((wx, wy)):
        j0, f0 = xi0[ax]
        sg = sigma.T @ wv_ if ns else np.zeros(0)          # sum_k w_k sigma_k
        jj = j0 - C["K1"] @ wv_[:N] - (Y @ sg if ns else 0.0)
        ff = f0 + sg
        x[ax * N:(ax + 1) * N] = jj
        x[2 * N + ax * ns:2 * N + (ax + 1) * ns] = ff
    lagr = np.zeros(m + 1)
    for j in range(q): lagr[1 + W[j]] = u[j]
    return x, lagr, iters


if __name__ == "__main__":
    import sys, os, ctypes as Cc
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import herdt_oracle as ho
    ev = {1000: lambda s: s.vel_ref(0.2, 0, 0), 2000: lambda s: s.vel_ref(0, 0.2, 0)}
    sim, rows = ho.run_online_script(3000, ev, logging=True)
    ins, X, U, meta = sim.log()
    p = ho.default_params()
    P = dict(cop_half_x=p.cop_half_x, cop_half_y=p.cop_half_y, ds_feet_distance=p.ds_feet_distance,
             foot_hull_x=list(p.foot_hull_x), foot_hull_y=list(p.foot_hull_y))
    C = constants()
    worst = 0; its = []
    for k in range(len(ins)):
        x, lagr, it = solve(C, P, ins[k])
        n, m = meta[k, 0], meta[k, 1]
        ex = np.abs(x - X[k, :n]).max()
        big = max(U[k].max(), 1e-12)
        same = set(np.nonzero(U[k, :m] > 1e-7 * big)[0]) == set(np.nonzero(lagr[:m] > 1e-7 * big)[0])
        worst = max(worst, ex); its.append(it)
        if ex > 1e-6 or not same:
            print("QP", k, "n", n, "err", ex, "same active", same, "iters", it)
    print("worst |x - x_QLD| =", worst, " iterations mean/max", np.mean(its), max(its))
