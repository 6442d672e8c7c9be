"""Probe: closed-loop warm vs cold start - iterations per QP, trajectory differences, timing."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jrl_walkgen_b200 as wg

ctx = wg.Context(0)
rng = np.random.default_rng(11)
B = 16384
v = np.column_stack([rng.uniform(-0.2, 0.3, B), rng.uniform(-0.15, 0.15, B), rng.uniform(-0.2, 0.2, B)])
res = {}
for warm in (0, 1):
    ctx.herdt_set_params()
    p = wg.herdt_mpc_default_params(); p.warm_start = warm
    ctx.herdt_mpc_set_params(p)
    st = ctx.herdt_mpc_init(B)
    t0 = time.time()
    _, s, _ = ctx.herdt_mpc_run(st, 60, vel_ref=v, steps=True)
    dt = time.time() - t0
    res[warm] = (s, st.copy())
    print(f"warm={warm}: {dt:.3f}s host wall, iterations/QP {st['iterations_total'].sum() / st['qp_count'].sum():.2f}, fails {st['fail_count'].sum()}")
    it = s["iterations"]
    print("  per-period mean iterations:", np.round(it.mean(axis=0)[::4], 1))
sc, sw = res[0][0], res[1][0]
for f in ("jerk_x", "jerk_y", "next_foot_x", "next_foot_y"):
    print(f, "max |warm-cold|", np.abs(sc[f] - sw[f]).max())
print("com_x", np.abs(sc["com_x"] - sw["com_x"]).max())
k = np.unravel_index(np.abs(sc["jerk_y"] - sw["jerk_y"]).argmax(), sc["jerk_y"].shape)
print("worst at", k, sc["jerk_y"][k], sw["jerk_y"][k], "n_active", sc["n_active"][k], sw["n_active"][k], "iters", sc["iterations"][k], sw["iterations"][k])
