"""Debug helper: find closed-loop QPs the GPU solver flags as failed; dump their inputs to gpurun_out/mpc_fails.npz."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import jrl_walkgen_b200 as wg
ctx = wg.Context(0); ctx.herdt_set_params(); ctx.herdt_mpc_set_params()
rng = np.random.default_rng(5)
B = 16384
v = np.column_stack([rng.uniform(-0.2, 0.3, B), rng.uniform(-0.15, 0.15, B), rng.uniform(-0.2, 0.2, B)])
st = ctx.herdt_mpc_init(B)
bad_in, bad_code, bad_step, bad_b = [], [], [], []
iters = []
for k in range(30):
    _, steps, qin = ctx.herdt_mpc_run(st, 1, vel_ref=v if k == 0 else None, steps=True, qp_in=True)
    f = steps["fail"][:, 0]
    iters.append(steps["iterations"][:, 0].copy())
    idx = np.nonzero(f)[0]
    for i in idx[:50]:
        bad_in.append(qin[i].copy()); bad_code.append(int(f[i])); bad_step.append(k); bad_b.append(int(i))
print("fails", len(bad_in), "codes", np.unique(bad_code, return_counts=True))
it = np.array(iters)
print("iterations mean per step", it.mean(axis=1).round(1), "max", it.max())
np.savez(os.path.join(ROOT, "gpurun_out", "mpc_fails.npz"), qin=np.array(bad_in), code=np.array(bad_code), step=np.array(bad_step), b=np.array(bad_b), v=v)
