import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import jrl_walkgen_b200 as wg
import dimitrov_oracle as do, zmpdisc_oracle as zo
name = sys.argv[1] if len(sys.argv) > 1 else "PbFlorentSeq2"
steps = zo.profile_steps(name)
o = zo.run(zo.default_params(), steps)
par = do.default_params(); par.cold_restart = 1; par.merge_duplicate_rows = 1
ref = do.run(o["left"], o["right"], o["types"][:, 1].copy(), par)
ctx = wg.Context(0)
gp = wg.dimitrov_default_params(); gp.cold_restart = 1; gp.merge_duplicate_rows = 1
ctx.dimitrov_set_params(gp)
out = ctx.dimitrov_run([steps], [zo.INIT_FEET])
per, rper = out["periods"][0], ref["periods"]
dj = np.abs(per["jerk_x"] - rper["jerk_x"]) + np.abs(per["jerk_y"] - rper["jerk_y"])
bad = np.nonzero(dj > 1e-6)[0]
print("first bad period", bad[:5], "of", len(per))
k0 = max(0, bad[0] - 2)
for k in range(k0, min(len(per), bad[0] + 4)):
    a, b = per[k], rper[k]
    print(k, "t", a["t_start"], "m", a["m"], b["m"], "st", a["status"], b["status"], "it", a["iterations"], b["iterations"], "nact", a["n_active"], b["n_active"],
          "djerk", dj[k], "dxk", np.abs(a["xk"] - b["xk"]).max())
    print("   gpu act", a["active"][:a["n_active"]]); print("   ref act", b["active"][:b["n_active"]])
P = ctx.fcals_build(o["left"], o["right"], o["types"])
Pr = do.fcals(o["left"], o["right"], o["types"][:, 1].copy(), par)
t = per["t_start"][bad[0]]
for i in range(len(P)):
    if P["t_end"][i] >= t and P["t_start"][i] <= t + 1.6:
        print("poly", i, P["state"][i], P["rows"][i], Pr["rows"][i], "dA", np.abs(P["A"][i] - Pr["A"][i]).max(), "dB", np.abs(P["B"][i] - Pr["B"][i]).max())
        print(Pr["A"][i][:Pr["rows"][i]].tolist(), Pr["B"][i][:Pr["rows"][i]].tolist())
