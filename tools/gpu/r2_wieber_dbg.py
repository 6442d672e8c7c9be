import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import jrl_walkgen_b200 as wg
import wieber_oracle as wo, zmpdisc_oracle as zo
from test_wieber import walk_steps
ctx = wg.Context(0)
steps = walk_steps(4, 0.2)
out = ctx.wieber_run([steps], np.array([zo.INIT_FEET]), want_feet=True)
w = zo.run(zo.default_params(), steps)
k, com_o, zmp_o, info = wo.run(w)
print("periods", k, out["periods_done"], "status", out["status"])
# inputs equal?
print("zmp ref in: max diff feet x", np.abs(out["left"]["x"] - w["left"][:, 0]).max(), np.abs(out["right"]["z"] - w["right"][:, 2]).max())
err = np.abs(out["com"][:4 * k] - com_o[:4 * k, :6]).max(axis=1)
bad = np.nonzero(err > 1e-9)[0]
print("first bad row", bad[:5], "period", bad[0] // 4 if len(bad) else None)
if len(bad):
    r = bad[0]
    print("gpu", out["com"][r], "\nora", com_o[r, :6])
    p = r // 4
    print("info around", info[max(0, p - 2):p + 3])
    for rr in range(max(0, r - 4), r + 8):
        print(rr, out["com"][rr, [0, 2, 3, 5]], com_o[rr, [0, 2, 3, 5]])
print("---- trajectories (x, y, zmp x, zmp y): gpu | oracle(ref QLD)")
for r in range(400, 4 * k, 150):
    print(r, np.round(out["com"][r, [0, 3]], 5), np.round(out["zmp"][r], 5), "|", np.round(com_o[r, [0, 3]], 5), np.round(zmp_o[r, :2], 5), "| zmpref", np.round(w["zmp"][r, :2], 4))
print("max |com pos diff|", np.abs(out["com"][:4 * k, [0, 3]] - com_o[:4 * k, [0, 3]]).max(), "max |zmp diff|", np.abs(out["zmp"][:4 * k] - zmp_o[:4 * k, :2]).max())
print("max |zmp - zmpref| gpu", np.abs(out["zmp"][:4 * k] - w["zmp"][:4 * k, :2]).max(), "oracle", np.abs(zmp_o[:4 * k, :2] - w["zmp"][:4 * k, :2]).max())
