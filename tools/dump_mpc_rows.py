"""Debug helper: run a walk-then-stop script on the GPU and dump rows/steps to gpurun_out/mpc_rows.npz."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import jrl_walkgen_b200 as wg
from test_herdt_mpc_gpu import gpu_run_script, rows_from_ticks
ctx = wg.Context(0); ctx.herdt_set_params(); ctx.herdt_mpc_set_params()
ticks, steps, st = gpu_run_script(ctx, 1, [(5, (0.2, 0.0, 0.0)), (60, (0.0, 0.0, 0.0))], 130)
np.savez(os.path.join(ROOT, "gpurun_out", "mpc_rows.npz"), rows=rows_from_ticks(ticks[0]), steps=steps[0])
