import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import jrl_walkgen_b200 as wg, herdt_oracle as ho
import test_herdt_mpc_gpu as T
ctx = wg.Context(0)
ctx.herdt_set_params(); p = wg.herdt_mpc_default_params(); p.foot_vel_limit=0.0; p.return_to_centre=0; ctx.herdt_mpc_set_params(p)
for name, script, nqp in (("online", ho.ONLINE_SCRIPT, 1117), ("emergency", ho.EMERGENCY_SCRIPT, 225)):
    sched = [(t // 20 + 1, v) for t, v, _ in script]; stop = [t // 20 + 1 for t, _, s in script if s][0]
    ticks, steps, st = T.gpu_run_script(ctx, 1, sched, nqp + 2, initial_support=(0.0, 0.1, 0.0), stop_at=stop)
    rows = T.rows_from_ticks(ticks[0][:nqp*20]); np.save('gpurun_out/rows_%s.npy' % name, rows)
    gold = ho.load_golden(name); g = gold[7:7+len(rows),1:37]; err=np.abs(rows[:len(g)]-g)
    print(name, st['qp_count'], st['online_mode'], 'max per col', np.round(err.max(axis=0)*1e7,1))
