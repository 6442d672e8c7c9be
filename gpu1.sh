python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cat > /tmp/pv.py <<'PY'
import numpy as np, time, sys
sys.path.insert(0,'.')
import jrl_walkgen_b200 as wg
ctx = wg.Context(0)
g = wg.preview_gains(); ctx.preview_set_gains(g)
rng = np.random.default_rng(0)
B=4096
lens = rng.integers(2800,4600,size=B)
off = np.concatenate([[0],np.cumsum(lens)]).astype(np.int64)
n=int(off[-1])
z = rng.normal(size=(n,2))
plan = ctx.preview_plan(off)
dz = ctx.to_device(z); ds = ctx.to_device(np.zeros((B,8))); dc = ctx.alloc(n*48); dzo = ctx.alloc(n*16)
for i in range(3): plan.run(dz, ds, dc, dzo, True, mem=wg.WG_MEM_DEVICE)
ctx.sync()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_preview_v1.csv python /tmp/pv.py > gpurun_out/ncu.log 2>&1
tail -8 gpurun_out/launches_preview_v1.csv
