#!/bin/bash
mkdir -p gpurun_out
for s in 3 4; do
WG_PREVIEW_DBG=$s timeout 600 python bench.py --steps 20 --warmup 3 --no-herdt --no-pldp --cpu-seconds 0.2 > gpurun_out/bench_g$s.json 2> gpurun_out/bench_g$s.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_g$s.json').read().strip().splitlines()[-1])
print($s, d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels'])
PY
done
