#!/bin/bash
mkdir -p gpurun_out
for c in 3 4; do for w in 4096 16384; do
WG_DIMITROV_CTAS=$c timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 0.5 --no-herdt --no-kajita --no-pldp --dimitrov-walks $w > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err; tail -3 gpurun_out/bench_o.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_o.json').read().strip().splitlines()[-1])
q=d['dimitrov_front_to_back']; print($c, $w, 'dimitrov', q['qp_periods_per_s'], q['ms_per_pass'], q['kernels']['dimitrov_kernel']['avg_ms'])
PY
done; done
