#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.5 --no-herdt --no-kajita --no-dimitrov > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; tail -3 gpurun_out/bench_x.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_x.json').read().strip().splitlines()[-1])
p=d['pldp']; print('pldp', p['pldp_solves_per_s'], p['ms_per_launch'], p['failures'])
PY
