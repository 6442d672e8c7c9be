#!/bin/bash
# bench with the new legs (Kajita front end, config-5 sweep) at N=1
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 --sweep --cpu-seconds 2 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -3 gpurun_out/bench_j.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_j.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step']); print(d['kajita_front_end']); print(d['sweep'])
PY
