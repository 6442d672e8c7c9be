#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dimitrov.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 0.5 --no-herdt --no-kajita --no-pldp > gpurun_out/bench_fc.json 2> gpurun_out/bench_fc.err; tail -3 gpurun_out/bench_fc.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_fc.json').read().strip().splitlines()[-1])
q=d['dimitrov_front_to_back']; print('dimitrov', q['qp_periods_per_s'], q['ms_per_pass'], q['kernels'])
PY
