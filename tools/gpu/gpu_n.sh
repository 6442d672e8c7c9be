#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dimitrov.py tests/test_pldp_gpu.py tests/test_host_cpp_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.5 --no-herdt --no-kajita > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; tail -3 gpurun_out/bench_n.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1])
p=d['pldp']; print('pldp', p['pldp_solves_per_s'], p['ms_per_launch'], p['failures'])
q=d['dimitrov_front_to_back']; print('dimitrov', q['qp_periods_per_s'], q['ms_per_pass'], q['kernels'], q['walks_completed'], q['e2e']['value'])
PY
