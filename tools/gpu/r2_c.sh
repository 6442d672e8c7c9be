#!/bin/bash
# quick: PLDP / Dimitrov tests + pldp/dimitrov bench legs
timeout 900 python -m pytest tests/test_pldp_gpu.py tests/test_dimitrov.py tests/test_host_cpp_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 2 --no-herdt --no-kajita --no-sweep --passes-per-step 2 --e2e-passes 2 --cpu-seconds 1 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print({k:v['avg_ms'] for k,v in d['kernels'].items()})
print('pldp ranked', d['pldp']['pldp_solves_per_s'], 'e2e', d['pldp']['e2e']['value'], 'dense', d['pldp']['dense_entry']['pldp_solves_per_s'], d['pldp']['dense_entry']['e2e']['value'])
print('dimitrov', d['dimitrov_front_to_back']['qp_periods_per_s'], d['dimitrov_front_to_back']['ms_per_pass'])
PY
