#!/bin/bash
# Herdt warm start: tests + herdt bench leg
timeout 1500 python -m pytest tests/test_herdt_gpu.py tests/test_herdt_mpc_gpu.py tests/test_host_cpp_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 600 python bench.py --steps 10 --warmup 2 --no-kajita --no-pldp --no-dimitrov --no-sweep --passes-per-step 2 --e2e-passes 2 --cpu-seconds 1 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print({k:v['avg_ms'] for k,v in d['kernels'].items()})
print(json.dumps(d['herdt'])[:1500])
PY
