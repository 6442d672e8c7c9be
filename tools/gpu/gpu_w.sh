#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_herdt_gpu.py tests/test_herdt_mpc_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 --no-kajita --no-pldp --no-dimitrov > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -3 gpurun_out/bench_w.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w.json').read().strip().splitlines()[-1])
h=d['herdt']; print('herdt open', h['qp_solves_per_s'], h['qp_ms_per_launch'], 'closed', h['closed_loop_qp_solves_per_s'], 'fails', h['failures'], h['closed_loop_failures'], 'e2e', h['e2e']['value'])
PY
