#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 --no-kajita --no-pldp --no-dimitrov --sweep > gpurun_out/bench_sp.json 2> gpurun_out/bench_sp.err; tail -3 gpurun_out/bench_sp.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sp.json').read().strip().splitlines()[-1])
h=d['herdt']; print('open', h['qp_solves_per_s'], 'closed', h['closed_loop_qp_solves_per_s'], h['closed_loop_ms_per_launch'], 'fails', h['failures'], h['closed_loop_failures']); print(d['sweep'])
PY
