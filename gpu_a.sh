#!/bin/bash
# Re-entry check: parity tests, smoke, bench N=1, config-5 sized closed loop on 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 300 gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-pldp --herdt-instances 500000 --herdt-periods 100 --cpu-seconds 1 > gpurun_out/bench_a_c5.json 2> gpurun_out/bench_a_c5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_a_c5.json').read().strip().splitlines()[-1])
print({k:d['herdt'][k] for k in ('instances','qp_solves_per_s','closed_loop_qp_solves_per_s','closed_loop_ms_per_launch','closed_loop_failures','failures','iterations_mean')})
PY
tail -3 gpurun_out/bench_a_c5.err
