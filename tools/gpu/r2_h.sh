#!/bin/bash
# Wieber bench leg only
timeout 1200 python bench.py --steps 3 --warmup 1 --no-herdt --no-kajita --no-pldp --no-dimitrov --no-sweep --passes-per-step 2 --e2e-passes 2 --cpu-seconds 6 --wieber-walks ${1:-512} > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -5 gpurun_out/bench_r2h.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2h.json'))
print(json.dumps(d['wieber_front_to_back'], indent=1)[:3000])
PY
