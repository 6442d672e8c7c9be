#!/bin/bash
# smoke + bench at the current head (new Dimitrov leg) + launch list
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 1500 python bench.py --steps 10 --warmup 3 --cpu-seconds 6 > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -5 gpurun_out/bench_l.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_l.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac']); print(json.dumps(d['dimitrov_front_to_back'], indent=1))
PY
