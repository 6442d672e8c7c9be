#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4; do
WG_HERDT_CHUNKS=$c timeout 900 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 --no-kajita --no-pldp --no-dimitrov > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -3 gpurun_out/bench_w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_w.json').read().strip().splitlines()[-1])
h=d['herdt']; print($c, 'herdt', h['qp_solves_per_s'], 'e2e', h['e2e']['value'])
PY
done
