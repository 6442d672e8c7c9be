#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zmpdisc.py tests/test_dimitrov.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.5 --no-herdt --no-pldp --no-dimitrov > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; tail -3 gpurun_out/bench_y.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_y.json').read().strip().splitlines()[-1])
k=d['kajita_front_end']; print('kajita', k['preview_steps_per_s'], k['ms_per_pass'], k['zmpdisc_ms_per_launch'], k['zmpdisc_roofline']['frac'], 'e2e', k['e2e']['value'])
PY
