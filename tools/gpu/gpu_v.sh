#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preview.py tests/test_zmpdisc.py tests/test_host_cpp_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.5 --no-herdt --no-kajita --no-pldp --no-dimitrov > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; tail -3 gpurun_out/bench_v.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e'])
PY
