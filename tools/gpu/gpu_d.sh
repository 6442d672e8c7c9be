#!/bin/bash
# preview kernel iteration: parity + bench of the preview leg only
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preview.py tests/test_zmpdisc.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 --no-herdt --no-pldp --cpu-seconds 1 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernels'])
PY
tail -3 gpurun_out/bench_d.err
