#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preview.py tests/test_zmpdisc.py -m gpu -x -q 2>&1 | tail -3
for s in 0 2 3; do
WG_PREVIEW_SHAPE=$s timeout 600 python bench.py --steps 20 --warmup 3 --no-herdt --no-pldp --cpu-seconds 0.2 > gpurun_out/bench_f$s.json 2> gpurun_out/bench_f$s.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_f$s.json').read().strip().splitlines()[-1])
print($s, d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels'])
PY
done
WG_PREVIEW_SHAPE=3 timeout 900 python -m pytest tests/test_preview.py -m gpu -x -q 2>&1 | tail -3
