#!/bin/bash
timeout 900 python -m pytest tests/test_preview.py tests/test_preview_ref.py tests/test_two_stage.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --passes-per-step 8 --e2e-passes 8 --no-herdt --no-pldp --no-kajita --no-dimitrov --no-wieber --no-sweep --cpu-seconds 2 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -3 gpurun_out/bench_r2i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2i.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'pos', d['e2e']['com_position_only']['value'], d['e2e']['host_link'])
PY
