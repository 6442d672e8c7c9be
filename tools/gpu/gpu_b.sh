#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-herdt --no-pldp --no-kajita --no-dimitrov --cpu-seconds 0.5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -2 gpurun_out/bench_b.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline'])"
