#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_herdt_gpu.py tests/test_herdt_mpc_gpu.py -m gpu -x -q -s 2>&1 | tail -25
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 6 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print(json.dumps(d['herdt'],indent=1)); print(d['value'], d['kernels'])"; tail -5 gpurun_out/bench2.err
