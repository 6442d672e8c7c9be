#!/bin/bash
# Re-entry check: parity tests, smoke, both bench arms (N=1)
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 300 gpurun_out/bench_c.json; tail -3 gpurun_out/bench_c.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c_reference.json 2>> gpurun_out/bench_c.err; cut -c1-300 gpurun_out/bench_c_reference.json
