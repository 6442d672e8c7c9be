#!/bin/bash
# round 2, step b: parity suite, then both bench arms at the head
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r2b_reference.json 2>> gpurun_out/bench_r2b.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_r2b_reference.json
