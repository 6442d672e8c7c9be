#!/bin/bash
# GPU session: parity tests, bench line, ncu launch list + full capture of the fused preview kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:preview_fused -s 3 -c 1 -o gpurun_out/prof_preview_fused python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
