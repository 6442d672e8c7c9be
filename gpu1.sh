#!/bin/bash
# Round-1 GPU session A: parity tests, bench line, ncu launch list + full capture of the preview kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_v1.json 2> gpurun_out/bench_r1_v1.err; tail -c 3000 gpurun_out/bench_r1_v1.json; tail -5 gpurun_out/bench_r1_v1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v1.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:preview_ -s 6 -c 2 -o gpurun_out/prof_preview_v1 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
