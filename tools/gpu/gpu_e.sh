#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:preview_fused -s 3 -c 1 -f -o gpurun_out/prof_preview_fused_v3 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --no-herdt --no-pldp > gpurun_out/ncu_full1.log 2>&1
tail -3 gpurun_out/ncu_full1.log
ls -la gpurun_out
