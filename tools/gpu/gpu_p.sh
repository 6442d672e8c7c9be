#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dimitrov_kernel -s 1 -c 1 -f -o gpurun_out/prof_dimitrov_v5 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.2 --no-herdt --no-pldp --no-kajita --dimitrov-walks 8192 > gpurun_out/ncu_dimitrov.log 2>&1
tail -2 gpurun_out/ncu_dimitrov.log
