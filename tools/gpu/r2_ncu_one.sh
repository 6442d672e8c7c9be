#!/bin/bash
# one ncu --set full capture: K=<kernel regex> SKIP=<launches to skip> TAG=<name> EXTRA="<bench flags>"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${K} -s ${SKIP:-0} -c ${CNT:-1} -f -o gpurun_out/prof_${TAG} python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --passes-per-step 2 --e2e-passes 2 --no-sweep $EXTRA > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log | cut -c1-200
