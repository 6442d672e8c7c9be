#!/bin/bash
# ncu full capture of the Dimitrov loop kernel + polygon kernel (source-level), launch list of the whole bench
mkdir -p gpurun_out
for k in dimitrov fcals; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 1 -c 1 -f -o gpurun_out/prof_${k}_v4 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.2 --no-herdt --no-pldp --no-kajita --dimitrov-walks 2048 > gpurun_out/ncu_${k}.log 2>&1
tail -2 gpurun_out/ncu_${k}.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/launches_v4.log 2>&1
tail -3 gpurun_out/launches_v4.log | cut -c1-300
ls -la gpurun_out
