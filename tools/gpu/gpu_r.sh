#!/bin/bash
# Round-1 v8 evidence at the current head: parity, smoke, both arms, launch list, full ncu captures per kernel
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; tail -c 300 gpurun_out/bench_v8.json; tail -3 gpurun_out/bench_v8.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_v8_reference.json 2>> gpurun_out/bench_v8.err; cut -c1-300 gpurun_out/bench_v8_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/launches_v8.log 2>&1
for k in preview_fused herdt_qp mpc_pre mpc_post pldp zmpdisc fcals dimitrov; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 2 -c 1 -f -o gpurun_out/prof_${k}_v8 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --dimitrov-walks 2048 > gpurun_out/ncu_${k}.log 2>&1
tail -1 gpurun_out/ncu_${k}.log
done
ls -la gpurun_out
