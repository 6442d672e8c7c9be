#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:herdt_qp_kernel -s 3 -c 1 -o gpurun_out/prof_herdt_qp python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --herdt-periods 2 > gpurun_out/ncu_qp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:herdt_mpc_kernel -s 18 -c 1 -o gpurun_out/prof_herdt_mpc python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --herdt-periods 2 > gpurun_out/ncu_mpc.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu_qp.log gpurun_out/ncu_mpc.log
