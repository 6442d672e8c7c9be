#!/bin/bash
# Full GPU session: parity tests, smoke, both bench arms, ncu launch list + full captures of the four kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_v2.json 2> gpurun_out/bench_r1_v2.err; tail -c 600 gpurun_out/bench_r1_v2.json; tail -3 gpurun_out/bench_r1_v2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_v2_reference.json 2>> gpurun_out/bench_r1_v2.err; cat gpurun_out/bench_r1_v2_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v2.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:preview_fused -s 3 -c 1 -o gpurun_out/prof_preview_fused_v2 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --no-herdt --no-pldp > gpurun_out/ncu_full1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:herdt_qp_kernel -s 3 -c 1 -o gpurun_out/prof_herdt_qp_v2 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --herdt-periods 2 --no-pldp > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:herdt_mpc_kernel -s 18 -c 1 -o gpurun_out/prof_herdt_mpc_v2 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --herdt-periods 2 --no-pldp > gpurun_out/ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pldp_kernel -s 3 -c 1 -o gpurun_out/prof_pldp_v2 python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --no-herdt > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out | tail -20
