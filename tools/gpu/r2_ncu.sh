#!/bin/bash
# round 2: launch list + one `ncu --set full` capture per kernel of the bench command (names: $KERNELS or all)
mkdir -p gpurun_out
TAG=${TAG:-r2a}
KERNELS=${KERNELS:-"preview_fused herdt_qp mpc_pre mpc_post pldp zmpdisc fcals dimitrov"}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --passes-per-step 2 --e2e-passes 2 --no-sweep > gpurun_out/launches_${TAG}.log 2>&1
for k in $KERNELS; do
  extra=""
  case $k in
    preview_fused) extra="--no-herdt --no-pldp --no-kajita --no-dimitrov";;
    herdt_qp|mpc_pre|mpc_post) extra="--no-pldp --no-kajita --no-dimitrov";;
    pldp) extra="--no-herdt --no-kajita --no-dimitrov";;
    zmpdisc) extra="--no-herdt --no-pldp --no-dimitrov";;
    fcals|dimitrov) extra="--no-herdt --no-pldp --no-kajita --dimitrov-walks 2048";;
  esac
  cnt=1; [ $k = pldp ] && cnt=2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 3 -c $cnt -f -o gpurun_out/prof_${k}_${TAG} python bench.py --steps 2 --warmup 1 --cpu-seconds 0.5 --passes-per-step 2 --e2e-passes 2 --no-sweep $extra > gpurun_out/ncu_${k}.log 2>&1
  tail -1 gpurun_out/ncu_${k}.log | cut -c1-200
done
ls -la gpurun_out | tail -12
