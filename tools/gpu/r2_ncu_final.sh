#!/bin/bash
# round 2 (final): launch list of the bench command + one `ncu --set full` capture per kernel
mkdir -p gpurun_out
TAG=${TAG:-r2b}
COMMON="--steps 2 --warmup 1 --cpu-seconds 0.5 --passes-per-step 2 --e2e-passes 2 --no-sweep"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $COMMON --wieber-walks 64 > gpurun_out/launches_${TAG}.log 2>&1
for k in ${KERNELS:-preview_rec_warp preview_fused herdt_qp pldp zmpdisc fcals dimitrov qld wieber_pre}; do
  extra=""; skip=3
  case $k in
    preview_rec_warp|preview_fused) extra="--no-herdt --no-pldp --no-kajita --no-dimitrov --no-wieber";;
    herdt_qp) extra="--no-pldp --no-kajita --no-dimitrov --no-wieber";;
    pldp) extra="--no-herdt --no-kajita --no-dimitrov --no-wieber";;
    zmpdisc) extra="--no-herdt --no-pldp --no-dimitrov --no-wieber"; skip=2;;
    fcals|dimitrov) extra="--no-herdt --no-pldp --no-kajita --no-wieber --dimitrov-walks 2048"; skip=1;;
    qld|wieber_pre) extra="--no-herdt --no-pldp --no-kajita --no-dimitrov --wieber-walks 296"; skip=400;;
  esac
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s $skip -c 1 -f -o gpurun_out/prof_${k}_${TAG} python bench.py $COMMON $extra > gpurun_out/ncu_${k}.log 2>&1
  tail -1 gpurun_out/ncu_${k}.log | cut -c1-160
done
ls -la gpurun_out | grep ${TAG}
