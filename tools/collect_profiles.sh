#!/bin/bash
# copy the evidence of a tools/gpu/r2_final.sh run from gpurun_out/ (scratch) into profiles/ (tracked):
#   bash tools/collect_profiles.sh <TAG> <NAME>     e.g. r2c r2
TAG=${1:-r2c}; NAME=${2:-r2}
set -e
tail -1 gpurun_out/bench_${TAG}.json > profiles/${NAME}_bench.json
tail -1 gpurun_out/bench_${TAG}_reference.json > profiles/${NAME}_bench_reference.json
cp gpurun_out/launches_${TAG}.csv profiles/${NAME}_launches.csv
cp gpurun_out/gputests_${TAG}.log profiles/${NAME}_gputests.log
cp gpurun_out/smoke_${TAG}.log profiles/${NAME}_smoke.log
reps=$(ls gpurun_out/prof_*_${TAG}.ncu-rep)
{ echo "# ncu --set full summaries, round 2 (captures: gpurun_out/prof_*_${TAG}.ncu-rep; tools/ncu_summary.py)"; echo;
  python tools/ncu_summary.py $reps; } > profiles/${NAME}_ncu_summary.md
python tools/ncu_summary.py --traffic 4096 profiles/ncu_traffic.json gpurun_out/prof_preview_rec_warp_${TAG}.ncu-rep gpurun_out/prof_preview_fused_${TAG}.ncu-rep gpurun_out/prof_zmpdisc_${TAG}.ncu-rep gpurun_out/prof_pldp_${TAG}.ncu-rep > /dev/null
ls -la profiles | grep ${NAME}_
