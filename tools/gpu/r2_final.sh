#!/bin/bash
# round 2, evidence run at the head: the driver's GPU tier (whole -m gpu suite + smoke), both bench arms at their defaults,
# the launch list of the bench command and one `ncu --set full` capture per kernel.  TAG names the outputs under gpurun_out/.
TAG=${TAG:-r2c}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/gputests_${TAG}.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke_${TAG}.log
s=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
e=$(date +%s); echo "reference arm: $((e-s)) s"
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
f=$(date +%s); echo "cuda arm: $((f-e)) s"
tail -2 gpurun_out/bench_${TAG}.err
cut -c1-400 gpurun_out/bench_${TAG}.json
TAG=$TAG bash tools/gpu/r2_ncu_final.sh 2>&1 | tail -14
