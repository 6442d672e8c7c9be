#!/bin/bash
# which CTA shape of the recursive preview kernel wins at which batch size (the launch heuristic of wgi_preview_launch_range)
mkdir -p gpurun_out
for w in ${WALKS:-256 512 1024 1536 2048 3072}; do
  for sh in ${SHAPE_LIST:-0 1 2 3}; do
    WG_PREVIEW_SHAPE=$sh timeout 300 python bench.py --walks $w --steps 5 --warmup 3 --no-herdt --no-pldp --no-kajita --no-dimitrov --no-wieber --no-sweep --passes-per-step 24 --e2e-passes 1 --cpu-seconds 0.1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=list(d['kernels'].values())[0]
print('walks $w shape $sh: %.2f G steps/s, kernel %.4f ms' % (d['value']/1e9, k['avg_ms']))"
  done
done
