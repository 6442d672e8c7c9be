#!/bin/bash
# round 2: one-warp-per-trajectory recursive preview kernel - GPU parity tests of everything that runs through the preview kernels
# under the CTA shape(s) in SHAPES (WG_PREVIEW_SHAPE 0: 64 x 8, 1: 128 x 4, 2: one warp), then the headline leg per shape
mkdir -p gpurun_out
for sh in ${TEST_SHAPES:-2}; do
  WG_PREVIEW_SHAPE=$sh timeout 900 python -m pytest tests/test_preview.py tests/test_preview_ref.py tests/test_two_stage.py tests/test_zmpdisc.py tests/test_host_cpp_gpu.py -m gpu -x -q 2>&1 | tail -4
done
for sh in ${SHAPES:-0 2}; do
  WG_PREVIEW_SHAPE=$sh timeout 600 python bench.py --steps 5 --warmup 3 --no-herdt --no-pldp --no-kajita --no-dimitrov --no-wieber --no-sweep --passes-per-step 24 --e2e-passes 4 --cpu-seconds 0.5 > gpurun_out/bench_rec_s$sh.json 2> gpurun_out/bench_rec_s$sh.err
  tail -3 gpurun_out/bench_rec_s$sh.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_rec_s$sh.json'))
print("shape $sh value %.3e ms_per_pass %.4f" % (d['value'], d['ms_per_step']/d['config']['passes_per_step']), "frac %.3f" % d['roofline']['frac'], "e2e", d['e2e']['value'], d['e2e']['com_position_only']['value'])
PY
done
