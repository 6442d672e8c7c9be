#!/bin/bash
# round 2: compute-sanitizer over the preview kernels (recursive sum in every CTA shape, 256-bit stores, cp.async ring) and the
# dense QP / Wieber kernels added this round: memcheck, then racecheck (shared-memory hazards) on the preview tests
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck_r2.log python -m pytest tests/test_preview.py tests/test_preview_ref.py tests/test_two_stage.py tests/test_qld_gpu.py -m gpu -q -x -k "not full_size" 2>&1 | tail -4
echo "memcheck exit: $?"
tail -6 gpurun_out/memcheck_r2.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 --log-file gpurun_out/racecheck_r2.log python -m pytest tests/test_preview.py -m gpu -q -x -k "recursive_and_direct or ragged or position_only" 2>&1 | tail -4
echo "racecheck exit: $?"
grep -c "Race reported\|hazard" gpurun_out/racecheck_r2.log; tail -8 gpurun_out/racecheck_r2.log | cut -c1-220
