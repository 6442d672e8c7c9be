#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the solver kernels' parity tests
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 --log-file gpurun_out/racecheck.log python -m pytest tests/test_dimitrov.py tests/test_pldp_gpu.py tests/test_herdt_gpu.py -m gpu -q -x -k "not batch_properties and not full_size" 2>&1 | tail -5
echo "exit: $?"
grep -c "Race reported\|hazard" gpurun_out/racecheck.log; tail -25 gpurun_out/racecheck.log | cut -c1-220
