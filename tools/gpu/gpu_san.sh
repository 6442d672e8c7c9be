#!/bin/bash
# compute-sanitizer memcheck over the GPU parity tests (small batches)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck.log python -m pytest tests/test_dimitrov.py tests/test_pldp_gpu.py tests/test_zmpdisc.py tests/test_preview.py tests/test_herdt_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "exit: $?"
tail -15 gpurun_out/memcheck.log
