#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck_mpc.log python -m pytest tests/test_herdt_mpc_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "exit: $?"; tail -4 gpurun_out/memcheck_mpc.log
