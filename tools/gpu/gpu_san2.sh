#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/memcheck_v8.log python -m pytest tests/test_herdt_mpc_gpu.py tests/test_zmpdisc.py tests/test_herdt_gpu.py tests/test_pldp_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "exit: $?"; tail -3 gpurun_out/memcheck_v8.log
