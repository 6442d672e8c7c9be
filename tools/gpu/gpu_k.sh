#!/bin/bash
# Dimitrov front-to-back pipeline on the device + regression of the PLDP refactor
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dimitrov.py -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/k_dimitrov.log
timeout 900 python -m pytest tests/test_pldp_gpu.py tests/test_host_cpp_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/k_pldp.log
