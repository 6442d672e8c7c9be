#!/bin/bash
# Herdt whole-datref parity on the device (closed-loop kernel + host C++ mirror)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_herdt_mpc_gpu.py tests/test_host_cpp_gpu.py -m gpu -x -q -s 2>&1 | tail -25
