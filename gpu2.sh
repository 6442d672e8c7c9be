#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_host_cpp_gpu.py -m gpu -x -q 2>&1 | tail -25
