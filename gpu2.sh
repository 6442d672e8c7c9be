#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pldp_gpu.py -m gpu -x -q -s 2>&1 | tail -30
