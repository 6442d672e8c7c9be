#!/bin/bash
timeout 900 python -m pytest tests/test_host_cpp_gpu.py -m gpu -q -x 2>&1 | tail -25
