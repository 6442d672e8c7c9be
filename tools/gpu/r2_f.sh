#!/bin/bash
timeout 900 python -m pytest tests/test_qld_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -30
