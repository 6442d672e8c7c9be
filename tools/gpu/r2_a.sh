#!/bin/bash
# round 2, step a: parity suite + smoke at the head (preview pins to the reference object, per-tick latency)
mkdir -p gpurun_out
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
