#!/bin/bash
timeout 1200 python -m pytest tests/test_wieber.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -30
