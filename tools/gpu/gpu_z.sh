#!/bin/bash
timeout 900 python -m pytest tests/test_dimitrov.py -m gpu -q -x -k edge 2>&1 | tail -25
