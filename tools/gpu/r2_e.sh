#!/bin/bash
# two-stage preview: tests
timeout 900 python -m pytest tests/test_two_stage.py tests/test_preview_ref.py tests/test_preview.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -15
