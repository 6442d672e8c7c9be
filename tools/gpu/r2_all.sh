#!/bin/bash
# the driver's GPU tier: whole GPU suite, smoke
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
