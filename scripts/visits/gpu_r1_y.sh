#!/bin/bash
set -u
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/overlap_probe.py 2>&1 | grep -v "^rows" | tail -7
CONFIGS='LG_GATHER_DYNAMIC=1
LG_GATHER_DYNAMIC=0
LG_GATHER_DYNAMIC=1' bash scripts/gpu_ab.sh
