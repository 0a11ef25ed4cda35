#!/bin/bash
set -u
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python scripts/overlap_probe.py 2>&1 | grep -E "^alone|^together" | head -3
CONFIGS='LG_CARVEOUT=100
LG_CARVEOUT=-1
LG_CARVEOUT=100
LG_CARVEOUT=100 LG_TMA_CTAS=5' bash scripts/gpu_ab.sh
CONFIGS='LG_CARVEOUT=100
LG_CARVEOUT=-1' BENCH_ARGS='--inflight 1' bash scripts/gpu_ab.sh
