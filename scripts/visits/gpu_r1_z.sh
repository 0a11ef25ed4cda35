#!/bin/bash
set -u
for r in 8 16; do LG_TMA_ROWS=$r timeout 900 python -m pytest tests/test_gather_gpu.py tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -2; done
for r in 32 16 8; do echo "== LG_TMA_ROWS=$r"; LG_TMA_ROWS=$r python scripts/overlap_probe.py 2>&1 | grep -E "^alone|^together|spin 148x32|spin 1184" | head -5; done
CONFIGS='LG_TMA_ROWS=32
LG_TMA_ROWS=16
LG_TMA_ROWS=8
LG_TMA_ROWS=8 LG_TMA_CTAS=5
LG_TMA_ROWS=8 LG_TMA_STAGES=4' bash scripts/gpu_ab.sh
