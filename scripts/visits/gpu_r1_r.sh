#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
CONFIGS='LG_DEDUP=dense
LG_DEDUP=hash' bash scripts/gpu_ab.sh
CONFIGS='LG_DEDUP=dense
LG_DEDUP=hash' BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
