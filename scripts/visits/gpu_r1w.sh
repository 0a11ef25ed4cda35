#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== new parity test"; timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== ukunion: pre-check before the hashed atomicMin"
CONFIGS="LG_RED_PRECHECK=1
LG_RED_PRECHECK=3
LG_RED_PRECHECK=1
LG_RED_PRECHECK=3" BENCH_ARGS='--workload ukunion --steps 150' bash scripts/gpu_ab.sh
