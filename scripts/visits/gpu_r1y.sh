#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== products: gather stream priority"
CONFIGS="LG_SIDE_PRIORITY=0
LG_SIDE_PRIORITY=1
LG_SIDE_PRIORITY=0
LG_SIDE_PRIORITY=1" bash scripts/gpu_ab.sh
echo "== ukunion"
CONFIGS="LG_SIDE_PRIORITY=0
LG_SIDE_PRIORITY=1" BENCH_ARGS='--workload ukunion --steps 150' bash scripts/gpu_ab.sh
