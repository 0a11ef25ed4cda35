#!/bin/bash
# r01s: chain opt-in + its test, hashed layout without the per-hop relabel launch, batches in flight
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== ukunion"
CONFIGS="LG_L2_HINTS=4
LG_L2_HINTS=4" BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
CONFIGS="LG_L2_HINTS=4" BENCH_ARGS='--workload ukunion --steps 100 --inflight 3' bash scripts/gpu_ab.sh
echo "== products"
CONFIGS="LG_L2_HINTS=4" bash scripts/gpu_ab.sh
CONFIGS="LG_L2_HINTS=4" BENCH_ARGS='--inflight 3' bash scripts/gpu_ab.sh
CONFIGS="LG_L2_HINTS=4" BENCH_ARGS='--inflight 4' bash scripts/gpu_ab.sh
