#!/bin/bash
# r01l: release fused into the batch's last kernel, chunked rank kernel (edges per thread A/B)
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== products"
CONFIGS="LG_RANK_ITEMS=12
LG_RANK_ITEMS=16
LG_RANK_ITEMS=8
LG_RANK_ITEMS=12" bash scripts/gpu_ab.sh
echo "== ukunion"
CONFIGS="LG_RANK_ITEMS=16
LG_RANK_ITEMS=12
LG_RANK_ITEMS=8
LG_RANK_ITEMS=16 LG_PDL=1" BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
echo "== products serial"
CONFIGS="LG_RANK_ITEMS=12" BENCH_ARGS='--inflight 1 --overlap 0' bash scripts/gpu_ab.sh
echo "== products 3 in flight"
CONFIGS="LG_RANK_ITEMS=12" BENCH_ARGS='--inflight 3' bash scripts/gpu_ab.sh
