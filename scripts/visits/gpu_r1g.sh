#!/bin/bash
# r01g: new defaults (8-row gather tiles / 130 KB, PDL per layout, one-wave sampler kernels) + row prefetch A/B
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== products"
CONFIGS="LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0
LG_ROW_PREFETCH=1 LG_PM_FILL_MB=0
LG_ROW_PREFETCH=1 LG_RANK_ITEMS=8
LG_ROW_PREFETCH=1 LG_SAMPLE_MINB=0
LG_ROW_PREFETCH=1 LG_TMA_ROWS=32 LG_GATHER_SMEM_KB=220 LG_TMA_CTAS=4
LG_ROW_PREFETCH=1" bash scripts/gpu_ab.sh
echo "== ukunion"
CONFIGS="LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0
LG_ROW_PREFETCH=1 LG_PDL=1
LG_ROW_PREFETCH=1 LG_GATHER_SMEM_KB=164" BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
echo "== products serial"
CONFIGS="LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0" BENCH_ARGS='--inflight 1 --overlap 0' bash scripts/gpu_ab.sh
echo "== sampler timeline (products, no gather)"
timeout 300 python scripts/trace_sampler.py 2>&1 | grep -v "phase [567]" | tail -40
