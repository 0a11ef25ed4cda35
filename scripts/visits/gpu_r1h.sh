#!/bin/bash
# r01h: two-level tile prefix (group words) — parity, A/B of row prefetch on top, timeline
set -u
mkdir -p gpurun_out
echo "== pytest gpu (sampler, full size, server)"; timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py tests/test_server_gpu.py -x -q -m gpu 2>&1 | tail -4
echo "== products"
CONFIGS="LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0
LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0" bash scripts/gpu_ab.sh
echo "== ukunion"
CONFIGS="LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0
LG_ROW_PREFETCH=0 LG_PDL=1" BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
echo "== products serial"
CONFIGS="LG_ROW_PREFETCH=1
LG_ROW_PREFETCH=0" BENCH_ARGS='--inflight 1 --overlap 0' bash scripts/gpu_ab.sh
echo "== sampler timeline (products, no gather), prefetch off then on"
LG_ROW_PREFETCH=0 timeout 300 python scripts/trace_sampler.py 2>&1 | grep -v "phase [567]" | tail -28
LG_ROW_PREFETCH=1 timeout 300 python scripts/trace_sampler.py 2>&1 | grep -v "phase [567]" | tail -28
