#!/bin/bash
# r01i: rank writes agg_src / relabel reads agg_src[p_first]; RED pre-check A/B
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== pytest gpu (sampler, full) with RED pre-check"; LG_RED_PRECHECK=1 timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== products"
CONFIGS="LG_RED_PRECHECK=0
LG_RED_PRECHECK=1
LG_RED_PRECHECK=0
LG_RED_PRECHECK=1
LG_RED_PRECHECK=0 LG_RANK_ITEMS=8" bash scripts/gpu_ab.sh
echo "== ukunion"
CONFIGS="LG_RED_PRECHECK=0
LG_RED_PRECHECK=0" BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
echo "== products serial"
CONFIGS="LG_RED_PRECHECK=0
LG_RED_PRECHECK=1" BENCH_ARGS='--inflight 1 --overlap 0' bash scripts/gpu_ab.sh
echo "== sampler timeline (products, no gather), precheck off then on"
timeout 300 python scripts/trace_sampler.py 2>&1 | grep -v "phase [567]" | grep -A8 "h2"
LG_RED_PRECHECK=1 timeout 300 python scripts/trace_sampler.py 2>&1 | grep -v "phase [567]" | grep -A8 "sample h2"
