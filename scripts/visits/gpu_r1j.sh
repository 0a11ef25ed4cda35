#!/bin/bash
# r01j: overlap probe with the new gather tiles, UK-Union timeline, gather budget around the default
set -u
mkdir -p gpurun_out
echo "== parity"; timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== products"
CONFIGS="LG_L2_HINTS=4
LG_GATHER_SMEM_KB=100
LG_GATHER_SMEM_KB=164
LG_L2_HINTS=4" bash scripts/gpu_ab.sh
echo "== products, 3 in flight"
CONFIGS="LG_L2_HINTS=4" BENCH_ARGS='--inflight 3' bash scripts/gpu_ab.sh
echo "== overlap probe (new gather tiles)"
N_ITERS=100 timeout 300 python scripts/overlap_probe.py 2>&1 | tail -16
echo "== sampler timeline UK-Union"
timeout 600 python scripts/trace_sampler.py --workload ukunion 2>&1 | grep -v "phase [567]" | tail -28
