#!/bin/bash
# r01p: chain_kernel (one persistent launch for the dense sampler chain) — parity, A/B
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu (sampler + full size)"; timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "== pytest gpu (rest)"; timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_sampler_gpu.py --deselect tests/test_full_size_gpu.py 2>&1 | tail -3
echo "== products"
CONFIGS="LG_CHAIN=1
LG_CHAIN=0
LG_CHAIN=1
LG_CHAIN=0" BENCH_TIMEOUT=200 bash scripts/gpu_ab.sh
echo "== products serial"
CONFIGS="LG_CHAIN=1
LG_CHAIN=0" BENCH_ARGS='--inflight 1 --overlap 0' BENCH_TIMEOUT=200 bash scripts/gpu_ab.sh
echo "== products 3 in flight"
CONFIGS="LG_CHAIN=1
LG_CHAIN=0" BENCH_ARGS='--inflight 3' BENCH_TIMEOUT=200 bash scripts/gpu_ab.sh
