#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== products serial: chain CTAs per SM"
CONFIGS="LG_CHAIN=1 LG_CHAIN_CTAS=4
LG_CHAIN=1 LG_CHAIN_CTAS=5
LG_CHAIN=1 LG_CHAIN_CTAS=6
LG_CHAIN=1 LG_CHAIN_CTAS=3
LG_CHAIN=0" BENCH_ARGS='--inflight 1 --overlap 0' BENCH_TIMEOUT=200 bash scripts/gpu_ab.sh
echo "== products pipelined: gather carve-out leaves room for the chain's shared memory"
CONFIGS="LG_CHAIN=1 LG_GATHER_CARVEOUT=72
LG_CHAIN=0 LG_GATHER_CARVEOUT=72
LG_CHAIN=1 LG_GATHER_CARVEOUT=72 LG_GATHER_SMEM_KB=110
LG_CHAIN=1 LG_GATHER_CARVEOUT=86" BENCH_TIMEOUT=200 bash scripts/gpu_ab.sh
echo "== sampler timeline (chain)"
timeout 300 python scripts/trace_sampler.py 2>&1 | grep -v "phase [567]" | tail -28
