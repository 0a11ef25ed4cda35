#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== products: sampler streams at high priority (2 of 3 runners)"
CONFIGS="BENCH_SAMPLER_PRIORITY=0
BENCH_SAMPLER_PRIORITY=1
BENCH_SAMPLER_PRIORITY=0
BENCH_SAMPLER_PRIORITY=1" bash scripts/gpu_ab.sh
