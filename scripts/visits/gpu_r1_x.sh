#!/bin/bash
set -u
LG_DEDUP=hash timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu -k hash 2>&1 | tail -3
CONFIGS='X=uk' BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
CONFIGS='LG_DEDUP=hash' bash scripts/gpu_ab.sh
