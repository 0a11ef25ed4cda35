#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/trace_sampler.py 2>&1 | tail -60 | tee gpurun_out/trace_summary.txt
