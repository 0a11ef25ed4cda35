#!/bin/bash
# r01d: PDL chain + wave-quantisation knobs of the sampler (LG_SAMPLE_MINB / LG_RANK_ITEMS / LG_RANK_MINB / LG_PM_FILL_MB)
set -u
mkdir -p gpurun_out
echo "== parity, defaults (PDL on)"
timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu > gpurun_out/r01d_parity_default.log 2>&1
rc=$?; tail -3 gpurun_out/r01d_parity_default.log
if [ $rc -ne 0 ]; then echo "PARITY FAILED with PDL on (rc=$rc): retrying with LG_PDL=0"; tail -40 gpurun_out/r01d_parity_default.log
  LG_PDL=0 timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -15; exit 1; fi
echo "== parity, tuned"
LG_SAMPLE_MINB=6 LG_RANK_ITEMS=12 LG_PM_FILL_MB=16 timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== parity, tuned 2 (hashed: minb 6)"
LG_SAMPLE_MINB=6 LG_RANK_ITEMS=12 LG_RANK_MINB=6 LG_PDL=0 timeout 600 python -m pytest tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== products sweep"
CONFIGS='LG_PDL=0
LG_PDL=1
LG_PDL=1 LG_SAMPLE_MINB=6
LG_PDL=1 LG_SAMPLE_MINB=6 LG_RANK_ITEMS=12
LG_PDL=1 LG_SAMPLE_MINB=6 LG_RANK_ITEMS=12 LG_PM_FILL_MB=16
LG_PDL=1 LG_SAMPLE_MINB=6 LG_RANK_ITEMS=16 LG_PM_FILL_MB=16
LG_PDL=0 LG_SAMPLE_MINB=6 LG_RANK_ITEMS=12 LG_PM_FILL_MB=16' bash scripts/gpu_ab.sh
echo "== products, serial (no overlap, 1 in flight)"
CONFIGS='LG_PDL=0
LG_PDL=1 LG_SAMPLE_MINB=6 LG_RANK_ITEMS=12 LG_PM_FILL_MB=16' BENCH_ARGS='--inflight 1 --overlap 0' bash scripts/gpu_ab.sh
echo "== ukunion sweep"
CONFIGS='LG_PDL=0
LG_PDL=1 LG_SAMPLE_MINB=6
LG_PDL=1 LG_SAMPLE_MINB=6 LG_RANK_MINB=6 LG_RANK_ITEMS=12' BENCH_ARGS='--workload ukunion --steps 100' bash scripts/gpu_ab.sh
