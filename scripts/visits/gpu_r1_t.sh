#!/bin/bash
set -u
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled; grep -i huge /proc/meminfo | head -5; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m | head -5
CONFIGS='LG_HOST_HUGEPAGES=0
LG_HOST_HUGEPAGES=1' BENCH_ARGS='--workload clueweb --scale 0.05 --topo host --topo-cache-ratio 0.1 --cache-ratio 0.2 --steps 30' bash scripts/gpu_ab.sh
CONFIGS='LG_HOST_HUGEPAGES=0
LG_HOST_HUGEPAGES=1' BENCH_ARGS='--cache-ratio 0.2 --steps 50' bash scripts/gpu_ab.sh
