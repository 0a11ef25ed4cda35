#!/bin/bash
# compute-sanitizer passes over a small slice of the parity suite (memcheck + racecheck + synccheck)
set -u
mkdir -p gpurun_out
SEL='tests/test_sampler_gpu.py::test_edge_cases tests/test_sampler_gpu.py::test_small_table_forces_collisions tests/test_gather_gpu.py::test_gather_without_cache_reads_backing'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -x -q 2>&1 | tail -6 | tee gpurun_out/sanitize_$tool.log
done
