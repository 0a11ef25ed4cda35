#!/bin/bash
# compute-sanitizer passes over a small slice of the parity suite (memcheck + racecheck + synccheck), both layouts of the position map
set -u
mkdir -p gpurun_out
SEL='tests/test_sampler_gpu.py::test_edge_cases tests/test_sampler_gpu.py::test_position_map_is_released_between_batches tests/test_gather_gpu.py::test_gather_without_cache_reads_backing tests/test_blocks.py::test_blocks_of_a_sampled_batch tests/test_sampler_gpu.py::test_lazy_relabel_op_by_op tests/test_sampler_gpu.py::test_chain_kernel_matches_oracle'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${TAG:-r01e}_sanitize_$tool.log
done
