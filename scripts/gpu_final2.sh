#!/bin/bash
# end-of-round visit: gpu_final.sh + sanitizer passes over the gather's new paths (tile::gather4, claimed chunks)
TAG=${TAG:-r02u}
bash scripts/gpu_final.sh
echo "== sanitizer (gather4 / dynamic chunks)"
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "tests/test_gather_gpu.py::test_gather4_tensor_copies[128-2]" "tests/test_gather_gpu.py::test_gather4_tensor_copies[100-0]" "tests/test_sampler_gpu.py::test_dynamic_gather_tiles" -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_sanitize_$tool.log
done
