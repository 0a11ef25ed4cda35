#!/bin/bash
# end-of-round visit: what the driver runs (GPU tests, smoke, both bench arms) + the ncu evidence for profiles/
# usage: TAG=r02 bash scripts/gpu_final.sh        (one GPU)
#   SHORT=1     skip the reference arm and the ncu --set full capture (tests, smoke, default bench, launch list)
#   SANITIZE=1  add compute-sanitizer memcheck + racecheck over the gather's tensor-copy and claimed-chunk paths
set -u
TAG=${TAG:-r02}
SCALE=${SCALE:-0.25}   # ncu runs: UK-Union shape scaled so that kernel replay does not have to save/restore 100 GB
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_default.json
if [ -z "${SHORT:-}" ]; then
  echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench_reference.json
fi
echo "== launch list (ncu)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gather_|sample_hop|rank_kernel|relabel_kernel|batch_generate|release_kernel|seed_local" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --scale $SCALE --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-extras --no-parity-check --inflight 1 --overlap 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
if [ -z "${SHORT:-}" ]; then
  echo "== full capture (ncu --set full)"
  TAG=$TAG SKIP=${SKIP:-42} COUNT=${COUNT:-16} BENCH_ARGS="--scale $SCALE" bash scripts/gpu_ncu.sh > gpurun_out/${TAG}_ncu_table.txt 2>&1; cat gpurun_out/${TAG}_ncu_table.txt
fi
timeout 300 python bench.py --scale $SCALE --steps 50 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/${TAG}_bench_scaled.json 2>> gpurun_out/${TAG}_bench.err
if [ -n "${SANITIZE:-}" ]; then
  echo "== sanitizer (gather4 / claimed chunks)"
  for tool in memcheck racecheck; do
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "tests/test_gather_gpu.py::test_gather4_tensor_copies[128-2]" "tests/test_gather_gpu.py::test_gather4_tensor_copies[100-0]" "tests/test_sampler_gpu.py::test_dynamic_gather_tiles" -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_sanitize_$tool.log
  done
fi
