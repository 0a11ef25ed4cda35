#!/bin/bash
# short end-of-round visit: GPU tests, smoke, default bench, ncu launch list (no full capture, no reference arm)
TAG=${TAG:-r02v}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench_default.json
echo "== launch list (ncu)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gather_|sample_hop|rank_kernel|relabel_kernel|batch_generate|release_kernel|seed_local" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --scale 0.25 --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-extras --no-parity-check --inflight 1 --overlap 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 300 python bench.py --scale 0.25 --steps 50 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/${TAG}_bench_scaled.json 2>> gpurun_out/${TAG}_bench.err
