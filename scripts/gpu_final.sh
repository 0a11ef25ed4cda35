#!/bin/bash
# end-of-round visit: what the driver runs (GPU tests, smoke, both bench arms) + the ncu evidence for profiles/
set -u
TAG=${TAG:-r01e}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/${TAG}_bench_products.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench_products.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/${TAG}_bench_serial.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
TAG=$TAG bash scripts/gpu_ncu.sh > gpurun_out/${TAG}_ncu_table.txt 2>&1; cat gpurun_out/${TAG}_ncu_table.txt
