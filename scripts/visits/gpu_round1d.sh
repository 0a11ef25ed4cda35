#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu4.log
echo "== sweep v3"; timeout 1500 python scripts/gather_sweep.py 2>&1 | tee gpurun_out/gather_sweep_v3.txt
for ov in 2 1; do
echo "== bench overlap $ov"; timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --overlap $ov > gpurun_out/bench_d_ov$ov.json 2> gpurun_out/bench_d_ov$ov.err; python -c "
import json;j=json.load(open('gpurun_out/bench_d_ov$ov.json'));print(j['value'],j['ms_per_step'],j['e2e']['value'],j['roofline']['frac'],j['breakdown_ms'])"; tail -2 gpurun_out/bench_d_ov$ov.err
done
