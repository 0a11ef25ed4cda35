#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (server + sampler + gather)"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu2.log
echo "== sweep"; timeout 1500 python scripts/gather_sweep.py 2>&1 | tee gpurun_out/gather_sweep.txt
echo "== bench overlap on"; timeout 600 python bench.py --no-cpu-baseline --no-server-e2e > gpurun_out/bench_overlap.json 2> gpurun_out/bench_overlap.err; python -c "
import json;j=json.load(open('gpurun_out/bench_overlap.json'));print(j['value'],j['ms_per_step'],j['e2e']['value'],j['roofline']['frac'],j['breakdown_ms'])"
echo "== bench overlap off"; timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --no-overlap > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err; python -c "
import json;j=json.load(open('gpurun_out/bench_nooverlap.json'));print(j['value'],j['ms_per_step'],j['e2e']['value'],j['roofline']['frac'])"
