#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
j=json.loads(line);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu9.log
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-server-e2e > gpurun_out/bench_j_$i.json 2> gpurun_out/bench_j.err; show gpurun_out/bench_j_$i.json "products run$i"; done
timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --inflight 4 > gpurun_out/bench_j_if4.json 2> gpurun_out/bench_j.err; show gpurun_out/bench_j_if4.json "products inflight4"
timeout 900 python bench.py --workload ukunion --steps 100 > gpurun_out/bench_j_uk.json 2> gpurun_out/bench_j.err; show gpurun_out/bench_j_uk.json ukunion
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sample_hop|rank_relabel" -s 16 -c 4 -o gpurun_out/prof_sampler2 -f \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/ncu_f5.log 2>&1
