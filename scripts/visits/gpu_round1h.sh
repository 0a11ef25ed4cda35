#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
j=json.loads(line);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()}, j['config']['num_nodes'], j['config']['num_edges'], j['roofline']['rows_per_step'], j.get('cpu_baseline',{}).get('value'))"; }
echo "== ref gpu kernels"; timeout 600 python scripts/ref_gpu_baseline.py 2> gpurun_out/ref_gpu.err | tee gpurun_out/ref_gpu_baseline.json; tail -3 gpurun_out/ref_gpu.err
echo "== bench products (full, with cpu baseline)"; timeout 900 python bench.py > gpurun_out/bench_h_products.json 2> gpurun_out/bench_h.err; show gpurun_out/bench_h_products.json products; tail -2 gpurun_out/bench_h.err
echo "== bench paper100m"; timeout 900 python bench.py --workload paper100m --steps 100 > gpurun_out/bench_h_paper.json 2> gpurun_out/bench_h.err; show gpurun_out/bench_h_paper.json paper100m; tail -3 gpurun_out/bench_h.err
echo "== bench ukunion"; timeout 1200 python bench.py --workload ukunion --steps 100 > gpurun_out/bench_h_uk.json 2> gpurun_out/bench_h.err; show gpurun_out/bench_h_uk.json ukunion; tail -3 gpurun_out/bench_h.err
nvidia-smi --query-gpu=memory.used --format=csv
echo "== ncu launches (serial schedule)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/ncu_l2.log 2>&1
echo "== ncu full (gather_tma, sample_hop, rank_relabel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_tma|sample_hop|rank_relabel" -s 25 -c 10 -o gpurun_out/prof_final -f \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/ncu_f4.log 2>&1
ls -la gpurun_out | tail -8
