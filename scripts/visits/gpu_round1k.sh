#!/bin/bash
set -u
mkdir -p gpurun_out
for ins in 0 1 2; do
LG_SAMPLE_INS=$ins timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sample_hop|rank_relabel" -s 12 -c 12 --csv --log-file gpurun_out/l_ins$ins.csv python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > /dev/null 2>&1
python - <<PY
import csv
rows=[l for l in open('gpurun_out/l_ins$ins.csv') if not l.startswith('==')]
from collections import defaultdict
a=defaultdict(list)
for r in csv.DictReader(rows): a[r['Kernel Name'][:60]+r['Grid Size']].append(float(r['Metric Value'])/1e3)
print('INS=$ins', {k:round(sum(v)/len(v),1) for k,v in a.items()})
PY
LG_SAMPLE_INS=$ins timeout 600 python bench.py --no-cpu-baseline --no-server-e2e > gpurun_out/bench_k_$ins.json 2>/dev/null; python -c "
import json;j=json.loads([l for l in open('gpurun_out/bench_k_$ins.json') if l.startswith('{')][-1]);print('INS=$ins',round(j['value']/1e6,2),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"
done
