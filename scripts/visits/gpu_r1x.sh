#!/bin/bash
# last check of the round: full GPU suite, smoke, default bench (as the driver runs them)
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/r01f_bench_products.json 2> gpurun_out/r01f_bench.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r01f_bench_products.json') if l.startswith('{')][-1])
print(round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms; e2e', round(j['e2e']['value']/1e6,2), 'server', round(j['e2e_server']['value']/1e6,2), 'frac', round(j['roofline']['frac'],3), 'cpu', round(j['cpu_baseline']['value']/1e6,3), j['clocks'])
PY
