#!/bin/bash
# quick visit: parity tests, sampler timeline, one bench line
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
j=json.loads(line);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_q.log
timeout 600 python scripts/trace_sampler.py 2>&1 | tail -45 | tee gpurun_out/trace_summary.txt
timeout 600 python bench.py --no-cpu-baseline --no-server-e2e ${BENCH_ARGS:-} > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err; show gpurun_out/bench_q.json "products"
