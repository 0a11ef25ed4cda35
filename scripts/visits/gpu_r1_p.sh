#!/bin/bash
# L2 eviction-priority hints A/B (LG_L2_HINTS bitmask, see csrc/common.cuh)
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(r['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
LG_L2_HINTS=15 timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for h in ${HINTS:-0 4 3 7 12 15}; do
  LG_L2_HINTS=$h timeout 600 python bench.py --no-cpu-baseline --no-server-e2e ${BENCH_ARGS:-} > gpurun_out/bench_p_$h.json 2> gpurun_out/bench_p_$h.err || tail -5 gpurun_out/bench_p_$h.err
  show gpurun_out/bench_p_$h.json "hints=$h"
done
