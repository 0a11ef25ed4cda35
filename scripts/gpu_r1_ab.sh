#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gather_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -4
CONFIGS='X=tma
X=cpa4
LG_CPA_CTAS=3
LG_CPA_CTAS=6
LG_CPA_CTAS=8' 
i=0
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(r['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
for cfg in "X=1 tma" "X=1 cpa" "LG_CPA_CTAS=3 cpa" "LG_CPA_CTAS=6 cpa" "LG_CPA_CTAS=8 cpa"; do
  set -- $cfg
  env $1 timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --gather $2 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err || tail -5 gpurun_out/bench_ab.err
  show gpurun_out/bench_ab.json "$cfg"
done
