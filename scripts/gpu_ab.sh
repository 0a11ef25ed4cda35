#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(r['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()},'enq',round(j.get('host_enqueue_ms_per_step',0),4))"; }
i=0
while read -r line; do
  i=$((i+1))
  env $line timeout ${BENCH_TIMEOUT:-400} python bench.py --no-cpu-baseline --no-server-e2e ${BENCH_ARGS:-} > gpurun_out/bench_q_$i.json 2> gpurun_out/bench_q_$i.err || tail -5 gpurun_out/bench_q_$i.err
  show gpurun_out/bench_q_$i.json "$line"
done <<EOL
${CONFIGS}
EOL
