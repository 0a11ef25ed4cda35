#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];D=j['config']['feature_dim'];rows=r['rows_per_step'];h=r['hit_mix']['host']
print('$2', round(j['value']/1e6,2),'M seeds/s','rows',int(rows),'host',round(h,3),'gather ms',round(r['gather_ms_per_step'],3),'PCIe GB/s',round(rows*h*4*D/1e9/(r['gather_ms_per_step']*1e-3),1),'N',j['config']['num_nodes'],{k:round(v,3) for k,v in j['breakdown_ms'].items()})"; }
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline --no-server-e2e "$@" > gpurun_out/bench_u_$name.json 2> gpurun_out/bench_u_$name.err || tail -5 gpurun_out/bench_u_$name.err; show gpurun_out/bench_u_$name.json "$name"; }
run clue_ldg --workload clueweb --scale 0.05 --topo host --topo-cache-ratio 0.1 --cache-ratio 0.2 --steps 30 --gather ldg
run clue_hbmtopo --workload clueweb --scale 0.05 --cache-ratio 0.2 --steps 30
run clue_fuse0 --workload clueweb --scale 0.05 --cache-ratio 0.2 --steps 30 --fuse 0
