#!/bin/bash
# host tiers: feature misses in pinned host memory, topology in pinned host memory (UVA), in-flight sweep
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(r['frac'],3),'mix',{k:(round(v,3) if isinstance(v,float) else v) for k,v in r['hit_mix'].items() if k in ('local','peer','host','bound','frac_of_mix_roofline')},{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
run() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline --no-server-e2e "$@" > gpurun_out/bench_o_$name.json 2> gpurun_out/bench_o_$name.err || tail -5 gpurun_out/bench_o_$name.err; show gpurun_out/bench_o_$name.json "$name"; }
for nf in 1 3 4; do run if$nf --inflight $nf; done
run cr50 --cache-ratio 0.5
run cr20 --cache-ratio 0.2
run cr05 --cache-ratio 0.05
run th00 --topo host
run th10 --topo host --topo-cache-ratio 0.1
run th50 --topo host --topo-cache-ratio 0.5
run th10cr20 --topo host --topo-cache-ratio 0.1 --cache-ratio 0.2
run clue05 --workload clueweb --scale 0.05 --topo host --topo-cache-ratio 0.1 --cache-ratio 0.2 --steps 50
ls -la gpurun_out | head -40
