#!/bin/bash
# LG_L2_HINTS bit 8: the sampler's neighbour reads (random 4-byte reads, one use) with an evict_first policy
run() { env $ENVV timeout 200 python bench.py --steps 150 --warmup 5 --no-extras --no-cpu-baseline --no-parity-check "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
ENVV="LG_L2_HINTS=12" run --workload ukunion
ENVV="LG_L2_HINTS=4" run --workload ukunion
