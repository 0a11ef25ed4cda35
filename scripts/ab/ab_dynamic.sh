#!/bin/bash
# gather tile order: chunks of consecutive tiles per CTA, static (LG_GATHER_CHUNK) or claimed from a counter one chunk ahead
# (LG_GATHER_DYNAMIC = tiles per claim), against one tile per round-robin step; pipelined, 1 GPU
mkdir -p gpurun_out
LG_GATHER_CHUNK=4 timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_gather_gpu.py -x -q -m gpu -k "dynamic_gather or stream_schedules or bit_exact" 2>&1 | tail -3
run() { env $ENVV python bench.py --steps 150 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
for wl in ukunion products; do
ENVV="A=0" run --workload $wl
ENVV="LG_GATHER_CHUNK=2" run --workload $wl
ENVV="LG_GATHER_CHUNK=4" run --workload $wl
ENVV="LG_GATHER_CHUNK=8" run --workload $wl
ENVV="LG_GATHER_CHUNK=16" run --workload $wl
ENVV="LG_GATHER_DYNAMIC=16" run --workload $wl
ENVV="LG_GATHER_DYNAMIC=32" run --workload $wl
ENVV="LG_GATHER_DYNAMIC=64" run --workload $wl
done
