#!/bin/bash
# hashed rank kernel: probe continuations in lockstep + straight-line write phase; parity, trace, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py tests/test_ref_kernels_gpu.py tests/test_blocks.py -x -q -m gpu 2>&1 | tail -3
timeout 250 python scripts/trace_sampler.py --workload ukunion --scale 0.25 2>&1 | grep -A6 "rank h2\|sample h2" | grep -v "phase [567]" | cut -c1-130
run() { env $ENVV timeout 300 python bench.py --steps 150 --warmup 5 --no-extras --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()}, 'parity', (j.get('parity_selfcheck') or {}).get('ok'))"; }
ENVV="A=0" run --workload ukunion
ENVV="A=0" run --workload products --no-parity-check
ENVV="A=0" run --workload ukunion --no-parity-check
