#!/bin/bash
# tile::gather4 tensor copies in the TMA gather (LG_GATHER4=1) against per-row bulk copies; parity first, then 1-GPU A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gather_gpu.py -x -q -m gpu -k "gather4" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
if [ $rc -ne 0 ]; then echo "gather4 parity FAILED (rc $rc): no A/B"; exit 1; fi
if [ -z "${SKIP_SUITE:-}" ]; then LG_GATHER4=1 timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_gather_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3; fi
run() { env $ENVV timeout 300 python bench.py --steps 150 --warmup 5 --no-extras --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()}, 'parity', (j.get('parity_selfcheck') or {}).get('ok'))"; }
ENVV="A=0" run --workload ukunion --no-parity-check
ENVV="LG_GATHER4=1" run --workload ukunion --no-parity-check
ENVV="LG_GATHER4=1 LG_TMA_ROWS=16" run --workload ukunion --no-parity-check
ENVV="LG_GATHER4=1 LG_TMA_ROWS=16 LG_GATHER_SMEM_KB=160" run --workload ukunion --no-parity-check
ENVV="LG_GATHER4=1 LG_TMA_ROWS=32 LG_GATHER_SMEM_KB=200" run --workload ukunion --no-parity-check
ENVV="LG_GATHER4=1 LG_GATHER4_PROMO=3" run --workload ukunion --no-parity-check
ENVV="LG_GATHER4=1 LG_GATHER4_PROMO=0" run --workload ukunion --no-parity-check
ENVV="LG_TMA_ROWS=16" run --workload ukunion --no-parity-check
