run() { env $ENVV python bench.py --workload ukunion --steps 100 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
ENVV="A=0" run --inflight 2
ENVV="A=0" run --inflight 3
ENVV="A=0" run --inflight 4
ENVV="A=0" run --inflight 6
ENVV="LG_GATHER_SMEM_KB=100" run
ENVV="LG_GATHER_SMEM_KB=160" run
ENVV="LG_GATHER_SMEM_KB=200" run
ENVV="LG_TMA_ROWS=16" run
ENVV="LG_TMA_ROWS=32 LG_GATHER_SMEM_KB=200" run
ENVV="LG_TMA_STAGES=4" run
ENVV="LG_PDL=1" run
ENVV="LG_L2_HINTS=0" run
ENVV="LG_L2_HINTS=7" run
