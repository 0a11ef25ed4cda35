run() { env $ENVV python bench.py --steps 100 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
ENVV="A=0" run --workload ukunion
ENVV="LG_RANK_ITEMS=8" run --workload ukunion
ENVV="LG_RANK_ITEMS=12" run --workload ukunion
ENVV="LG_RANK_ITEMS=8 LG_RANK_CTAS_PER_SM=4" run --workload ukunion
ENVV="LG_RANK_ITEMS=8 LG_SAMPLE_TILE=128" run --workload ukunion
ENVV="LG_RANK_ITEMS=8" run --workload products
ENVV="LG_RANK_ITEMS=4" run --workload products
ENVV="A=0" run --workload products
