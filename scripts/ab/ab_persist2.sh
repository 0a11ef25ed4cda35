run() { env $ENVV python bench.py --steps 100 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
for wl in ukunion products; do
ENVV="A=0" run --workload $wl
ENVV="LG_RELABEL_CTAS_PER_SM=4" run --workload $wl
ENVV="LG_RELABEL_CTAS_PER_SM=2" run --workload $wl
ENVV="LG_GATHER_DYNAMIC=1" run --workload $wl
ENVV="LG_GATHER_DYNAMIC=1 LG_RELABEL_CTAS_PER_SM=4" run --workload $wl
ENVV="LG_TMA_ROWS=16 LG_RELABEL_CTAS_PER_SM=4" run --workload $wl
ENVV="A=1" run --workload $wl
done
