run() { env $ENVV python bench.py --steps 100 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
python -m pytest tests/test_sampler_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -1
LG_SAMPLE_CTAS_PER_SM=3 LG_RANK_CTAS_PER_SM=3 python -m pytest tests/test_full_size_gpu.py tests/test_large_graph_gpu.py -m gpu -x -q 2>&1 | tail -1
for wl in ukunion products; do
ENVV="A=0" run --workload $wl
ENVV="LG_SAMPLE_CTAS_PER_SM=3" run --workload $wl
ENVV="LG_SAMPLE_CTAS_PER_SM=2" run --workload $wl
ENVV="LG_SAMPLE_CTAS_PER_SM=3 LG_RANK_CTAS_PER_SM=3" run --workload $wl
ENVV="LG_SAMPLE_CTAS_PER_SM=3 LG_RANK_CTAS_PER_SM=2" run --workload $wl
ENVV="LG_SAMPLE_CTAS_PER_SM=2 LG_RANK_CTAS_PER_SM=2" run --workload $wl
done
