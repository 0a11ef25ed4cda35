run() { python bench.py "$@" --steps 100 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()}, 'parity', (j.get('parity_selfcheck') or {}).get('all_ranks_ok'))"; }
python -m pytest tests/test_gather_gpu.py -m gpu -x -q 2>&1 | tail -2
run --workload ukunion
run --workload ukunion --no-identity
run --workload products
run --workload products --no-identity
