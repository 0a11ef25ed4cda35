run() { env "$@" python bench.py --workload ukunion --steps 100 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"; }
run LG_SAMPLE_MINB_HASHED=0
run LG_SAMPLE_MINB_HASHED=5
run LG_SAMPLE_MINB_HASHED=6
run LG_SAMPLE_MINB_HASHED=0 LG_SAMPLE_TILE=128
run LG_SAMPLE_MINB_HASHED=6 LG_PDL=1
run LG_SAMPLE_MINB_HASHED=6 LG_RED_PRECHECK=3
