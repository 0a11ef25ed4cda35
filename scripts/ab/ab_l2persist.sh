for wl in products ukunion; do for mb in 0 32 64 96; do
LG_L2_PERSIST_MB=$mb python bench.py --workload $wl --steps 100 --warmup 5 --no-extras --no-parity-check --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); b=j['breakdown_ms']
print('$wl persist=$mb', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms  e2e',round(j['e2e']['value']/1e6,2), 'frac',round(j['roofline']['frac'],3), {k:round(v,4) for k,v in b.items()})"
done; done
