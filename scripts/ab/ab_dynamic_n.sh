#!/bin/bash
# usage: ab_dynamic_n.sh <gpus>: the gather's tile order on the NVLink-bound layout (Kc=1, Kg=N)
n=$1; port=29800
run() { port=$((port+1)); env $ENVV python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 100 --warmup 5 --no-extras --no-parity-check "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; m=r['hit_mix']
print('$ENVV $*', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms e2e', round(j['e2e']['value']/1e6,2),'| gather',round(r['gather_ms_per_step'],4),'nvlink GB/s',round(m['link_GBps_achieved']['nvlink'],1),'frac',round(r['frac'],3))"; }
ENVV="A=0" run
ENVV="LG_GATHER_DYNAMIC=4" run
ENVV="LG_GATHER_DYNAMIC=4 LG_GATHER_STATIC_PCT=75" run
ENVV="LG_GATHER_DYNAMIC=2" run
