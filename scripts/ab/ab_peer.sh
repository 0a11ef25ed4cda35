#!/bin/bash
# usage: ab_peer.sh <gpus>: gather tuning for the NVLink-bound (Kg=N) layout
n=$1; port=29700
run() { port=$((port+1)); env $ENVV python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 100 --warmup 5 --no-extras --no-parity-check 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); r=j['roofline']; m=r['hit_mix']
print('$ENVV', round(j['value']/1e6,2),'M', round(j['ms_per_step'],4),'ms | gather',round(r['gather_ms_per_step'],4),'nvlink GB/s',round(m['link_GBps_achieved']['nvlink'],1),'frac',round(r['frac'],3))"; }
ENVV="A=0" run
ENVV="LG_TMA_STAGES=4" run
ENVV="LG_TMA_STAGES=6" run
ENVV="LG_TMA_STAGES=6 LG_GATHER_SMEM_KB=200 LG_TMA_CTAS=16" run
ENVV="LG_TMA_STAGES=4 LG_GATHER_SMEM_KB=200 LG_TMA_CTAS=16" run
ENVV="LG_TMA_STAGES=3 LG_GATHER_SMEM_KB=200 LG_TMA_CTAS=16" run
ENVV="LG_TMA_ROWS=16 LG_TMA_STAGES=4 LG_GATHER_SMEM_KB=200 LG_TMA_CTAS=16" run
ENVV="LG_TMA_ROWS=32 LG_TMA_STAGES=4 LG_GATHER_SMEM_KB=220 LG_TMA_CTAS=16" run
