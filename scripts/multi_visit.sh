#!/bin/bash
# usage: scripts/multi_visit.sh <gpus> <tag> <file with one "name args..." bench line per row>
# runs every line under torchrun on <gpus> GPUs; outputs in gpurun_out/<tag>_<name>.{json,log}
gpus=$1; tag=$2; list=$3; port=29600
mkdir -p gpurun_out
while read -r name args; do
  [ -z "$name" ] && continue
  port=$((port+1))
  echo "== $name: $args"
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $gpus --master-addr 127.0.0.1 --master-port $port bench.py --gpus $gpus $args ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.log
  echo "rc=$?"; grep "^\[bench\]\|Error\|real" gpurun_out/${tag}_${name}.log | tail -5 | cut -c1-300
  tail -1 gpurun_out/${tag}_${name}.json | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); r=j['roofline']; m=r['hit_mix']
    print('  ->', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms/step e2e',round(j['e2e']['value']/1e6,2),'| gather',round(r['gather_ms_per_step'],4),'ms bound',r['bound'],'frac',round(r['frac'],3),'mix l/p/h',round(m['local'],3),round(m['peer'],3),round(m['host'],3),'| links GB/s',{k:round(v,1) for k,v in m['link_GBps_achieved'].items()},'| parity',(j.get('parity_selfcheck') or {}).get('all_ranks_ok'))
    print('     breakdown', {k:round(v,4) for k,v in j['breakdown_ms'].items()})
except Exception as e: print('  no json', e)
"
done < $list
