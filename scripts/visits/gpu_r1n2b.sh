#!/bin/bash
# r01d, 2 GPUs: bytes in flight of the NVLink-bound gather (UK-Union shape, Kg=2)
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2','NO JSON'); sys.exit(0)
j=json.loads(line[-1]);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'mix',j['roofline']['hit_mix']['bound'],round(j['roofline']['hit_mix']['frac_of_mix_roofline'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
run() { env $2 timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline --no-server-e2e $3 > gpurun_out/bench_n2b_$1.json 2> gpurun_out/bench_n2b_$1.err || tail -5 gpurun_out/bench_n2b_$1.err
  show gpurun_out/bench_n2b_$1.json "$1"; }
run uk_kg2_130 "LG_GATHER_SMEM_KB=130" "--workload ukunion --kg 2 --steps 100"
run uk_kg2_200 "LG_GATHER_SMEM_KB=200" "--workload ukunion --kg 2 --steps 100"
run uk_kg2_r16_200 "LG_GATHER_SMEM_KB=200 LG_TMA_ROWS=16" "--workload ukunion --kg 2 --steps 100"
run uk_kg2_ldg "LG_L2_HINTS=4" "--workload ukunion --kg 2 --steps 100 --gather ldg"
