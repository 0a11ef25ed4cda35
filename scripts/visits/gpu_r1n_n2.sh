#!/bin/bash
# r01d, 2 GPUs: the new gather tiles on the NVLink-partitioned layouts (Kg=2), multi-GPU parity tests
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2','NO JSON'); sys.exit(0)
j=json.loads(line[-1]);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),'mix',j['roofline']['hit_mix']['bound'],round(j['roofline']['hit_mix']['frac_of_mix_roofline'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
OLD="LG_TMA_ROWS=32 LG_GATHER_SMEM_KB=220 LG_TMA_CTAS=4"
echo "== pytest gpu (multi-GPU + server)"; timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_server_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { # name, env, args
  env $2 timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline --no-server-e2e $3 > gpurun_out/bench_n2_$1.json 2> gpurun_out/bench_n2_$1.err || tail -5 gpurun_out/bench_n2_$1.err
  show gpurun_out/bench_n2_$1.json "$1"; }
run products_kg1 "LG_L2_HINTS=4" ""
run products_kg2 "LG_L2_HINTS=4" "--kg 2"
run products_kg2_oldtiles "$OLD" "--kg 2"
run products_kg2_r16 "LG_TMA_ROWS=16" "--kg 2"
run ukunion_kg2 "LG_L2_HINTS=4" "--workload ukunion --kg 2 --steps 100"
run ukunion_kg2_oldtiles "$OLD" "--workload ukunion --kg 2 --steps 100"
