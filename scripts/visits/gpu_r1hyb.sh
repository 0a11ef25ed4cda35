#!/bin/bash
# r01f, 2 GPUs: hybrid placement (hottest rows replicated, rest partitioned over NVLink), products shape
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for r in 0.1 0.3; do
  timeout 120 $TR bench.py --gpus 2 --no-cpu-baseline --no-server-e2e --kg 2 --replicate-ratio $r --steps 100 > gpurun_out/bench_hyb_$r.json 2> gpurun_out/bench_hyb_$r.err || tail -3 gpurun_out/bench_hyb_$r.err
  python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/bench_hyb_$r.json') if l.startswith('{')][-1])
m=j['roofline']['hit_mix']
print('replicate-ratio $r:', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms; mix local/peer', round(m['local'],3), round(m['peer'],3), m['bound'], round(m['frac_of_mix_roofline'],3), 'gather ms', round(j['roofline']['gather_ms_per_step'],4))
PY
done
