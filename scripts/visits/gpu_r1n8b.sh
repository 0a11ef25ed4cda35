#!/bin/bash
# end of round 1, 8 GPUs: the default bench line (products shape, replicated cache, 3 runners in flight per GPU)
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --no-cpu-baseline --no-server-e2e > gpurun_out/r01f_bench_n8_products.json 2> gpurun_out/r01f_bench_n8.err || tail -5 gpurun_out/r01f_bench_n8.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r01f_bench_n8_products.json') if l.startswith('{')][-1])
print(round(j['value']/1e6,2),'M seeds/s on', j['n_gpus'],'GPUs', round(j['ms_per_step'],4),'ms; e2e', round(j['e2e']['value']/1e6,2), 'frac', round(j['roofline']['frac'],3), j['config']['cache'][:40])
PY
