#!/bin/bash
# 2-GPU box: parity tests (incl. partitioned cache over CUDA IPC), bench at N=1 and N=2
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu6.log
echo "== bench N=1"; timeout 600 python bench.py --no-cpu-baseline --no-server-e2e > gpurun_out/bench_f_n1.json 2> gpurun_out/bench_f_n1.err; python -c "
import json;j=json.load(open('gpurun_out/bench_f_n1.json'));print(round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()}, j['tier_rows'])"; tail -2 gpurun_out/bench_f_n1.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_f_n2.json 2> gpurun_out/bench_f_n2.err; python -c "
import json;j=json.load(open('gpurun_out/bench_f_n2.json'));print(round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()}, j['tier_rows'])"; tail -5 gpurun_out/bench_f_n2.err
