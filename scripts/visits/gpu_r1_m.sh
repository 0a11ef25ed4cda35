#!/bin/bash
# re-entry visit: parity tests, bench (default + in-flight sweep), ncu full capture of the sampler kernels
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
j=json.loads(line);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_m.log
timeout 600 python bench.py > gpurun_out/bench_m_default.json 2> gpurun_out/bench_m.err; show gpurun_out/bench_m_default.json "products default"
for nf in 1 3 4; do timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --inflight $nf > gpurun_out/bench_m_if$nf.json 2> gpurun_out/bench_m.err; show gpurun_out/bench_m_if$nf.json "products inflight$nf"; done
timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 1 > gpurun_out/bench_m_ov1.json 2> gpurun_out/bench_m.err; show gpurun_out/bench_m_ov1.json "products inflight1 overlap1"
timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/bench_m_ov0.json 2> gpurun_out/bench_m.err; show gpurun_out/bench_m_ov0.json "products inflight1 overlap0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sample_hop|rank_relabel" -s 16 -c 4 -o gpurun_out/prof_sampler_m -f \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/ncu_m.log 2>&1
ls -la gpurun_out
