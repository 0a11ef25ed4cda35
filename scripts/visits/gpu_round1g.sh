#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
j=json.loads(line);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),'mix',j['roofline']['hit_mix']['bound'],round(j['roofline']['hit_mix']['frac_of_mix_roofline'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()}, j['clocks'])"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu7.log
for w in 1 8; do
LG_LOOKBACK_WIDE=$w timeout 600 python bench.py --no-cpu-baseline --no-server-e2e > gpurun_out/bench_g_n1_w$w.json 2> gpurun_out/bench_g.err; show gpurun_out/bench_g_n1_w$w.json "N1 wide=$w"; tail -2 gpurun_out/bench_g.err
done
timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_g_n2_kg1.json 2> gpurun_out/bench_g.err; show gpurun_out/bench_g_n2_kg1.json "N2 kg=auto"; grep -v "^\*\|OMP" gpurun_out/bench_g.err | tail -2
timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline --kg 2 > gpurun_out/bench_g_n2_kg2.json 2> gpurun_out/bench_g.err; show gpurun_out/bench_g_n2_kg2.json "N2 kg=2 tma3"
LG_TMA_STAGES=4 timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline --kg 2 > gpurun_out/bench_g_n2_kg2_s4.json 2> gpurun_out/bench_g.err; show gpurun_out/bench_g_n2_kg2_s4.json "N2 kg=2 tma4"
LG_TMA_STAGES=6 timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline --kg 2 > gpurun_out/bench_g_n2_kg2_s6.json 2> gpurun_out/bench_g.err; show gpurun_out/bench_g_n2_kg2_s6.json "N2 kg=2 tma6"
timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline --kg 2 --gather ldg > gpurun_out/bench_g_n2_kg2_ldg.json 2> gpurun_out/bench_g.err; show gpurun_out/bench_g_n2_kg2_ldg.json "N2 kg=2 ldg"
timeout 900 $TR bench.py --gpus 2 --impl reference --steps 5 --warmup 1 > gpurun_out/bench_g_n2_ref.json 2> gpurun_out/bench_g.err; head -c 300 gpurun_out/bench_g_n2_ref.json
