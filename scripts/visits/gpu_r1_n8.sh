#!/bin/bash
# 8 GPUs of one box: multi-GPU parity test, default bench (what the driver's scaling run launches), NVSwitch-partitioned
# cache (Kg=8) on the products and UK-Union shapes
set -u
N=${NGPU:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];D=j['config']['feature_dim'];m=r['hit_mix'];rows=r['rows_per_step']
print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'rows',int(rows),'gather ms',round(r['gather_ms_per_step'],4),'peer GB/s/GPU',round(rows*m['peer']*4*D/1e9/(r['gather_ms_per_step']*1e-3),1),'mix',round(m['local'],3),round(m['peer'],3),round(m['host'],3),'mixfrac',round(m['frac_of_mix_roofline'],3),'bound',m['bound'])"; }
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err || tail -5 gpurun_out/bench_n${N}_$name.err; show gpurun_out/bench_n${N}_$name.json "N=$N $name"; }
run default
run kgN --kg $N --no-cpu-baseline --steps 100
run uk_kgN --kg $N --workload ukunion --steps 50 --no-cpu-baseline
run uk_kgN_cr50 --kg $N --workload ukunion --steps 20 --cache-ratio 0.5 --no-cpu-baseline
