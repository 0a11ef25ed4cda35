#!/bin/bash
# N GPUs on one box: multi-GPU parity test, then bench with the cache replicated (Kg=1) and NVSwitch-partitioned (Kg=N)
set -u
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(r['frac'],3),'mix',{k:(round(v,3) if isinstance(v,float) else v) for k,v in r['hit_mix'].items() if k in ('local','peer','host','bound','frac_of_mix_roofline')},{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --no-cpu-baseline "$@" > gpurun_out/bench_s_n${N}_$name.json 2> gpurun_out/bench_s_n${N}_$name.err || tail -5 gpurun_out/bench_s_n${N}_$name.err; show gpurun_out/bench_s_n${N}_$name.json "N=$N $name"; }
run kg1
run kgN --kg $N
run kgN_ldg --kg $N --gather ldg
run kgN_cr50 --kg $N --cache-ratio 0.5
