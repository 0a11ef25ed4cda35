#!/bin/bash
set -u
N=${NGPU:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')]
if not line: print('$2', 'NO JSON'); sys.exit(0)
j=json.loads(line[-1]);r=j['roofline'];D=j['config']['feature_dim'];m=r['hit_mix'];rows=r['rows_per_step']
print('$2', round(j['value']/1e6,2),'M seeds/s','rows',int(rows),'gather ms',round(r['gather_ms_per_step'],4),'peer GB/s/GPU',round(rows*m['peer']*4*D/1e9/(r['gather_ms_per_step']*1e-3),1),'mixfrac',round(m['frac_of_mix_roofline'],3),'bound',m['bound'], 'N', j['config']['num_nodes'])"; }
run() { name=$1; shift; env ${ENVV:-X=1} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --no-cpu-baseline --steps 40 "$@" > gpurun_out/bench_w_n${N}_$name.json 2> gpurun_out/bench_w_n${N}_$name.err || tail -5 gpurun_out/bench_w_n${N}_$name.err; show gpurun_out/bench_w_n${N}_$name.json "N=$N $name"; }
run d100_vmm --kg $N
run d128_mid_vmm --kg $N --workload paper100m --scale 0.1
ENVV="LG_SHARD_IPC=legacy" run d128_mid_legacy --kg $N --workload paper100m --scale 0.1
run d100_x4_vmm --kg $N --scale 4
run uk_vmm --kg $N --workload ukunion --steps 30
