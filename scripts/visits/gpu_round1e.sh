#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu5.log
for cfg in "1 0 0" "1 2 0" "1 2 2" "2 0 0" "2 0 2" "2 2 2" "3 0 2" "3 2 2" "4 0 2"; do
set -- $cfg
echo "== bench inflight=$1 overlap=$2 fuse=$3"; timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --steps 60 --inflight $1 --overlap $2 --fuse $3 > gpurun_out/bench_e_$1_$2_$3.json 2> gpurun_out/bench_e.err; python -c "
import json;j=json.load(open('gpurun_out/bench_e_$1_$2_$3.json'));print(round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; tail -1 gpurun_out/bench_e.err
done
