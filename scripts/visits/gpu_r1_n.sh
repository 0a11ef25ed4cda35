#!/bin/bash
# re-entry visit 2: parity tests, smoke, both bench arms, ncu launch list + full capture of the current kernels
set -u
mkdir -p gpurun_out
show() { python -c "
import json,sys
txt=open('$1').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
j=json.loads(line);print('$2', round(j['value']/1e6,2),'M seeds/s', round(j['ms_per_step'],4),'ms e2e',round(j['e2e']['value']/1e6,2),'sync',round(j['e2e_sync_per_step']['value']/1e6,2),'frac',round(j['roofline']['frac'],3),{k:round(v,4) for k,v in j['breakdown_ms'].items()})"; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_n.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_n.log
timeout 600 python bench.py > gpurun_out/bench_n_default.json 2> gpurun_out/bench_n.err; show gpurun_out/bench_n_default.json "products default"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_n_reference.json 2>> gpurun_out/bench_n.err; tail -c 600 gpurun_out/bench_n_reference.json
timeout 600 python bench.py --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/bench_n_ov0.json 2>> gpurun_out/bench_n.err; show gpurun_out/bench_n_ov0.json "products inflight1 overlap0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_n.csv \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/ncu_launch_n.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_|sample_hop|rank_kernel|relabel_kernel|batch_generate|pm_clear" -s 36 -c 8 -o gpurun_out/prof_all_n -f \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-server-e2e --inflight 1 --overlap 0 > gpurun_out/ncu_full_n.log 2>&1
ncu -i gpurun_out/prof_all_n.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum > gpurun_out/prof_all_n_raw.csv 2>&1
ls -la gpurun_out
