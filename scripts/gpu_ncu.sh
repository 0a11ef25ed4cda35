#!/bin/bash
# one ncu --set full capture of every step kernel (serialised), raw CSV of the key metrics -> gpurun_out/prof_${TAG}_raw.csv
set -u
TAG=${TAG:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none ${NCU_EXTRA:-} --import-source on -k regex:"gather_|sample_hop|rank_kernel|relabel_kernel|batch_generate|release_kernel|seed_local" -s ${SKIP:-36} -c ${COUNT:-9} -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 3 --warmup 3 --presample 2 --no-cpu-baseline --no-extras --no-parity-check --inflight 1 --overlap 0 ${BENCH_ARGS:-} > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__waves_per_multiprocessor,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,dram__sectors_read.sum > gpurun_out/prof_${TAG}_raw.csv 2>&1
python - <<PY
import csv
rows=[l for l in open('gpurun_out/prof_${TAG}_raw.csv') if not l.startswith('==')]
for r in list(csv.DictReader(rows))[1:]:
    print(r['Kernel Name'][:40].ljust(40), r['Grid Size'].ljust(14), 'us',r['gpu__time_duration.sum'][:6],'dramR',r['dram__bytes_read.sum'][:7],'W',r['dram__bytes_write.sum'][:7],'L2hit',r['lts__t_sector_hit_rate.pct'][:5],'dram%',r['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][:5])
PY
