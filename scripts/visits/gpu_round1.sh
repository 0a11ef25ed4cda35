#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both gather movers), ncu launch list + full capture.
# Everything is wrapped in `timeout` so a hung kernel cannot hold the box.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench ldg"; timeout 600 python bench.py --gather ldg > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err; tail -c 3000 gpurun_out/bench_ldg.json; tail -3 gpurun_out/bench_ldg.err
echo "== bench tma"; timeout 600 python bench.py --gather tma --no-cpu-baseline > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; tail -c 1500 gpurun_out/bench_tma.json; tail -3 gpurun_out/bench_tma.err
echo "== reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 800 gpurun_out/bench_ref.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ldg.csv \
  python bench.py --gather ldg --steps 3 --warmup 3 --presample 2 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "== ncu full gather (ldg, tma) + sampler"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_ldg -s 12 -c 3 -o gpurun_out/prof_gather_ldg -f \
  python bench.py --gather ldg --steps 3 --warmup 3 --presample 2 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_tma -s 12 -c 3 -o gpurun_out/prof_gather_tma -f \
  python bench.py --gather tma --steps 3 --warmup 3 --presample 2 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sample_hop|rank_relabel" -s 16 -c 4 -o gpurun_out/prof_sampler -f \
  python bench.py --gather ldg --steps 3 --warmup 3 --presample 2 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out
