#!/bin/bash
# r01u: server reads the sticky status asynchronously (no sync on a stream that already holds the next batch)
set -u
mkdir -p gpurun_out
echo "== pytest gpu (server)"; timeout 600 python -m pytest tests/test_server_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== server e2e"; timeout 300 python scripts/server_e2e.py 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['seeds_per_s']/1e6,2),'M seeds/s', round(j['ms_per_batch'],4),'ms/batch')"
echo "== server e2e again"; timeout 300 python scripts/server_e2e.py 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(round(j['seeds_per_s']/1e6,2),'M seeds/s', round(j['ms_per_batch'],4),'ms/batch')"

CONFIGS="LG_L2_HINTS=4" BENCH_ARGS='--inflight 1' bash scripts/gpu_ab.sh
