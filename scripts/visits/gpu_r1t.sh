#!/bin/bash
# r01t: lazy relabel in the server path; full GPU suite; server e2e
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== server e2e (lazy relabel + fused gathers)"; timeout 600 python scripts/server_e2e.py 2>/dev/null | tail -1 | cut -c1-600
echo "== server e2e (reference schedule)"; LEGION_LAZY_RELABEL=0 LEGION_FUSE_GATHERS=0 timeout 600 python scripts/server_e2e.py 2>/dev/null | tail -1 | cut -c1-600
