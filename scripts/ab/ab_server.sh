python -m pytest tests/test_server_gpu.py tests/test_boundary_ref_gpu.py -m gpu -x -q 2>&1 | tail -3
for st in 3 0 3 0; do for a in "" "--csc --server-csc"; do LEGION_STAGING=$st python scripts/server_e2e.py --epochs 10 $a 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('staging=$st $a', round(j['seeds_per_s']/1e6,2),'M seeds/s', round(j['ms_per_batch'],4),'ms/batch', 'csc_by_server',j['csc_built_by_server'])"; done; done
