#!/usr/bin/env python
"""Random-row gather bandwidth from a PEER GPU's memory as a function of how the peer buffer is mapped and how big
it is.  One process, GPU 0 reads GPU 1: (a) cudaMalloc + cudaDeviceEnablePeerAccess, (b) cuMemCreate/cuMemMap (VMM)."""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from legion_b200 import capi
L = capi.load()
D, ROWS = int(os.environ.get("D", "128")), 1_000_000
capi.check(L.lg_enable_peer_access(2))
res = []
for gb in (0.5, 2.0, 8.0, 24.0):
    n = int(gb * 1e9 / (D * 4))
    for kind in ("cudaMalloc+peer", "vmm", "local"):
        torch.cuda.set_device(0 if kind == "local" else 1)
        p = C.c_void_p()
        if kind == "vmm":
            fd = C.c_int32(-1)
            capi.check(L.lg_vmm_alloc(n * D * 4, C.byref(p), C.byref(fd)))
            own = p.value
            torch.cuda.set_device(0)
            q = C.c_void_p()
            capi.check(L.lg_vmm_import(fd.value, n * D * 4, C.byref(q)))
            os.close(fd.value)
            src = q.value
        else:
            capi.check(L.lg_device_alloc(C.byref(p), n * D * 4))
            own = src = p.value
        torch.cuda.set_device(0)
        ids = torch.randint(0, n, (ROWS,), dtype=torch.int32, device="cuda:0")
        dst = torch.empty((ROWS, D), dtype=torch.float32, device="cuda:0")
        fc = capi.FeatureCache()
        fc.n_parts, fc.shard_rows, fc.dim, fc.num_nodes = 0, 0, D, n
        fc.backing, fc.directory = src, None
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for variant in (capi.GATHER_TMA, capi.GATHER_LDG):
            for _ in range(3):
                capi.check(L.lg_gather_rows(st, C.byref(fc), C.c_void_p(ids.data_ptr()), ROWS, C.c_void_p(dst.data_ptr()), 0, variant, None))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                capi.check(L.lg_gather_rows(st, C.byref(fc), C.c_void_p(ids.data_ptr()), ROWS, C.c_void_p(dst.data_ptr()), 0, variant, None))
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            res.append(dict(GB=gb, mapping=kind, mover="tma" if variant == capi.GATHER_TMA else "ldg", ms=round(ms, 4),
                            read_GBps=round(ROWS * D * 4 / 1e9 / (ms * 1e-3), 1)))
            print(res[-1], flush=True)
        if kind == "vmm":
            capi.check(L.lg_vmm_free(C.c_void_p(src)))
            torch.cuda.set_device(1)
            capi.check(L.lg_vmm_free(C.c_void_p(own)))
        else:
            torch.cuda.set_device(0 if kind == "local" else 1)
            capi.check(L.lg_device_free(C.c_void_p(own)))
        del ids, dst
print(json.dumps(res))
