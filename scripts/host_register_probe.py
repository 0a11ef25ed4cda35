#!/usr/bin/env python
"""Does lg_host_register take a /dev/shm mapping of this size on this box?  usage: host_register_probe.py GB [chunk_mb]"""
import ctypes as C, mmap, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legion_b200 import capi
gb = float(sys.argv[1])
if len(sys.argv) > 2: os.environ["LG_HOST_REGISTER_CHUNK_MB"] = sys.argv[2]
L = capi.load(); torch.cuda.init()
n = int(gb * (1 << 30)); path = f"/dev/shm/lg_probe_{os.getpid()}"
fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o600); t = time.time(); os.posix_fallocate(fd, 0, n); t_alloc = time.time() - t
mm = mmap.mmap(fd, n); os.close(fd); os.unlink(path)
buf = (C.c_char * n).from_buffer(mm); dp = C.c_void_p(); t = time.time()
rc = L.lg_host_register(C.c_void_p(C.addressof(buf)), n, C.byref(dp))
print(f"{gb} GB chunk={os.environ.get('LG_HOST_REGISTER_CHUNK_MB', '1024')} MB: rc={rc} {L.lg_last_error().decode() if rc else 'ok'} fallocate {t_alloc:.1f}s register {time.time() - t:.1f}s")
if rc == 0:
    x = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    capi.check(L.lg_memcpy_d2h(C.c_void_p(C.addressof(buf) + n - (1 << 20)), C.c_void_p(x.data_ptr()), 1 << 20, None)); torch.cuda.synchronize()
    capi.check(L.lg_host_unregister(C.c_void_p(C.addressof(buf))))
