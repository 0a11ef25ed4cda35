#!/usr/bin/env python
"""How much do the sampler chain and the gather slow each other down when they run concurrently?
Times N sampler-only batches on one stream, N gathers of a prepared batch on another, then both together."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from legion_b200 import capi
from legion_b200.runner import DataPath

class A: workload="products"; scale=1.0; batch=0
shape = bench.shape_of(A)
N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
H = len(fanout)
dev = "cuda:0"; torch.cuda.set_device(0)
ip, ix, feat, lab, E = bench.device_dataset(shape, 0)
train = bench.train_split(shape, 1)[0]
d_train = torch.from_numpy(train).to(dev); d_lab = lab[d_train.long()].contiguous()
steps = (len(train) - 1) // B
dp = DataPath(0, fanout, B, N, D)
dp.set_full_graph(ip.data_ptr(), ix.data_ptr(), keep=[ip, ix]); dp.set_backing_features(feat.data_ptr(), keep=[feat])
hot = torch.bincount(ix.long(), minlength=N)
order, _ = dp.rank_hotness(hot)
dp.build_feature_cache(order, N)
dp.set_overlap(0); dp.set_gather_fusion(2)
if os.environ.get("PROBE_GATHER") == "ldg":
    dp.set_gather_variant(capi.GATHER_LDG)
bg = dp.alloc_batch()            # prepared batch for the gather-only stream
dp.run_once(dp.params(d_train, d_lab, B, 0, seed=1, batch_id=0), bg); torch.cuda.synchronize()
rows = int(bg.node_counter[9 + H].item())
SMALL = float(os.environ.get("SAMPLER_SCALE", "1.0"))
if SMALL < 1.0:  # the sampler walks its own, smaller graph (L2-resident topology): separates DRAM from L2/SM contention
    class A2: workload = "products"; scale = SMALL; batch = 0
    sh2 = bench.shape_of(A2)
    ip2, ix2, _f2, lab2, E2 = bench.device_dataset(dict(sh2, dense=False), 0)
    tr2 = bench.train_split(sh2, 1)[0]
    d_train = torch.from_numpy(tr2).to(dev); d_lab = lab2[d_train.long()].contiguous()
    steps = (len(tr2) - 1) // B
    dp2 = DataPath(0, fanout, B, sh2["N"], D); dp2.set_full_graph(ip2.data_ptr(), ix2.data_ptr(), keep=[ip2, ix2]); dp2.set_overlap(0)
    print("sampler graph: N", sh2["N"], "E", E2, "topology MB", (E2 * 4 + sh2["N"] * 8) / 1e6)
else:
    dp2 = DataPath(0, fanout, B, N, D); dp2.share_storage_from(dp); dp2.set_overlap(0)
bs = dp2.alloc_batch(feature_rows=1)
S1, S2 = torch.cuda.Stream(), torch.cuda.Stream()
n = int(os.environ.get("N_ITERS", "100"))

def sampler(k):
    with torch.cuda.stream(S1):
        for i in range(k):
            dp2.run_once(dp2.params(d_train, d_lab, B, i % steps, seed=1, batch_id=i), bs, gather=False)
def gather(k):
    with torch.cuda.stream(S2):
        for i in range(k):
            capi.check(dp.L.lg_feature_cache_lookup_range(dp.sampler, C.c_void_p(S2.cuda_stream), C.byref(dp.cache), 3 * H + 1, 0, 0, C.byref(bg.c), None))

def timed(fs):
    torch.cuda.synchronize()
    ev = {}
    for name, st, f in fs:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st): a.record()
        ev[name] = (a, b, st)
    for name, st, f in fs: f(n)
    for name, (a, b, st) in ev.items():
        with torch.cuda.stream(st): b.record()
    torch.cuda.synchronize()
    return {name: a.elapsed_time(b) / n for name, (a, b, st) in ev.items()}

sampler(5); gather(5); torch.cuda.synchronize()
print("rows", rows, "alg GB", rows * (8 * D + 8) / 1e9)
print("alone   ", timed([("sampler", S1, sampler)]), timed([("gather", S2, gather)]))
print("together", timed([("sampler", S1, sampler), ("gather", S2, gather)]))
print("alone   ", timed([("sampler", S1, sampler)]), timed([("gather", S2, gather)]))

# ---- what kind of neighbour slows the gather down? ----
S3 = torch.cuda.Stream()
xa = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16); xb = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)
small = torch.randn(1 << 20, device=dev)            # 4 MB: L2-resident elementwise traffic
big = torch.randn(1 << 28, device=dev)              # 1 GB: DRAM streaming
idx = torch.randint(0, 1 << 28, (1 << 21,), device=dev)   # 2 M random 4-byte reads out of 1 GB
def matmul(k):
    with torch.cuda.stream(S3):
        for _ in range(k): torch.mm(xa, xb)
def l2_elementwise(k):
    with torch.cuda.stream(S3):
        for _ in range(k * 8): small.mul_(1.0001)
def dram_stream(k):
    with torch.cuda.stream(S3):
        for _ in range(k): big[: 1 << 25].mul_(1.0001)   # 128 MB read + 128 MB written per call
def random_reads(k):
    with torch.cuda.stream(S3):
        for _ in range(k * 2): big[idx]
for name, fn in (() if os.environ.get("PROBE") == "tiny" else (("bf16 matmul 4096^3", matmul), ("L2-resident elementwise", l2_elementwise), ("DRAM streaming 256 MB/iter", dram_stream),
                 ("2x2M random 4-byte reads/iter", random_reads))):
    fn(3); torch.cuda.synchronize()
    print(name.ljust(32), "alone", timed([("other", S3, fn)]), "with gather", timed([("other", S3, fn), ("gather", S2, gather)]))

# spinning neighbours: no memory traffic at all, varying SM occupancy; one long launch per gather iteration
for ctas, thr in ((1, 32), (148, 32), (148 * 4, 256), (148 * 8, 256)):
    def spin(k, ctas=ctas, thr=thr):
        for _ in range(k):
            capi.check(dp.L.lg_debug_spin(C.c_void_p(S3.cuda_stream), ctas, thr, 250000))   # ~0.13 ms at 1.9 GHz
    spin(3); torch.cuda.synchronize()
    print(f"spin {ctas}x{thr}".ljust(32), "alone", timed([("other", S3, spin)]), "with gather", timed([("other", S3, spin), ("gather", S2, gather)]))
# the same launches spread over independent streams: is it the launch, or the completion -> launch dependency?
many = [torch.cuda.Stream() for _ in range(9)]
def tiny_streams(k):
    for i in range(k * 27):
        capi.check(dp.L.lg_debug_spin(C.c_void_p(many[i % len(many)].cuda_stream), 1, 128, 2000))
def timed_many(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(S2): a.record()
    fn(n); gather(n)
    with torch.cuda.stream(S2): b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
tiny_streams(3); torch.cuda.synchronize()
print("27 tiny launches/iter over 9 independent streams: gather", timed_many(tiny_streams))
# many short launches of an empty-ish kernel: is the disturbance per SM (only where the CTAs land) or global?
for ctas, thr, per_iter in ((148, 128, 27), (1, 128, 27), (16, 128, 27), (148, 128, 7), (148 * 4, 256, 7)):
    def tiny(k, ctas=ctas, thr=thr, per_iter=per_iter):
        for _ in range(k * per_iter):
            capi.check(dp.L.lg_debug_spin(C.c_void_p(S3.cuda_stream), ctas, thr, 2000))
    tiny(3); torch.cuda.synchronize()
    print(f"{per_iter} tiny launches/iter {ctas}x{thr}".ljust(32), "alone", timed([("other", S3, tiny)]), "with gather", timed([("other", S3, tiny), ("gather", S2, gather)]))
