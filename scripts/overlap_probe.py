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
bg = dp.alloc_batch()            # prepared batch for the gather-only stream
dp.run_once(dp.params(d_train, d_lab, B, 0, seed=1, batch_id=0), bg); torch.cuda.synchronize()
rows = int(bg.node_counter[9 + H].item())
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
