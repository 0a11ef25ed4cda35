#!/usr/bin/env python
"""Per-tile phase timeline of the sampler kernels on the products-shaped bench workload (diagnostics)."""
import argparse, os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from legion_b200 import capi
from legion_b200.runner import DataPath

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="products")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--out", default="gpurun_out/trace.npy")
args = ap.parse_args()
shape = bench.shape_of(args)
N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
ip, ix, feat, lab, E = bench.device_dataset(shape, 0)
train = bench.train_split(shape, 1)[0]
d_train = torch.from_numpy(train).cuda()
d_lab = lab[d_train.long()].contiguous()
dp = DataPath(0, fanout, B, N, D)
dp.set_full_graph(ip.data_ptr(), ix.data_ptr(), keep=[ip, ix])
dp.set_overlap(0)
buf = dp.alloc_batch(feature_rows=1)
L = dp.L
words = int(L.lg_debug_trace_words())
tr = torch.zeros(words, dtype=torch.int64, device="cuda")
for it in range(5):
    p = dp.params(d_train, d_lab, B, it, seed=bench.SEED, batch_id=it)
    dp.run_once(p, buf, gather=False)
torch.cuda.synchronize()
capi.check(L.lg_debug_set_trace(dp.sampler, C.c_void_p(tr.data_ptr())))
p = dp.params(d_train, d_lab, B, 7, seed=bench.SEED, batch_id=7)
dp.run_once(p, buf, gather=False)
torch.cuda.synchronize()
capi.check(L.lg_debug_set_trace(dp.sampler, None))
t = tr.cpu().numpy().reshape(-1, 2048, 8, 2)
np.save(args.out, t[:4])
names = {0: "sample h1", 1: "rank h1", 2: "sample h2", 3: "rank h2"}
for k in range(4):
    g = t[k, :, :, 0].astype(np.float64)  # globaltimer ns
    used = g[:, 0] > 0
    n = int(used.sum())
    if n == 0:
        continue
    g = g[used]
    t0 = g[:, 0].min()
    last = g.max()
    print(f"== {names[k]}: {n} tiles, span {(last - t0) / 1e3:.1f} us")
    nph = 5 if k % 2 == 0 else 8
    for ph in range(nph):
        col = g[:, ph] - t0
        print(f"   phase {ph}: at min {col.min() / 1e3:7.1f}  p50 {np.median(col) / 1e3:7.1f}  max {col.max() / 1e3:7.1f} us"
              + ("" if ph == 0 else f"   delta from prev p50 {np.median(g[:, ph] - g[:, ph - 1]) / 1e3:6.2f} max {(g[:, ph] - g[:, ph - 1]).max() / 1e3:6.2f}"))
    # per-tile lifetime by tile index deciles
    life = (g[:, nph - 1] - g[:, 0]) / 1e3
    idx = np.linspace(0, n - 1, 9).astype(int)
    print("   tile:start/life us  " + "  ".join(f"{i}:{(g[i, 0] - t0) / 1e3:.1f}/{life[i]:.1f}" for i in idx))
