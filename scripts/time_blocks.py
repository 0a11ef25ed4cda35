#!/usr/bin/env python
"""Time lg_block_csc on the blocks of full-size products batches (B=8000, [25,10]) next to torch's own COO->CSC
route (argsort + bincount), which is what a framework does when handed the COO."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from legion_b200.blocks import BlockBuilder
from legion_b200.runner import DataPath

class A: workload = "products"; scale = 1.0; batch = 0
shape = bench.shape_of(A)
N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
H = len(fanout)
ip, ix, feat, lab, E = bench.device_dataset(shape, 0)
train = bench.train_split(shape, 1)[0]
d_train = torch.from_numpy(train).cuda(); d_lab = lab[d_train.long()].contiguous()
dp = DataPath(0, fanout, B, N, D)
dp.set_full_graph(ip.data_ptr(), ix.data_ptr(), keep=[ip, ix]); dp.set_backing_features(feat.data_ptr(), keep=[feat])
buf = dp.alloc_batch(feature_rows=1)
dp.run_once(dp.params(d_train, d_lab, B, 0, seed=1, batch_id=0), buf, gather=False); torch.cuda.synchronize()
nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
bb = BlockBuilder(int(ec[9 + H]))
out = {}
for h in range(H, 0, -1):
    e, num_dst = int(ec[9 + h]), int(nc[9 + h - 1])
    src, dst = buf.agg_src[:e], buf.agg_dst[:e]
    def ours(): return bb.csc(src, dst, num_dst)
    def torch_route():
        order = torch.argsort(dst, stable=True)
        indptr = torch.zeros(num_dst + 1, dtype=torch.int64, device="cuda")
        indptr[1:] = torch.cumsum(torch.bincount(dst, minlength=num_dst), 0)
        return indptr, src[order], order
    res = {}
    for name, fn in (("lg_block_csc", ours), ("torch_argsort", torch_route)):
        for _ in range(3): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        res[name + "_us"] = round(a.elapsed_time(b) * 1000 / 20, 1)
    o, t = ours(), torch_route()
    res["identical"] = bool(torch.equal(o[0].long(), t[0]) and torch.equal(o[1], t[1]) and torch.equal(o[2].long(), t[2]))
    out[f"block{h}"] = dict(edges=e, num_dst=num_dst, **res)
# all blocks of the batch in one set of launches, sizes read on the device (lg_block_csc_batch: what the server runs)
max_edges, max_dst, per, nodes, edges = [], [], B, B, 0
for f in fanout:
    max_dst.append(nodes); per *= f; edges += per; nodes += per; max_edges.append(edges)
for _ in range(3): bb.csc_batch(buf.c, H, max_edges, max_dst)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(20): bb.csc_batch(buf.c, H, max_edges, max_dst)
b.record(); torch.cuda.synchronize()
out["all_blocks_one_call_us"] = round(a.elapsed_time(b) * 1000 / 20, 1)
print(json.dumps(out))
