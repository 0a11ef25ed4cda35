#!/usr/bin/env python
"""Context number: the REFERENCE's own kernels (oracle/_ref/libref_ops.so = /root/reference sources compiled in place
for sm_100a, launch configurations as in the reference: 16x1024 sampler, 32x1024 gather) on the bench workload.
Only the device kernels are timed (batch_generate, counter_update, random_sample, construct_graph,
multiGPU_feat_cache_lookup); the reference's bcht probe kernels, its >= 10 blocking cudaMemcpy per batch and the
O(N) bitmap memset are... the memset IS included (it is inside BatchGenerate), the probes and host syncs are not.
So this is an optimistic bound for the reference on B200."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def P(t):
    return C.c_void_p(t.data_ptr())


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_ops.so"))

    class A:
        workload, scale, batch = "products", 1.0, 0
    shape = bench.shape_of(A)
    N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
    ip, ix, feat, lab, E = bench.device_dataset(shape, 0)
    train = bench.train_split(shape, 1)[0]
    d_train = torch.from_numpy(train).cuda()
    d_lab = lab[d_train.long()].contiguous()
    num_ids = B + B * fanout[0] + B * fanout[0] * fanout[1]
    z = lambda n, dt=torch.int32: torch.zeros(n, dtype=dt, device="cuda")  # noqa: E731
    ids, labels = z(num_ids), z(B)
    a_src, a_dst, o_src, o_dst = z(num_ids), z(num_ids), z(num_ids), z(num_ids)
    accessed, posmap, nc, ec = z(N // 32 + 1), z(N), z(16), z(16)
    part_idx = torch.full((num_ids,), -2, dtype=torch.int8, device="cuda")
    part_off = z(num_ids)
    tab_ip = torch.tensor([ip.data_ptr()], dtype=torch.int64, device="cuda")
    tab_ix = torch.tensor([ix.data_ptr()], dtype=torch.int64, device="cuda")
    ptrs = torch.tensor([feat.data_ptr()], dtype=torch.int64, device="cuda")
    out = torch.empty((num_ids, D), dtype=torch.float32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    steps, warm = 20, 3
    tot, rows = 0.0, 0
    parts = {"batch_generate": 0.0, "sample": 0.0, "gather": 0.0}
    for s in range(warm + steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        k = 0
        ev[k].record()
        ref.ref_batch_generate(st, P(ids), P(labels), B, s % 30, P(d_train), P(d_lab), d_train.numel(), P(posmap), P(accessed),
                               N, P(nc), P(ec), 2)
        k += 1; ev[k].record()
        gather_ms = 0.0
        for op in (1, 4, 7):
            if op > 1:
                h = op // 3
                ref.ref_random_sample(st, P(ids), 3 * h, P(tab_ip), P(tab_ix), P(part_idx), P(part_off), fanout[h - 1], 0,
                                      P(a_src), P(a_dst), P(o_src), P(o_dst), P(accessed), P(posmap), P(nc), P(ec))
                k += 1; ev[k].record()
            ref.ref_counter_update(st, P(nc), P(ec), op)
            # FindFeat result for a fully cached table in identity order: cache_index = id (excluded from timing like the bcht probe)
            torch.cuda.synchronize()
            h_nc = nc.cpu().numpy()
            off, cnt = int(h_nc[2]), int(h_nc[3])
            cache_index = ids[off:off + cnt].contiguous()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            ref.ref_feat_cache_lookup(st, P(feat), P(ptrs), D, P(ids), P(cache_index), N, P(nc), P(out), N, op)
            g1.record()
            torch.cuda.synchronize()
            gather_ms += g0.elapsed_time(g1)
            k += 1; ev[k].record()
        torch.cuda.synchronize()
        if s >= warm:
            bg = ev[0].elapsed_time(ev[1])
            # sampling = events around the two ref_random_sample calls
            smp = ev[2].elapsed_time(ev[3]) + ev[4].elapsed_time(ev[5])
            parts["batch_generate"] += bg; parts["sample"] += smp; parts["gather"] += gather_ms
            tot += bg + smp + gather_ms
            rows += int(nc.cpu().numpy()[11])
    ms = tot / steps
    print(json.dumps({"what": "reference kernels (sm_100a recompiled, reference launch configs), device time only",
                      "ms_per_batch": ms, "seeds_per_s": B / (ms * 1e-3), "rows_per_batch": rows / steps,
                      "gather_GBps": rows / steps * (8 * D + 8) / (parts["gather"] / steps * 1e-3) / 1e9,
                      "parts_ms": {k: v / steps for k, v in parts.items()}}))


if __name__ == "__main__":
    main()
