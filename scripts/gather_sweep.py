#!/usr/bin/env python
"""Gather-kernel sweep on one B200: movers x tuning knobs x access patterns.  Each configuration runs in a
child process because the knobs are read once per process (LG_LDG_R, LG_LDG_HINT, LG_TMA_STAGES, ...).
Usage: python scripts/gather_sweep.py            (parent: runs the matrix, prints a table)
       python scripts/gather_sweep.py --child    (one configuration, JSON on stdout)"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    from legion_b200 import capi
    L = capi.load()
    dim = int(os.environ.get("SW_DIM", "100"))
    pattern = os.environ.get("SW_PATTERN", "skew")
    variant = int(os.environ.get("SW_VARIANT", "1"))
    cached = int(os.environ.get("SW_CACHED", "1"))
    N, n_rows = 2_449_029, 780_000
    dev = "cuda:0"
    feat = torch.empty((N, dim), dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(L.lg_synth_features(st, 0, N, dim, 1, feat.data_ptr()))
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    if pattern == "seq":
        ids = torch.arange(n_rows, dtype=torch.int32, device=dev)
    elif pattern == "rand":
        ids = torch.randperm(N, device=dev, generator=g)[:n_rows].to(torch.int32)
    else:  # skewed popularity like the synthetic graph: rank = N*u^3, unique, shuffled
        u = torch.rand(6_000_000, device=dev, generator=g, dtype=torch.float64)
        r = torch.unique((u * u * u * N).long().clamp_(max=N - 1))
        r = r[torch.randperm(r.numel(), device=dev, generator=g)][:n_rows]
        ids = r.to(torch.int32)  # rank == row index of a hotness-ordered shard
        n_rows = ids.numel()
    cache = capi.FeatureCache()
    cache.dim, cache.num_nodes, cache.backing = dim, N, feat.data_ptr()
    if cached:  # identity directory: row v of shard 0 (so `hot` rows = low ids)
        directory = torch.arange(N, dtype=torch.int32, device=dev)
        cache.n_parts, cache.shard_rows, cache.directory = 1, N, directory.data_ptr()
        cache.shard[0] = feat.data_ptr()
    out = torch.empty((n_rows, dim), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for rep in range(12):
        if os.environ.get("SW_FLUSH", "1") == "1":
            flush.fill_(rep)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.check(L.lg_gather_rows(st, C.byref(cache), ids.data_ptr(), n_rows, out.data_ptr(), 0, variant, None))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ok = bool(torch.equal(out.view(torch.int32), feat[ids.long()].view(torch.int32)))
    times = sorted(times[2:])
    ms = times[len(times) // 2]
    print(json.dumps({"ms": ms, "GBps": n_rows * (8 * dim + 8) / ms / 1e6, "rows": n_rows, "ok": ok}))


def main():
    if "--child" in sys.argv:
        return child()
    peak = 6547.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    matrix = []
    for dim in (100, 128):
        for pattern in ("seq", "skew"):
            matrix.append(dict(SW_DIM=dim, SW_PATTERN=pattern, SW_VARIANT=2))
    matrix.append(dict(SW_DIM=100, SW_PATTERN="skew", SW_VARIANT=1))
    matrix.append(dict(SW_DIM=100, SW_PATTERN="skew", SW_VARIANT=1, LG_LDG_R=4))
    for stages in (4, 6):
        matrix.append(dict(SW_DIM=100, SW_PATTERN="skew", SW_VARIANT=2, LG_TMA_STAGES=stages))
    matrix.append(dict(SW_DIM=100, SW_PATTERN="skew", SW_VARIANT=2, SW_CACHED=0))
    matrix.append(dict(SW_DIM=100, SW_PATTERN="skew", SW_VARIANT=2, SW_FLUSH=0))
    print(f"{'config':70s} {'ms':>8s} {'GB/s':>8s} {'frac':>6s} ok")
    for cfg in matrix:
        env = dict(os.environ, **{k: str(v) for k, v in cfg.items()})
        r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=300)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{json.dumps(cfg):70s} {j['ms']:8.4f} {j['GBps']:8.1f} {j['GBps'] / peak:6.3f} {j['ok']}", flush=True)
        except Exception:
            print(f"{json.dumps(cfg):70s} FAILED {r.stderr[-300:]}", flush=True)


if __name__ == "__main__":
    main()
