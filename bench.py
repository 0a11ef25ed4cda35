#!/usr/bin/env python
"""bench.py — sampled+gathered seeds/s of the Legion mini-batch data path on B200.

A step = one mini-batch of `batch` seeds through the whole hot path on one GPU: batch_generate ->
feature gather of the seeds -> per hop (neighbour sampling -> dedup/reindex -> feature gather of the
new vertices), i.e. the ops of GPURunner::RunOnce (reference engine/server.cu:302-332).

  python bench.py [--gpus N --steps K --warmup W]          our arm (N>1: launched by torchrun)
  python bench.py --impl reference ...                      CPU arm: DGL-semantics sampler + index_select
                                                            (oracle port) on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the byte accounting.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x1E910
WORKLOADS = {
    # name: (shape key, fanout, batch, dmax, materialise the [N x D] matrix in vertex order?)
    "products": ("products", [25, 10], 8000, 20000, True),     # BASELINE.json configs[1] (default)
    "paper100m": ("paper100m", [25, 10], 8000, 20000, False),  # configs[2] shape
    "ukunion": ("ukunion", [25, 10], 8000, 20000, False),      # configs[3] shape: the metric's named graph
    "clueweb": ("clueweb", [15, 10, 5], 8000, 20000, False),   # configs[4] shape: GCN 3-hop, topology in host UVA
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: clocks DURING the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def recorded_traffic():
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# dataset
# ----------------------------------------------------------------------------------------------
def shape_of(args):
    from legion_b200 import synth
    key, fanout, batch, dmax, dense = WORKLOADS[args.workload]
    n, e, d, classes = synth.SHAPES[key]
    n = max(1000, int(n * args.scale))
    e_target = int(e * args.scale)
    return dict(name=key, N=n, E_target=e_target, D=d, classes=classes, fanout=fanout, batch=args.batch or batch,
                dmax=dmax, dmin=synth.dmin_for(n, e_target), dense=dense)


def device_dataset(shape, device):
    """graph + features + labels generated directly in HBM (include/legion_b200_synth.h)"""
    import torch
    from legion_b200 import capi
    L = capi.load()
    dev = f"cuda:{device}"
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    N, D = shape["N"], shape["D"]
    ip = torch.empty(N + 1, dtype=torch.int64, device=dev)
    capi.check(L.lg_synth_indptr(st, N, shape["dmin"], shape["dmax"], SEED, ip.data_ptr()))
    E = int(ip[N].item())
    ix = torch.empty(E, dtype=torch.int32, device=dev)
    capi.check(L.lg_synth_indices(st, N, ip.data_ptr(), SEED, ix.data_ptr()))
    feat = None
    if shape["dense"]:
        feat = torch.empty((N, D), dtype=torch.float32, device=dev)
        capi.check(L.lg_synth_features(st, 0, N, D, SEED, feat.data_ptr()))
    lab = torch.empty(N, dtype=torch.int32, device=dev)
    capi.check(L.lg_synth_labels(st, N, shape["classes"], lab.data_ptr()))
    torch.cuda.synchronize()
    return ip, ix, feat, lab, E


def train_split(shape, world):
    """random 10 % of the vertices (dataset/gen_sets.py:62-67), split by id % gpus (storage_management.cu:175-179)"""
    from legion_b200 import synth
    if shape["N"] <= 20_000_000:
        tr, _, _ = synth.split_sets(shape["N"], SEED)
    else:  # paper-scale: draw the permutation on the device
        import torch
        g = torch.Generator(device="cuda")
        g.manual_seed(SEED)
        tr = torch.randperm(shape["N"], device="cuda", generator=g)[: shape["N"] // 10].to(torch.int32).cpu().numpy()
    return synth.partition_ids(tr, world)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from legion_b200 import capi
    from legion_b200.runner import DataPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))
    capi.load()  # fails loudly if the CUDA library is missing

    shape = shape_of(args)
    N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
    H = len(fanout)
    t0 = time.time()
    ip, ix, feat, lab, E = device_dataset(shape, local)
    parts = train_split(shape, world)
    my_train = parts[rank]
    d_train = torch.from_numpy(my_train).to(dev)
    d_lab = lab[d_train.long()].contiguous()
    train_steps = (min(len(p) for p in parts) - 1) // B  # engine/ipc_service.cu:73-82
    assert train_steps >= 1, "training set smaller than one batch"
    if rank == 0:
        log(f"[bench] {shape['name']} N={N} E={E} D={D} world={world} train/gpu={len(my_train)} "
            f"train_steps={train_steps} setup {time.time() - t0:.1f}s")

    dp = DataPath(local, fanout, B, N, D, rank=rank, world=world)
    topo_host = args.topo == "host"
    host_keep = []
    if topo_host:
        # full CSR in cudaHostAllocMapped memory, read by the sampler through UVA (storage/storage_management.cu:100-115,
        # engine/operator_impl.cu:224-243); the device copy only feeds presampling/placement and is dropped afterwards
        from legion_b200.runner import MappedHostBuffer
        h_ip, h_ix = MappedHostBuffer((N + 1) * 8), MappedHostBuffer(max(E, 1) * 4)
        capi.check(dp.L.lg_memcpy_d2h(C.c_void_p(h_ip.host_ptr), C.c_void_p(ip.data_ptr()), (N + 1) * 8, dp._stream()))
        capi.check(dp.L.lg_memcpy_d2h(C.c_void_p(h_ix.host_ptr), C.c_void_p(ix.data_ptr()), E * 4, dp._stream()))
        torch.cuda.synchronize()
        host_keep += [h_ip, h_ix]
    dp.set_full_graph(ip.data_ptr(), ix.data_ptr(), keep=[ip, ix])  # topology replicated in each GPU's HBM
    feat_host = args.cache_ratio < 1.0
    if feat_host:
        # backing matrix in pinned host memory (cache/cache_impl.cuh:262-266 reads misses through UVA)
        from legion_b200.runner import MappedHostBuffer
        h_feat = MappedHostBuffer(N * D * 4)
        for r0 in range(0, N, 1 << 22):  # generated by the device straight into the mapped allocation
            capi.check(dp.L.lg_synth_features(dp._stream(), r0, min(1 << 22, N - r0), D, SEED,
                                              C.c_void_p(h_feat.dev_ptr + r0 * D * 4)))
        torch.cuda.synchronize()
        host_keep.append(h_feat)
        dp.set_backing_features(h_feat.dev_ptr, keep=[h_feat])
    elif feat is not None:
        dp.set_backing_features(feat.data_ptr(), keep=[feat])
    dp.set_overlap(args.overlap)
    dp.set_gather_variant({"auto": capi.GATHER_AUTO, "ldg": capi.GATHER_LDG, "tma": capi.GATHER_TMA}[args.gather])

    # --- presampling: hotness -> ranking -> interleaved placement (PreSc + CandidateSelection + FillUp) ---
    scratch = dp.alloc_batch(feature_rows=1)
    eh = torch.zeros(N, dtype=torch.int64, device=dev)
    nh = torch.zeros(N, dtype=torch.int64, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    pre = min(args.presample, train_steps)
    for it in range(pre):
        dp.run_presc(dp.params(d_train, d_lab, B, it, seed=SEED, batch_id=it), scratch, eh, nh, mx)
    torch.cuda.synchronize()
    if world > 1:
        dist.all_reduce(nh)  # init-time sum of hotness over the GPUs (cache/cache.cu:408-411); not on the serving path
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    order, _ = dp.rank_hotness(nh)
    # cache aggregation (Legion's cache_agg_mode: Kg GPUs share one partitioned cache, Kc = world/Kg replicas).
    # auto = the smallest power of two whose per-GPU shard fits the cache budget: on 180 GB parts the named
    # shapes replicate (Kg=1, no NVLink traffic); --kg N forces the NVSwitch-partitioned layout.
    table_bytes = N * D * 4
    if args.kg > 0:
        kg = args.kg
    else:
        kg = 1
        while kg < world and table_bytes / kg > args.cache_gb * 1e9:
            kg *= 2
    assert world % kg == 0, "world size must be a multiple of Kg"
    # rows cached across the clique: the whole table, or its hottest --cache-ratio fraction (misses -> pinned host)
    # --replicate-ratio r: the hottest r*N rows are stored on EVERY GPU of the clique (local reads for the head of the
    # distribution), only the rest of the cached rows is partitioned (hybrid placement, lg_place_features_hybrid)
    cached_rows = int(N * min(args.cache_ratio, 1.0))
    rep = min(int(N * max(args.replicate_ratio, 0.0)), cached_rows) if kg > 1 else 0
    cap = max(rep + (cached_rows - rep + kg - 1) // kg, 1)
    if feat is not None and not feat_host:
        dp.build_feature_cache(order, cap, kg=kg, j=rank % kg, dist=dist if world > 1 else None, replicate=rep)
    else:  # shards generated in place (paper-scale shapes: no [N x D] matrix in vertex order in HBM)
        dp.build_feature_cache_synth(order, cap, SEED, kg=kg, j=rank % kg, dist=dist if world > 1 else None,
                                     keep_backing=feat_host, replicate=rep)
    topo_cap = 0
    if topo_host:
        # hot-vertex topology cache in HBM (GraphCache, storage/graph_storage.cu:76-111), ranked by the presampled
        # edge hotness (QT, cache/cache.cu:420-440); everything else is read from the host CSR over PCIe
        if world > 1:
            dist.all_reduce(eh)
        if args.topo_cache_ratio > 0:
            order_t, _ = dp.rank_hotness(eh)
            topo_cap = max(1, (int(N * min(args.topo_cache_ratio, 1.0)) + kg - 1) // kg)
            dp.build_topology_cache(order_t, topo_cap, kg=kg, j=rank % kg, dist=dist if world > 1 else None)
            del order_t
        dp.repoint_full_graph(h_ip.dev_ptr, h_ix.dev_ptr, drop=[ip, ix])
        np_ip, np_ix = h_ip.numpy(np.int64, (N + 1,)), h_ix.numpy(np.int32, (E,))
        del ip, ix
    max_ids = int(mx.item())
    feature_rows = min(dp.num_ids, int(max_ids * 1.2) + 1)  # engine/server.cu:277
    del scratch, eh, nh
    torch.cuda.empty_cache()
    bufs = [dp.alloc_batch(feature_rows=feature_rows) for _ in range(2)]  # INTERBATCH_CON pipeline slots
    dp.set_gather_fusion(args.fuse)
    # additional batches in flight on the same GPU: own sampler scratch + buffers + stream, shared storage
    runners = [(dp, bufs, torch.cuda.current_stream())]
    for _ in range(args.inflight - 1):
        d2 = DataPath(local, fanout, B, N, D, rank=rank, world=world)
        d2.share_storage_from(dp)
        d2.set_overlap(args.overlap)
        d2.set_gather_fusion(args.fuse)
        d2.set_gather_variant({"auto": capi.GATHER_AUTO, "ldg": capi.GATHER_LDG, "tma": capi.GATHER_TMA}[args.gather])
        runners.append((d2, [d2.alloc_batch(feature_rows=feature_rows) for _ in range(2)], torch.cuda.Stream()))

    def run_steps(first, count, tier=False):
        """`count` batches round-robin over the in-flight runners; returns when all are enqueued and joined
        into the current stream"""
        cur = torch.cuda.current_stream()
        for _, _, st in runners[1:]:
            st.wait_stream(cur)
        for s in range(count):
            r, rb, st = runners[s % len(runners)]
            with torch.cuda.stream(st):
                r.run_once(params(first + s), rb[(s // len(runners)) % 2], tier=tier)
        for r, rb, st in runners:
            with torch.cuda.stream(st):
                for b in rb:
                    r.batch_wait(b)
            if st is not cur:
                cur.wait_stream(st)

    def params(step):
        return dp.params(d_train, d_lab, B, step % train_steps, seed=SEED, batch_id=step)

    # --- warm-up ---
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    torch.cuda.synchronize()
    run_steps(0, args.warmup)
    torch.cuda.synchronize()
    # self-check: the rows gathered for the last warm-up batch equal the feature function of their ids, bit for bit
    r0, rb0, _ = runners[(args.warmup - 1) % len(runners)]
    b0 = rb0[((args.warmup - 1) // len(runners)) % 2]
    n0 = int(b0.node_counter[9 + H].item())
    chk = torch.empty((n0, D), dtype=torch.float32, device=dev)
    capi.check(dp.L.lg_synth_feature_rows(dp._stream(), b0.ids.data_ptr(), n0, D, SEED, chk.data_ptr()))
    torch.cuda.synchronize()
    features_ok = bool(torch.equal(chk.view(torch.int32), b0.features[:n0].view(torch.int32)))
    assert features_ok, "gathered features differ from the feature function"
    del chk
    assert all(r.status() == 0 for r, _, _ in runners), "sampler overflow status"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- timed region: exactly K steps, device-timed, max over ranks ---
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_enq = time.perf_counter()
    run_steps(args.warmup, args.steps, tier=True)  # all in-flight batches are complete before the clock stops
    t_enq = time.perf_counter() - t_enq  # host time to enqueue the steps (no synchronisation inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    tiers = sum(r.tier_rows for r, _, _ in runners)
    if world > 1:
        dist.all_reduce(tiers)
    tiers = tiers.cpu().numpy().astype(np.int64)
    assert dp.status() == 0

    dp.set_overlap(min(args.overlap, 1))  # per-op timing and the synchronous e2e call: no cross-batch pipelining
    # --- instrumented pass: same steps, CUDA events around each op (per-kernel durations) ---
    L = dp.L
    st = dp._stream()
    rows_total = 0
    n_inst = min(args.steps, 20)
    acc = {}
    for s in range(n_inst):
        p = params(args.warmup + s)
        b = bufs[s % 2]
        ops = [("batch_generate", lambda: L.lg_batch_generate(dp.sampler, st, p.all_ids, p.all_labels, p.total_cap,
                                                              p.batch_size, p.counter, C.byref(b.c)))]
        pending = 0
        for hop in range(0, H + 1):
            if hop > 0:
                ops.append((f"sample{hop}", lambda hop=hop: L.lg_random_sample(
                    dp.sampler, st, C.byref(dp.topo), hop, p.rng_kind, p.rng_seed, p.batch_id, p.stream_id, C.byref(b.c), None)))
            last = hop == H
            if not last and (args.fuse == 2 or (args.fuse == 1 and hop == 0)):
                continue
            nm = f"gather{hop}" if pending == hop else f"gather{pending}-{hop}"  # rows of hops [pending, hop]
            ops.append((nm, lambda hop=hop, pending=pending: L.lg_feature_cache_lookup_range(
                dp.sampler, st, C.byref(dp.cache), 3 * hop + 1, pending, dp.local_part, C.byref(b.c), None)))
            pending = hop + 1
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
        evs[0].record()
        for i, (nm, fn) in enumerate(ops):
            capi.check(fn())
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, (nm, _) in enumerate(ops):
            acc[nm] = acc.get(nm, 0.0) + evs[i].elapsed_time(evs[i + 1])
        rows_total += int(b.node_counter[9 + H].item())
    clk = clocks.stop() if rank == 0 else None  # sampled from warm-up through the timed and instrumented passes
    breakdown = {k: v / n_inst for k, v in acc.items()}
    gather_ms = sum(v for k, v in breakdown.items() if k.startswith("gather"))
    n_gather_launches = sum(1 for k in breakdown if k.startswith("gather"))
    rows_per_step = rows_total / n_inst
    alg_bytes = rows_per_step * (8 * D + 8)  # 4D read + 4D written + id + location
    achieved = alg_bytes / (gather_ms * 1e-3) / 1e9
    peak, peak_kind = measured_peak()
    # roofline for the measured hit mix: per row, HBM moves 4D(l+p)+4D (own local reads + reads served
    # to peers + own writes), NVLink-in 4D*p, PCIe 4D*h; t_roof = max over the three links
    tsum = max(int(tiers.sum()), 1)
    fl, fp, fh = (float(x) / tsum for x in tiers)
    nvl_bw, pcie_bw = 770.0, 55.0  # GB/s: measured peer-copy figure of this pool (B200_PROFILING.md); PCIe Gen5 x16 practical
    t_parts = {"hbm": rows_per_step * (4 * D * (fl + fp) + 4 * D + 8) / (peak * 1e9),
               "nvlink": rows_per_step * 4 * D * fp / (nvl_bw * 1e9), "pcie": rows_per_step * 4 * D * fh / (pcie_bw * 1e9)}
    mix_bound = max(t_parts, key=t_parts.get)
    mix_frac = t_parts[mix_bound] / (gather_ms * 1e-3)

    # --- end-to-end: host seeds in (pinned), counters out, every step synchronised ---
    h_ids = torch.from_numpy(my_train[: B * train_steps].copy()).pin_memory()
    h_lab = lab[torch.from_numpy(my_train[: B * train_steps].astype(np.int64)).to(dev)].cpu().pin_memory()
    pin = pin_feat = None
    h_nc, h_ec = np.zeros(16, np.int32), np.zeros(16, np.int32)

    def e2e_step(s):
        c = (args.warmup + s) % train_steps
        p = params(args.warmup + s)
        dp.run_once_host(p, h_ids.numpy()[c * B:(c + 1) * B], h_lab.numpy()[c * B:(c + 1) * B], bufs[s % 2], h_nc, h_ec)

    # (a) latency view: one batch at a time, stream synchronised after every step
    for s in range(min(3, args.warmup)):
        e2e_step(s)
    barrier()
    t_a = time.perf_counter()
    for s in range(args.steps):
        e2e_step(s)
    torch.cuda.synchronize()
    t_sync = time.perf_counter() - t_a
    # (b) throughput view (the headline e2e): the same host-fed call without the per-step synchronisation, batches
    # round-robin over the in-flight runners exactly like the device-resident timed region.  Every step still does
    # its own H2D of seeds+labels (pinned) and its own D2H of both counter arrays (pinned).
    for r, _, _ in runners:
        r.set_overlap(args.overlap)
    h_cnt = torch.zeros((args.steps, 32), dtype=torch.int32).pin_memory()
    cnt_np = h_cnt.numpy()

    def e2e_async(first, count):
        cur = torch.cuda.current_stream()
        for _, _, st in runners[1:]:
            st.wait_stream(cur)
        for s in range(count):
            r, rb, st = runners[s % len(runners)]
            c = (first + s) % train_steps
            with torch.cuda.stream(st):
                r.run_once_host_async(params(first + s), h_ids.numpy()[c * B:(c + 1) * B], h_lab.numpy()[c * B:(c + 1) * B],
                                      rb[(s // len(runners)) % 2], cnt_np[s, :16], cnt_np[s, 16:])
        for _, _, st in runners[1:]:
            cur.wait_stream(st)

    e2e_async(args.warmup, min(4, args.steps))
    barrier()
    t_a = time.perf_counter()
    e2e_async(args.warmup, args.steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t_a
    assert int(cnt_np[args.steps - 1, 9]) == B and int(cnt_np[args.steps - 1, 8]) == H, "e2e counters not delivered"
    te = torch.tensor([t_e2e, t_sync], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e, t_sync = float(te[0].item()), float(te[1].item())
    dp.set_overlap(min(args.overlap, 1))
    # variant with the WHOLE result read back to pinned host memory (not what Legion's trainer does: it reads
    # the buffers in place over CUDA IPC) — reported for completeness on a few steps
    n_full = min(5, args.steps)
    pin = {k: torch.empty_like(getattr(bufs[0], k), device="cpu").pin_memory() for k in ("ids", "agg_src", "agg_dst")}
    pin_feat = torch.empty((feature_rows, D), dtype=torch.float32).pin_memory()
    barrier()
    t_a = time.perf_counter()
    d2h_full = 0
    for s in range(n_full):
        e2e_step(s)
        n, e = int(h_nc[9 + H]), int(h_ec[9 + H])
        b = bufs[s % 2]
        pin["ids"][:n].copy_(b.ids[:n], non_blocking=True)
        pin["agg_src"][:e].copy_(b.agg_src[:e], non_blocking=True)
        pin["agg_dst"][:e].copy_(b.agg_dst[:e], non_blocking=True)
        pin_feat[:n].copy_(b.features[:n], non_blocking=True)
        torch.cuda.synchronize()
        d2h_full += 4 * n + 8 * e + 4 * n * D + 128
    t_full = time.perf_counter() - t_a

    out = None
    if rank == 0:
        seeds = world * B * args.steps
        out = {
            "metric": "sampled+gathered seeds/sec", "value": seeds / (ms_max * 1e-3), "unit": "seeds/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 ids / fp32 rows moved bit-exact",
            "data": "synthetic",
            "config": {"workload": f"{shape['name']}-shaped synthetic graph (BASELINE.json " + {"products": "configs[1]", "paper100m": "configs[2] shape", "ukunion": "configs[3] shape, 128-d", "clueweb": "configs[4] shape"}[shape["name"]] + ")",
                       "num_nodes": N, "num_edges": E, "feature_dim": D, "fanout": fanout, "batch": B,
                       "scale": args.scale, "cache": f"Kc={world // kg},Kg={kg}: feature table interleaved by hotness rank over {kg} GPU(s) per clique" + (f", hottest {rep} rows replicated on every GPU" if rep else "")
                                + (f", hottest {args.cache_ratio:.3f} of the rows in HBM ({cap} rows/GPU), the rest in pinned host memory (UVA)" if feat_host else ", fully HBM-cached")
                                + (f"; topology in pinned host memory (UVA) with the hottest {args.topo_cache_ratio:.3f} of the adjacency lists cached in HBM ({topo_cap} rows/GPU)" if topo_host else "; topology replicated in HBM"),
                       "feature_cache_ratio": args.cache_ratio, "topology": args.topo, "topology_cache_ratio": args.topo_cache_ratio if topo_host else None,
                       "rng": "philox4x32-10", "position_map": ["dense u32[N]", "hashed L2-resident table"][dp.L.lg_sampler_dedup_layout(dp.sampler)], "gather_mover": args.gather, "gather_fusion": args.fuse, "batches_in_flight": args.inflight, "schedule": ["one stream", "gather overlaps next hop, joined per batch", "pipelined: every in-flight runner owns 2 INTERBATCH_CON buffer slots; the gather of batch k overlaps the sampling of the following batches"][args.overlap],
                       "l2": "working set (topology + features + per-batch output, >1.5 GB) exceeds the 126 MB L2; every step samples different seeds"},
            "e2e": {"value": seeds / t_e2e, "unit": "seeds/s", "h2d_bytes_per_step": 2 * 4 * B, "d2h_bytes_per_step": 128,
                    "note": "every step: seed ids+labels copied from pinned host memory (H2D), lg_run_batch_host_async, both counter "
                            "arrays copied back to pinned host memory (D2H, what get_next reads); batches in flight as in the device-"
                            "resident run; features/COO stay in the CUDA-IPC buffers as in Legion's hand-off (no host round trip)"},
            "e2e_sync_per_step": {"value": seeds / t_sync, "unit": "seeds/s",
                                  "note": "lg_run_batch_host: same copies, one batch at a time, stream synchronised after every step"},
            "e2e_host_result": {"value": world * B * n_full / t_full, "unit": "seeds/s", "steps": n_full,
                                "d2h_bytes_per_step": d2h_full // max(n_full, 1),
                                "note": "same, plus ids/COO/features copied back to pinned host memory every step"},
            "gpu_launches": 0,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": peak_kind, "kernel": f"feature gather ({n_gather_launches} launch(es) per step)",
                         "algorithmic_bytes_per_step": alg_bytes, "rows_per_step": rows_per_step,
                         "gather_ms_per_step": gather_ms,
                         "hit_mix": {"local": fl, "peer": fp, "host": fh, "bound": mix_bound, "frac_of_mix_roofline": mix_frac,
                                     "nvlink_GBps_assumed": nvl_bw, "pcie_GBps_assumed": pcie_bw}},
            "breakdown_ms": breakdown, "host_enqueue_ms_per_step": 1e3 * t_enq / args.steps,
            "features_bit_exact_selfcheck": features_ok,
            "tier_rows": {"local": int(tiers[0]), "peer": int(tiers[1]), "host_or_backing": int(tiers[2])},
            "clocks": clk,
        }
        # lg_run_batch: batch_generate + (sample, rank) per hop + the last hop's relabel (which also releases the
        # position map) + gathers; no memset nodes
        if dp.L.lg_sampler_dedup_layout(dp.sampler) == 1:  # hashed: + the seeds' local ids
            out["gpu_launches"] = args.steps * (2 + 2 * H + 1 + n_gather_launches)
        else:
            out["gpu_launches"] = args.steps * (1 + 2 * H + 1 + n_gather_launches)
        tr = recorded_traffic()
        if tr:
            out["roofline"]["traffic"] = tr.get("traffic_bytes_per_step")
            out["roofline"]["traffic_source"] = tr.get("source")
    # --- CPU baseline beside it (rank 0, N=1 only) ---
    if rank == 0 and world == 1 and not args.no_cpu_baseline and feat is not None:
        if not topo_host:
            np_ip, np_ix = ip.cpu().numpy(), ix.cpu().numpy()
        out["cpu_baseline"] = cpu_arm(shape, np_ip, np_ix, feat.cpu().numpy(), my_train, steps=args.cpu_steps,
                                      warmup=1)["cpu_baseline"]
    # --- the drop-in boundary itself (rank 0, N=1): sampling_server binary -> shm/semaphores/CUDA IPC -> ipc_service ---
    if rank == 0 and world == 1 and args.server_e2e and shape["name"] == "products" and not feat_host and not topo_host:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "server_e2e.py"), "--epochs", "10"],
                               capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                j = json.loads(line[-1])
                out["e2e_server"] = {"value": j["seeds_per_s"], "unit": "seeds/s", "ms_per_batch": j["ms_per_batch"],
                                     "steps": j["train_steps_per_epoch"] * j["epochs"], "server_says": j["server_says"],
                                     "note": "C++ sampling_server binary (dataset files, meta_config, presampling, cost model, cache fill) -> "
                                             "simpleIPCshm + semaphores + CUDA-IPC buffers -> ipc_service.get_next/get_block_size/"
                                             "synchronize consumer, wall clock over the training steps of 10 epochs"}
            else:
                out["e2e_server"] = {"unavailable": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001  (never lose the main line to the optional leg)
            out["e2e_server"] = {"unavailable": repr(ex)[:300]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port of the baseline BASELINE.json names): DGL-semantics sampler + index_select
# ----------------------------------------------------------------------------------------------
def cpu_arm(shape, indptr, indices, feat, train, steps, warmup):
    from oracle import oracle as O
    B, fanout, D = shape["batch"], shape["fanout"], shape["D"]
    base = O.DGLBaseline(indptr, indices, fanout, B)
    base.use_all_cores()  # torchrun sets OMP_NUM_THREADS=1; the CPU arm gets every host core
    out = np.empty((O.num_ids(B, fanout), D), np.float32)
    n_batches = (len(train) - 1) // B
    times, rows = [], 0
    for s in range(warmup + steps):
        seeds = train[(s % n_batches) * B:(s % n_batches + 1) * B]
        t = time.perf_counter()
        n = base.sample(seeds, rng_seed=SEED + s)
        base.gather(feat, n, out)
        dt = time.perf_counter() - t
        if s >= warmup:
            times.append(dt)
            rows += n
    tot = sum(times)
    val = B * steps / tot
    return {"value": val, "ms_per_step": 1e3 * tot / steps,
            "cpu_baseline": {"value": val, "unit": "seeds/s", "cores": base.threads(), "kind": "port",
                             "sample": f"{steps} batches of {B} seeds, fanout {fanout}, DGL-semantics sampler (without replacement, "
                                       f"unique frontier) + index_select over {rows // max(steps, 1)} rows/batch, OpenMP {base.threads()} threads",
                             "gather_GBps": rows * (8 * D + 8) / tot / 1e9}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = shape_of(args)
    t0 = time.time()
    indptr = indices = feat = None
    try:
        import torch
        if torch.cuda.is_available():  # data generation only; the timed path below is pure CPU
            ip, ix, f, _, _ = device_dataset(shape, 0)
            indptr, indices, feat = ip.cpu().numpy(), ix.cpu().numpy(), f.cpu().numpy()
            del ip, ix, f
    except Exception as ex:  # noqa: BLE001
        log(f"[reference] device generator unavailable ({ex}); using numpy")
    if indptr is None:
        from legion_b200 import synth
        indptr, indices = synth.graph(shape["N"], shape["dmin"], shape["dmax"], SEED)
        feat = np.concatenate([synth.features(r, min(65536, shape["N"] - r), shape["D"], SEED)
                               for r in range(0, shape["N"], 65536)])
    train = train_split(shape, 1)[0]
    log(f"[reference] dataset ready in {time.time() - t0:.1f}s")
    r = cpu_arm(shape, indptr, indices, feat, train, steps=args.steps, warmup=args.warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "sampled+gathered seeds/sec", "value": r["value"], "unit": "seeds/s",
           "n_gpus": max(args.gpus, world), "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 ids / fp32 rows", "data": "synthetic",
           "config": {"workload": f"{shape['name']}-shaped synthetic graph (BASELINE.json configs[0]: CPU sampler + index_select)",
                      "num_nodes": shape["N"], "num_edges": int(indptr[-1]), "feature_dim": shape["D"],
                      "fanout": shape["fanout"], "batch": shape["batch"], "scale": args.scale},
           "cpu_baseline": r["cpu_baseline"],
           "e2e": {"value": r["value"], "unit": "seeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale N and E of the named shape (1.0 = paper size)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--presample", type=int, default=20, help="presampling batches used for the hotness ranking")
    ap.add_argument("--gather", default="auto", choices=["auto", "ldg", "tma"])
    ap.add_argument("--fuse", type=int, default=2, choices=[0, 1, 2],
                    help="gather launches per batch: 0 one per lookup op, 1 seeds ride with hop 1, 2 single gather")
    ap.add_argument("--kg", type=int, default=0, help="GPUs sharing one partitioned cache (0 = auto by capacity)")
    ap.add_argument("--cache-gb", type=float, default=100.0, help="per-GPU feature-cache budget used by --kg auto")
    ap.add_argument("--replicate-ratio", type=float, default=0.0,
                    help="with --kg > 1: fraction of the rows (hottest first) stored on every GPU instead of partitioned")
    ap.add_argument("--cache-ratio", type=float, default=1.0,
                    help="fraction of the feature rows cached in HBM across the clique; < 1 puts the backing matrix in "
                         "pinned host memory and misses are read over PCIe (UVA)")
    ap.add_argument("--topo", default="hbm", choices=["hbm", "host"],
                    help="where the full CSR lives: replicated in HBM, or pinned host memory read through UVA")
    ap.add_argument("--topo-cache-ratio", type=float, default=0.0,
                    help="with --topo host: fraction of the vertices whose adjacency lists are cached in HBM")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight per GPU (own scratch + stream each); 3 measured best: host-fed e2e 36.3 -> 40.4 M seeds/s")
    ap.add_argument("--overlap", type=int, default=2, choices=[0, 1, 2],
                    help="0 one stream; 1 gathers overlap the next hop; 2 pipelined across the two batch slots")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-server-e2e", dest="server_e2e", action="store_false",
                    help="skip the e2e_server leg (scripts/server_e2e.py: the sampling_server binary feeding an ipc_service "
                         "consumer over shm/semaphores/CUDA IPC; N=1, products workload only, ~30 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 50 and "--steps" not in sys.argv:
            args.steps = 20
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
