#!/usr/bin/env python
"""bench.py — sampled+gathered seeds/s of the Legion mini-batch data path on B200.

A step = one mini-batch of `batch` seeds through the whole hot path on one GPU: batch_generate -> per hop
(neighbour sampling -> dedup/reindex) -> unified-cache lookup + feature gather of every vertex of the batch, i.e.
the ops of GPURunner::RunOnce (reference engine/server.cu:302-332).

  python bench.py [--gpus N --steps K --warmup W]     our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                 CPU arm: DGL-semantics sampler + index_select (oracle port)

Default workload = the graph BASELINE.json's metric is quoted on: UK-Union shape (133.6 M vertices, 5.5 B edges,
128-d), fan-out [25,10], batch 8000, feature table partitioned over the N GPUs of the NVSwitch domain
(Kc=1, Kg=N — legion_server.py:99-106 -> cache_agg_mode = log2(clique) -> cache/cache.cu:375-392).  The
replicated (Kg=1) and hybrid layouts and the products shape (BASELINE.json configs[1]) ride along as extra keys.

One JSON line on stdout (rank 0), printed last.  See DESIGN.md "Measurement" for the byte accounting.
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
    # name: (shape key, fanout, batch, dmax, materialise the [N x D] matrix in vertex order on the device?)
    "products": ("products", [25, 10], 8000, 20000, True),     # BASELINE.json configs[1]
    "paper100m": ("paper100m", [25, 10], 8000, 20000, False),  # configs[2] shape
    "ukunion": ("ukunion", [25, 10], 8000, 20000, False),      # configs[3] shape: the metric's named graph (default)
    "clueweb": ("clueweb", [15, 10, 5], 8000, 20000, False),   # configs[4] shape: GCN 3-hop, topology in host UVA
}
CONFIG_OF = {"products": "configs[1]", "paper100m": "configs[2] shape", "ukunion": "configs[3] shape, 128-d",
             "clueweb": "configs[4] shape"}
NVLINK_GUIDE_GBPS = 770.0  # B200_PROFILING.md: measured peer copy per direction on this pool (900 nominal)
PCIE_GUIDE_GBPS = 55.0     # PCIe Gen5 x16 practical; replaced by the H2D copy measured in this run


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_label(name):
    """identical in both arms (the driver compares the strings)"""
    return f"{name}-shaped synthetic graph (BASELINE.json {CONFIG_OF[name]})"


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
        # "under load": the samples of the busy part of the run (the upper half of the observed clocks)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def recorded_traffic(key):
    """ncu dram__bytes_read+write of the gather launch for this workload/layout (profiles/roofline_traffic.json), or None"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# dataset
# ----------------------------------------------------------------------------------------------
def shape_of(args, workload=None):
    from legion_b200 import synth
    key, fanout, batch, dmax, dense = WORKLOADS[workload or args.workload]
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
    """random 10 % of the vertices (dataset/gen_sets.py:62-67), split by id % gpus (storage_management.cu:175-179).
    Pure numpy (seeded), so both arms draw the same sets with or without a GPU."""
    from legion_b200 import synth
    tr, _, _ = synth.split_sets(shape["N"], SEED)
    return synth.partition_ids(tr, world)


class HostCSR:
    """The full CSR in host memory, for the checker (oracle) and the CPU baseline.  With several ranks on one box the
    copy lives once in /dev/shm: every rank writes its slice of the arrays from its (identical) device copy and maps the
    whole file; the names are unlinked as soon as everybody has them mapped."""

    def __init__(self, ip, ix, N, E, rank, world, dist):
        import torch
        if world == 1:
            self.indptr = np.empty(N + 1, np.int64)
            self.indices = np.empty(max(E, 1), np.int32)[:E]
            self._copy(ip, self.indptr, 0, N + 1)
            self._copy(ix, self.indices, 0, E)
            return
        tok = [f"/dev/shm/legion_b200_csr_{os.getpid()}_{int(time.time())}" if rank == 0 else None]
        dist.broadcast_object_list(tok, src=0)
        base = tok[0]
        if rank == 0:
            for suffix, nbytes in ((".indptr", (N + 1) * 8), (".indices", max(E, 1) * 4)):
                with open(base + suffix, "wb") as f:
                    f.truncate(nbytes)
        dist.barrier()
        self.indptr = np.memmap(base + ".indptr", dtype=np.int64, mode="r+", shape=(N + 1,))
        self.indices = np.memmap(base + ".indices", dtype=np.int32, mode="r+", shape=(max(E, 1),))[:E]
        dist.barrier()  # everybody holds a mapping: the names can go (a crash later leaves nothing behind)
        if rank == 0:
            os.unlink(base + ".indptr")
            os.unlink(base + ".indices")
        self._copy(ip, self.indptr, (N + 1) * rank // world, (N + 1) * (rank + 1) // world)
        self._copy(ix, self.indices, E * rank // world, E * (rank + 1) // world)  # .cpu() copies are synchronous
        dist.barrier()

    @staticmethod
    def _copy(src, dst, lo, hi, chunk=1 << 27):
        for a in range(lo, hi, chunk):
            b = min(a + chunk, hi)
            dst[a:b] = src[a:b].cpu().numpy()


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class Ctx:
    """everything one workload needs on one rank"""


def setup_workload(args, workload, rank, world, local, dist):
    import torch
    from legion_b200 import capi
    from legion_b200.runner import DataPath, SharedHostBuffer

    c = Ctx()
    c.args, c.rank, c.world, c.local, c.dist = args, rank, world, local, dist
    c.dev = dev = f"cuda:{local}"
    c.shape = shape = shape_of(args, workload)
    N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
    c.N, c.D, c.B, c.fanout, c.H = N, D, B, fanout, len(fanout)
    t0 = time.time()
    c.ip, c.ix, c.feat, c.lab, c.E = device_dataset(shape, local)
    parts = train_split(shape, world)
    c.my_train = parts[rank]
    c.d_train = torch.from_numpy(c.my_train).to(dev)
    c.d_lab = c.lab[c.d_train.long()].contiguous()
    c.train_steps = (min(len(p) for p in parts) - 1) // B  # engine/ipc_service.cu:73-82
    assert c.train_steps >= 1, "training set smaller than one batch"
    if rank == 0:
        log(f"[bench] {shape['name']} N={N} E={c.E} D={D} world={world} train/gpu={len(c.my_train)} "
            f"train_steps={c.train_steps} setup {time.time() - t0:.1f}s")

    c.dp = dp = DataPath(local, fanout, B, N, D, rank=rank, world=world)
    c.topo_host = args.topo == "host"
    c.host_keep = []
    if c.topo_host:
        # full CSR in cudaHostAllocMapped memory, read by the sampler through UVA (storage/storage_management.cu:100-115,
        # engine/operator_impl.cu:224-243); the device copy only feeds presampling/placement and is dropped afterwards
        # (one copy per box: every rank writes its slice of the arrays from its identical device copy)
        dd = dist if world > 1 else None
        c.h_ip = SharedHostBuffer((N + 1) * 8, rank, world, dd, "indptr")
        c.h_ix = SharedHostBuffer(max(c.E, 1) * 4, rank, world, dd, "indices")
        for hb, t, n_el, item in ((c.h_ip, c.ip, N + 1, 8), (c.h_ix, c.ix, c.E, 4)):
            lo, hi = n_el * rank // world, n_el * (rank + 1) // world
            hb.copy_from_device(t.data_ptr() + lo * item, lo * item, (hi - lo) * item, dp._stream())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        c.host_keep += [c.h_ip, c.h_ix]
    dp.set_full_graph(c.ip.data_ptr(), c.ix.data_ptr(), keep=[c.ip, c.ix])  # topology replicated in each GPU's HBM
    c.feat_host = args.cache_ratio < 1.0
    if c.feat_host:
        # backing matrix in pinned host memory (cache/cache_impl.cuh:262-266 reads misses through UVA)
        h_feat = SharedHostBuffer(N * D * 4, rank, world, dist if world > 1 else None, "features")
        lo, hi = N * rank // world, N * (rank + 1) // world  # each rank generates its share of the rows
        for r0 in range(lo, hi, 1 << 22):  # generated by the device straight into the mapped allocation
            capi.check(dp.L.lg_synth_features(dp._stream(), r0, min(1 << 22, hi - r0), D, SEED,
                                              C.c_void_p(h_feat.dev_ptr + r0 * D * 4)))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        c.host_keep.append(h_feat)
        dp.set_backing_features(h_feat.dev_ptr, keep=[h_feat])
    elif c.feat is not None:
        dp.set_backing_features(c.feat.data_ptr(), keep=[c.feat])
    else:
        dp._backing = 0
    dp.set_overlap(args.overlap)
    c.variant = {"auto": capi.GATHER_AUTO, "ldg": capi.GATHER_LDG, "tma": capi.GATHER_TMA}[args.gather]
    dp.set_gather_variant(c.variant)

    # --- presampling: hotness -> ranking (PreSc + CandidateSelection, engine/server.cu:90-117, cache/cache.cu:360-443) ---
    scratch = dp.alloc_batch(feature_rows=1)
    eh = torch.zeros(N, dtype=torch.int64, device=dev)
    nh = torch.zeros(N, dtype=torch.int64, device=dev)
    mx = torch.zeros(1, dtype=torch.int32, device=dev)
    for it in range(min(args.presample, c.train_steps)):
        dp.run_presc(dp.params(c.d_train, c.d_lab, B, it, seed=SEED, batch_id=it), scratch, eh, nh, mx)
    torch.cuda.synchronize()
    if world > 1:
        dist.all_reduce(nh)  # init-time sum of hotness over the GPUs (cache/cache.cu:408-411); not on the serving path
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    c.order, _ = dp.rank_hotness(nh)
    c.topo_cap = 0
    if c.topo_host:
        # hot-vertex topology cache in HBM (GraphCache, storage/graph_storage.cu:76-111), ranked by the presampled
        # edge hotness (QT, cache/cache.cu:420-440); everything else is read from the host CSR over PCIe
        kg_t = world
        if world > 1:
            dist.all_reduce(eh)
        if args.topo_cache_ratio > 0:
            order_t, _ = dp.rank_hotness(eh)
            c.topo_cap = max(1, (int(N * min(args.topo_cache_ratio, 1.0)) + kg_t - 1) // kg_t)
            dp.build_topology_cache(order_t, c.topo_cap, kg=kg_t, j=rank % kg_t, dist=dist if world > 1 else None)
            del order_t
        dp.repoint_full_graph(c.h_ip.dev_ptr, c.h_ix.dev_ptr, drop=[c.ip, c.ix])
        c.ip = c.ix = None  # the HBM copy of the CSR is gone from here on
    c.max_ids = int(mx.item())
    c.feature_rows = min(dp.num_ids, int(c.max_ids * 1.2) + 1)  # engine/server.cu:277
    del scratch, eh, nh
    torch.cuda.empty_cache()
    c.bufs = [dp.alloc_batch(feature_rows=c.feature_rows) for _ in range(2)]  # INTERBATCH_CON pipeline slots
    dp.set_gather_fusion(args.fuse)
    # additional batches in flight on the same GPU: own sampler scratch + buffers + stream, shared storage
    c.runners = [(dp, c.bufs, torch.cuda.current_stream())]
    for _ in range(args.inflight - 1):
        d2 = DataPath(local, fanout, B, N, D, rank=rank, world=world)
        d2.share_storage_from(dp)
        d2.set_overlap(args.overlap)
        d2.set_gather_fusion(args.fuse)
        d2.set_gather_variant(c.variant)
        c.runners.append((d2, [d2.alloc_batch(feature_rows=c.feature_rows) for _ in range(2)], torch.cuda.Stream()))
    return c


def build_cache(c, kg, replicate_ratio):
    """FillUp (cache/cache.cu:553-611) for one layout: the table interleaved by hotness rank over the kg GPUs of a clique
    (reference placement), optionally with the hottest rows replicated on every GPU (hybrid, an extension)."""
    import torch
    args, dp, dist = c.args, c.dp, c.dist
    N = c.N
    assert c.world % kg == 0, "world size must be a multiple of Kg"
    if dp._cache_keep:
        if c.world > 1:
            torch.cuda.synchronize()
            dist.barrier()  # nobody frees a shard a peer may still read
        dp.drop_feature_cache()
        torch.cuda.empty_cache()
    cached_rows = int(N * min(args.cache_ratio, 1.0))
    rep = min(int(N * max(replicate_ratio, 0.0)), cached_rows) if kg > 1 else 0
    cap = max(rep + (cached_rows - rep + kg - 1) // kg, 1)
    dd = dist if c.world > 1 else None
    identity = bool(args.identity and kg == 1 and cached_rows >= N and not c.feat_host)
    if identity:  # the cache holds every vertex on this GPU: rows at row index = vertex id, no directory (LG_CACHE_IDENTITY)
        dp.build_feature_cache_identity(seed=SEED if c.feat is None else None)
    elif c.feat is not None and not c.feat_host:
        dp.build_feature_cache(c.order, cap, kg=kg, j=c.rank % kg, dist=dd, replicate=rep)
    else:  # shards generated in place (paper-scale shapes: no [N x D] matrix in vertex order in HBM)
        dp.build_feature_cache_synth(c.order, cap, SEED, kg=kg, j=c.rank % kg, dist=dd, keep_backing=c.feat_host,
                                     replicate=rep)
    for r, _, _ in c.runners[1:]:
        r.share_storage_from(dp)
    c.kg, c.rep, c.cap = kg, rep, cap
    c.layout = ((f"Kc={c.world // kg},Kg=1: the whole feature table on every GPU, rows at row index = vertex id (no directory lookup)"
                 if identity else
                 f"Kc={c.world // kg},Kg={kg}: feature table interleaved by hotness rank over {kg} GPU(s) per clique")
                + (f", hottest {rep} rows replicated on every GPU (hybrid placement)" if rep else "")
                + (f", hottest {args.cache_ratio:.3f} of the rows in HBM ({cap} rows/GPU), the rest in pinned host memory (UVA)"
                   if c.feat_host else ", fully HBM-cached")
                + (f"; topology in pinned host memory (UVA) with the hottest {args.topo_cache_ratio:.3f} of the adjacency lists "
                   f"cached in HBM ({c.topo_cap} rows/GPU)" if c.topo_host else "; topology replicated in HBM"))


def params_of(c, step):
    return c.dp.params(c.d_train, c.d_lab, c.B, step % c.train_steps, seed=SEED, batch_id=step)


def run_steps(c, first, count, tier=False):
    """`count` batches round-robin over the in-flight runners; returns when all are enqueued and joined into the
    current stream"""
    import torch
    cur = torch.cuda.current_stream()
    for _, _, st in c.runners[1:]:
        st.wait_stream(cur)
    n = len(c.runners)
    for s in range(count):
        r, rb, st = c.runners[s % n]
        with torch.cuda.stream(st):
            r.run_once(params_of(c, first + s), rb[(s // n) % 2], tier=tier)
    for r, rb, st in c.runners:
        with torch.cuda.stream(st):
            for b in rb:
                r.batch_wait(b)
        if st is not cur:
            cur.wait_stream(st)


def barrier(c):
    import torch
    if c.world > 1:
        c.dist.barrier()
    torch.cuda.synchronize()


def features_check(c, buf):
    """rows gathered into `buf` == the feature function of their ids, bit for bit (all rows, on the device)"""
    import torch
    from legion_b200 import capi
    n0 = int(buf.node_counter[9 + c.H].item())
    chk = torch.empty((n0, c.D), dtype=torch.float32, device=c.dev)
    capi.check(c.dp.L.lg_synth_feature_rows(c.dp._stream(), buf.ids.data_ptr(), n0, c.D, SEED, chk.data_ptr()))
    torch.cuda.synchronize()
    ok = bool(torch.equal(chk.view(torch.int32), buf.features[:n0].view(torch.int32)))
    del chk
    return ok, n0


def parity_selfcheck(c, host_csr, step):
    """Untimed, every rank: one batch of this rank's own stream compared field by field with the CPU oracle (counters —
    all 16+16 slots —, ids, labels, COO) and its gathered rows with the feature function (all rows on the device; a sample
    of them against the numpy twin on the host).  Reference semantics: engine/operator_impl.cu:27-296,
    cache/cache_impl.cuh:239-272.  The run fails on any mismatch."""
    import torch
    from legion_b200 import synth
    from oracle import oracle as O
    dp, buf = c.dp, c.bufs[0]
    p = params_of(c, step)
    tier0 = dp.tier_rows.clone()
    dp.run_once(p, buf, tier=True)
    dp.batch_wait(buf)
    torch.cuda.synchronize()
    tier = (dp.tier_rows - tier0).cpu().numpy().astype(np.int64)
    dp.tier_rows.copy_(tier0)
    nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
    n_g, e_g = int(nc[9 + c.H]), int(ec[9 + c.H])
    got = dict(nc=nc, ec=ec, ids=buf.ids[:n_g].cpu().numpy(), labels=buf.labels[:int(nc[9])].cpu().numpy(),
               agg_src=buf.agg_src[:e_g].cpu().numpy(), agg_dst=buf.agg_dst[:e_g].cpu().numpy())
    t0 = time.time()
    orc = O.Oracle(host_csr.indptr, host_csr.indices, c.fanout, c.B)
    want = orc.run_batch(c.my_train, c.lab_host, c.B, step % c.train_steps, rng_kind=O.RNG_PHILOX, seed=SEED,
                         batch_id=step, stream_id=c.rank)
    del orc
    n, e = want["total_nodes"], want["total_edges"]
    res = {"counters": bool(np.array_equal(got["nc"], want["nc"]) and np.array_equal(got["ec"], want["ec"])),
           "ids": bool(np.array_equal(got["ids"], want["ids"][:n])),
           "labels": bool(np.array_equal(got["labels"], want["labels"][:int(want["nc"][9])])),
           "coo": bool(np.array_equal(got["agg_src"], want["agg_src"][:e]) and np.array_equal(got["agg_dst"], want["agg_dst"][:e]))}
    f_ok, _ = features_check(c, buf)
    rng = np.random.default_rng(SEED + c.rank)
    rows = np.sort(rng.choice(n, size=min(n, 2048), replace=False))
    sample = buf.features[torch.from_numpy(rows).to(c.dev)].cpu().numpy()
    ids = want["ids"][rows].astype(np.uint64)
    idx = (ids[:, None] * np.uint64(c.D) + np.arange(c.D, dtype=np.uint64)[None, :]).reshape(-1)
    twin = ((synth.hash2(SEED ^ synth._S3, idx) & np.uint64(0xFFFFFFFF)).astype(np.uint32) & np.uint32(0xBFFFFFFF)).reshape(len(rows), c.D)
    res["features"] = bool(f_ok and np.array_equal(sample.view(np.uint32), twin))
    res["status"] = dp.status() == 0
    ok = all(res.values())
    info = {"ok": ok, "fields": res, "batch_id": step, "nodes": int(n), "edges": int(e),
            "tier_rows_of_checked_batch": {"local": int(tier[0]), "peer": int(tier[1]), "host_or_backing": int(tier[2])},
            "oracle_s": round(time.time() - t0, 2)}
    return info


def timed_region(c, steps, warmup, first=0):
    """warm-up, then exactly `steps` steps: barrier + synchronize on both sides, CUDA events, max over ranks"""
    import torch
    for r, _, _ in c.runners:
        r.set_overlap(c.args.overlap)
        r.tier_rows.zero_()
    torch.cuda.synchronize()
    run_steps(c, first, warmup)
    torch.cuda.synchronize()
    assert all(r.status() == 0 for r, _, _ in c.runners), "sampler overflow status"
    L = c.dp.L
    barrier(c)
    L.lg_debug_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_enq = time.perf_counter()
    run_steps(c, first + warmup, steps, tier=True)  # all in-flight batches are complete before the clock stops
    t_enq = time.perf_counter() - t_enq  # host time to enqueue the steps (no synchronisation inside)
    e1.record()
    barrier(c)
    launches = int(L.lg_debug_launch_count(0))
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(tms, op=c.dist.ReduceOp.MAX)
    tiers = sum(r.tier_rows for r, _, _ in c.runners)
    if c.world > 1:
        c.dist.all_reduce(tiers)
    tiers = tiers.cpu().numpy().astype(np.int64)
    assert all(r.status() == 0 for r, _, _ in c.runners), "sampler overflow status"
    return float(tms.item()), tiers, t_enq, launches


def instrumented_pass(c, steps, first):
    """same steps, one batch at a time, CUDA events around every op: per-kernel durations of this rank"""
    import torch
    from legion_b200 import capi
    args, dp, H = c.args, c.dp, c.H
    dp.set_overlap(min(args.overlap, 1))
    L, st = dp.L, dp._stream()
    rows_total, acc = 0, {}
    n_inst = min(steps, 20)
    for s in range(n_inst):
        p = params_of(c, first + s)
        b = c.bufs[s % 2]
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
        ops.append(("io_complete", lambda: L.lg_io_complete(dp.sampler, st, 0, C.byref(b.c), None, None)))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
        evs[0].record()
        for i, (nm, fn) in enumerate(ops):
            capi.check(fn())
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, (nm, _) in enumerate(ops):
            acc[nm] = acc.get(nm, 0.0) + evs[i].elapsed_time(evs[i + 1])
        rows_total += int(b.node_counter[9 + H].item())
    breakdown = {k: v / n_inst for k, v in acc.items()}
    gather_ms = sum(v for k, v in breakdown.items() if k.startswith("gather"))
    n_gather = sum(1 for k in breakdown if k.startswith("gather"))
    return breakdown, gather_ms, n_gather, rows_total / n_inst


def link_peaks(c):
    """Measured on this box, all ranks at once: pinned H2D copy (PCIe) and, with a partitioned cache, a bulk copy out of a
    peer's shard (NVLink).  GB/s, min over ranks."""
    import torch
    from legion_b200 import capi
    out = {}
    n = 1 << 28
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=c.dev)
    best = 0.0
    barrier(c)
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    vals = [best, 0.0]
    del h
    if c.kg > 1:
        nb = min(c.cap * c.D * 4, 1 << 30)
        dst = torch.empty(nb, dtype=torch.uint8, device=c.dev)
        peer = c.dp.cache.shard[(c.dp.local_part + 1) % c.kg]
        barrier(c)
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.check(c.dp.L.lg_memcpy_d2d(C.c_void_p(dst.data_ptr()), C.c_void_p(peer), nb, c.dp._stream()))
            e1.record()
            torch.cuda.synchronize()
            vals[1] = max(vals[1], nb / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del dst
    t = torch.tensor(vals, dtype=torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MIN)
    out["pcie_h2d_GBps"] = float(t[0].item())
    out["nvlink_peer_copy_GBps"] = float(t[1].item()) if c.kg > 1 else None
    del d
    return out


def mix_roofline(c, tiers, rows_per_step, gather_ms, peaks, hbm_peak):
    """roofline of the gather for the measured hit mix: per row HBM moves 4D(l+p)+4D+8 (own local reads + reads
    served to peers + own writes + id + location), NVLink-in 4D*p, PCIe 4D*h; t_roof = max over the three links"""
    D = c.D
    tsum = max(int(tiers.sum()), 1)
    fl, fp, fh = (float(x) / tsum for x in tiers)
    nvl = peaks.get("nvlink_peer_copy_GBps") or NVLINK_GUIDE_GBPS
    pcie = peaks.get("pcie_h2d_GBps") or PCIE_GUIDE_GBPS
    bytes_per_row = {"hbm": 4 * D * (fl + fp) + 4 * D + 8, "nvlink": 4 * D * fp, "pcie": 4 * D * fh}
    bw = {"hbm": hbm_peak, "nvlink": nvl, "pcie": pcie}
    t_parts = {k: rows_per_step * bytes_per_row[k] / (bw[k] * 1e9) for k in bw}
    bound = max(t_parts, key=t_parts.get)
    t = gather_ms * 1e-3
    return {"local": fl, "peer": fp, "host": fh, "bound": bound, "frac_of_mix_roofline": t_parts[bound] / t,
            "link_GBps_achieved": {k: rows_per_step * bytes_per_row[k] / t / 1e9 for k in bw},
            "link_GBps_peak": bw,
            "peaks_source": {"hbm": "MEASURED_PEAKS.json", "nvlink": "peer copy measured in this run" if peaks.get("nvlink_peer_copy_GBps") else "B200_PROFILING.md (770 GB/s peer copy)",
                             "pcie": "pinned H2D copy measured in this run"}}


def e2e_region(c, steps, first):
    """host-fed end to end: every step copies its seeds + labels from pinned host memory, runs the batch and copies both
    counter arrays back to pinned host memory; batches in flight as in the device-resident run; wall clock"""
    import torch
    B, H, dp = c.B, c.H, c.dp
    n_seed = B * c.train_steps
    h_ids = torch.from_numpy(c.my_train[:n_seed].copy()).pin_memory()
    h_lab = torch.from_numpy(c.lab_host[:n_seed].copy()).pin_memory()
    ids_np, lab_np = h_ids.numpy(), h_lab.numpy()
    h_nc, h_ec = np.zeros(16, np.int32), np.zeros(16, np.int32)
    dp.set_overlap(min(c.args.overlap, 1))

    def e2e_step(s):
        k = (first + s) % c.train_steps
        dp.run_once_host(params_of(c, first + s), ids_np[k * B:(k + 1) * B], lab_np[k * B:(k + 1) * B], c.bufs[s % 2], h_nc, h_ec)

    # (a) latency view: one batch at a time, stream synchronised after every step
    for s in range(3):
        e2e_step(s)
    barrier(c)
    t_a = time.perf_counter()
    for s in range(steps):
        e2e_step(s)
    torch.cuda.synchronize()
    t_sync = time.perf_counter() - t_a
    # (b) throughput view (the headline e2e): the same host-fed call without the per-step synchronisation
    for r, _, _ in c.runners:
        r.set_overlap(c.args.overlap)
    h_cnt = torch.zeros((steps, 32), dtype=torch.int32).pin_memory()
    cnt_np = h_cnt.numpy()
    n = len(c.runners)

    def e2e_async(count):
        cur = torch.cuda.current_stream()
        for _, _, st in c.runners[1:]:
            st.wait_stream(cur)
        for s in range(count):
            r, rb, st = c.runners[s % n]
            k = (first + s) % c.train_steps
            with torch.cuda.stream(st):
                r.run_once_host_async(params_of(c, first + s), ids_np[k * B:(k + 1) * B], lab_np[k * B:(k + 1) * B],
                                      rb[(s // n) % 2], cnt_np[s, :16], cnt_np[s, 16:])
        for _, _, st in c.runners[1:]:
            cur.wait_stream(st)

    e2e_async(min(4, steps))
    barrier(c)
    t_a = time.perf_counter()
    e2e_async(steps)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t_a
    assert int(cnt_np[steps - 1, 9]) == B and int(cnt_np[steps - 1, 8]) == H, "e2e counters not delivered"
    te = torch.tensor([t_e2e, t_sync], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(te, op=c.dist.ReduceOp.MAX)
    dp.set_overlap(min(c.args.overlap, 1))
    return float(te[0].item()), float(te[1].item())


def measure_layout(c, steps, warmup, peaks, hbm_peak, full):
    """timed region (+ instrumented pass, + e2e when `full`) for the cache layout currently built"""
    ms_max, tiers, t_enq, launches = timed_region(c, steps, warmup)
    f_ok, _ = features_check(c, c.runners[(steps - 1) % len(c.runners)][1][((steps - 1) // len(c.runners)) % 2])
    assert f_ok, "gathered features differ from the feature function"
    breakdown, gather_ms, n_gather, rows_per_step = instrumented_pass(c, steps, warmup)
    mix = mix_roofline(c, tiers, rows_per_step, gather_ms, peaks, hbm_peak)
    seeds = c.world * c.B * steps
    res = {"value": seeds / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "cache": c.layout,
           "tier_rows": {"local": int(tiers[0]), "peer": int(tiers[1]), "host_or_backing": int(tiers[2])},
           "hit_mix": mix, "gather_ms_per_step": gather_ms, "rows_per_step": rows_per_step,
           "breakdown_ms": breakdown, "features_bit_exact_selfcheck": f_ok, "gpu_launches": launches,
           "host_enqueue_ms_per_step": 1e3 * t_enq / steps, "gather_launches_per_step": n_gather}
    if full:
        t_e2e, t_sync = e2e_region(c, steps, warmup)
        res["e2e_value"] = seeds / t_e2e
        res["e2e_sync_value"] = seeds / t_sync
    return res


def roofline_of(c, res, hbm_peak, peak_kind, traffic_key):
    """the dominant kernel = the feature gather; bound = the link that binds it for the measured hit mix"""
    D = c.D
    mix = res["hit_mix"]
    t = res["gather_ms_per_step"] * 1e-3
    rows = res["rows_per_step"]
    alg = rows * (8 * D + 8)  # 4D read + 4D written + id + location
    bound = mix["bound"]
    hbm = {"achieved": alg / t / 1e9, "peak": hbm_peak, "frac": alg / t / 1e9 / hbm_peak, "peak_kind": peak_kind,
           "algorithmic_bytes_per_launch": alg, "note": "8D+8 bytes per gathered row over the gather launch"}
    if bound == "hbm":
        r = {"bound": "hbm", "achieved": hbm["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm["frac"]}
    else:
        r = {"bound": bound, "achieved": mix["link_GBps_achieved"][bound], "peak": mix["link_GBps_peak"][bound], "unit": "GB/s",
             "frac": mix["frac_of_mix_roofline"],
             "bound_note": f"the gather of this hit mix is bound by {bound}: achieved = bytes that cross that link per launch / launch time"}
    tr = recorded_traffic(traffic_key)
    traffic = None
    if tr:  # the capture's bytes per gathered row x the rows of THIS launch (the capture may be of a scaled-down graph)
        traffic = tr["bytes_per_row"] * rows if "bytes_per_row" in tr else tr["bytes_per_launch"]
    r.update({"traffic": traffic, "traffic_source": tr["source"] if tr else None,
              "traffic_key": traffic_key, "peak_kind": peak_kind if bound == "hbm" else mix["peaks_source"][bound],
              "kernel": f"feature gather ({res['gather_launches_per_step']} launch(es) per step)",
              "algorithmic_bytes_per_step": alg, "rows_per_step": rows, "gather_ms_per_step": res["gather_ms_per_step"],
              "hbm_algorithmic": hbm, "hit_mix": mix})
    # the same algorithmic bytes over the WHOLE pipelined step of this rank: what the gather gets while the sampler's
    # kernels of the other batches in flight share the GPU with it (achieved / frac above are the gather launch alone)
    step_s = res["ms_per_step"] * 1e-3
    if step_s > 0:
        r["in_pipelined_step"] = {"hbm_algorithmic_GBps": alg / step_s / 1e9, "frac_of_hbm_peak": alg / step_s / 1e9 / hbm_peak,
                                  "note": "8D+8 bytes per gathered row over ms_per_step (gather + sampler chain of the batches in flight)"}
    return r


def run_workload(args, workload, rank, world, local, dist, steps, warmup, main_line):
    """everything for one workload; returns the result dict (rank 0) — the main line or an extra key"""
    import torch
    clocks = ClockSampler(local)
    if rank == 0 and main_line:
        clocks.start()
    c = setup_workload(args, workload, rank, world, local, dist)
    hbm_peak, peak_kind = measured_peak()
    kg = args.kg if args.kg > 0 else world  # Kc=1, Kg=N: one NVSwitch domain (legion_server.py:99-106)
    build_cache(c, kg, args.replicate_ratio)
    c.lab_host = c.lab.cpu().numpy()[c.my_train]
    peaks = link_peaks(c)
    # --- parity self-check against the CPU oracle (untimed; every rank; fails the run on mismatch) ---
    parity = None
    host_csr = None
    if args.parity_check:
        t0 = time.time()
        if c.topo_host:  # the CSR the sampler reads through UVA is the host copy already
            host_csr = Ctx()
            host_csr.indptr, host_csr.indices = c.h_ip.numpy(np.int64, (c.N + 1,)), c.h_ix.numpy(np.int32, (c.E,))
        else:
            host_csr = HostCSR(c.ip, c.ix, c.N, c.E, rank, world, dist)
        run_steps(c, 0, 2)  # warm
        torch.cuda.synchronize()
        parity = parity_selfcheck(c, host_csr, step=warmup + steps + 7)
        ok = torch.tensor([1 if parity["ok"] else 0], dtype=torch.int32, device=c.dev)
        tr = torch.tensor([parity["tier_rows_of_checked_batch"][k] for k in ("local", "peer", "host_or_backing")], dtype=torch.int64, device=c.dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            dist.all_reduce(tr)
        parity["ranks_checked"] = world
        parity["all_ranks_ok"] = bool(ok.item())
        parity["tier_rows_all_ranks"] = {"local": int(tr[0]), "peer": int(tr[1]), "host_or_backing": int(tr[2])}
        parity["what"] = ("one batch per rank vs oracle.Oracle.run_batch on the host CSR: node/edge counters (all 32 slots), ids, "
                          "labels, agg_src/agg_dst; gathered rows vs the feature function (all rows on the device, 2048 rows vs the numpy twin)")
        if rank == 0:
            log(f"[bench] parity self-check {parity['fields']} tiers={parity['tier_rows_all_ranks']} ({time.time() - t0:.1f}s incl. host CSR)")
        assert parity["all_ranks_ok"], f"parity self-check FAILED on some rank (rank {rank}: {parity['fields']})"
    main = measure_layout(c, steps, warmup, peaks, hbm_peak, full=True)
    # --- other layouts of the same table (extra keys): replicated (Kg=1) and hybrid ---
    layouts = {}
    if main_line and args.extras and world > 1 and args.kg <= 0 and not c.feat_host:
        for name, (kg2, rr) in {"replicated_kg1": (1, 0.0), f"hybrid_replicate_{args.extra_replicate_ratio:g}": (world, args.extra_replicate_ratio)}.items():
            build_cache(c, kg2, rr)
            r = measure_layout(c, steps, warmup, peaks, hbm_peak, full=False)
            layouts[name] = {k: r[k] for k in ("value", "ms_per_step", "cache", "tier_rows", "gather_ms_per_step", "features_bit_exact_selfcheck")}
            layouts[name]["hit_mix"] = {k: r["hit_mix"][k] for k in ("local", "peer", "host", "bound", "frac_of_mix_roofline")}
    clk = clocks.stop() if (rank == 0 and main_line) else None
    out = None
    if rank == 0:
        shape, dp = c.shape, c.dp
        layout_key = f"{shape['name']}/D{c.D}/kg{kg}" + ("/host" if c.feat_host else "")
        out = {
            "metric": "sampled+gathered seeds/sec", "value": main["value"], "unit": "seeds/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 ids / fp32 rows moved bit-exact",
            "data": "synthetic",
            "config": {"workload": workload_label(shape["name"]),
                       "num_nodes": c.N, "num_edges": c.E, "feature_dim": c.D, "fanout": c.fanout, "batch": c.B,
                       "scale": args.scale, "cache": main["cache"],
                       "feature_cache_ratio": args.cache_ratio, "topology": args.topo,
                       "topology_cache_ratio": args.topo_cache_ratio if c.topo_host else None,
                       "rng": "philox4x32-10", "position_map": ["dense u32[N]", "hashed L2-resident table"][dp.L.lg_sampler_dedup_layout(dp.sampler)],
                       "gather_mover": args.gather, "gather_fusion": args.fuse, "batches_in_flight": args.inflight,
                       "schedule": ["one stream", "gather overlaps next hop, joined per batch", "pipelined: every in-flight runner owns 2 INTERBATCH_CON buffer slots; the gather of batch k overlaps the sampling of the following batches"][args.overlap],
                       "l2": "working set (topology + features + per-batch output, tens of GB) exceeds the 126 MB L2; every step samples different seeds"},
            "e2e": {"value": main["e2e_value"], "unit": "seeds/s", "h2d_bytes_per_step": 2 * 4 * c.B, "d2h_bytes_per_step": 128,
                    "note": "every step: seed ids+labels copied from pinned host memory (H2D), lg_run_batch_host_async, both counter "
                            "arrays copied back to pinned host memory (D2H, what get_next reads); batches in flight as in the device-"
                            "resident run; features/COO stay in the CUDA-IPC buffers as in Legion's hand-off (no host round trip)"},
            "e2e_sync_per_step": {"value": main["e2e_sync_value"], "unit": "seeds/s",
                                  "note": "lg_run_batch_host: same copies, one batch at a time, stream synchronised after every step"},
            "gpu_launches": main["gpu_launches"],
            "gpu_launches_note": "kernels launched by liblegion_b200.so on rank 0 inside the timed region (counted by the library)",
            "roofline": roofline_of(c, main, hbm_peak, peak_kind, layout_key),
            "breakdown_ms": main["breakdown_ms"], "host_enqueue_ms_per_step": main["host_enqueue_ms_per_step"],
            "features_bit_exact_selfcheck": main["features_bit_exact_selfcheck"],
            "parity_selfcheck": parity,
            "tier_rows": main["tier_rows"],
            "link_peaks_measured": peaks,
            "clocks": clk,
        }
        if layouts:
            out["layouts"] = layouts
    # --- CPU baseline beside it (rank 0, N=1 only) ---
    if rank == 0 and world == 1 and main_line and not args.no_cpu_baseline:
        try:
            if host_csr is None and not c.topo_host:
                host_csr = HostCSR(c.ip, c.ix, c.N, c.E, rank, world, dist)
            out["cpu_baseline"] = cpu_arm(c.shape, host_csr.indptr, host_csr.indices, c.my_train, steps=args.cpu_steps,
                                          warmup=1)["cpu_baseline"]
        except MemoryError as ex:
            out["cpu_baseline"] = {"unavailable": repr(ex)[:200]}
    # release this workload's device memory (the next one may need it)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    c.dp.drop_feature_cache()
    for r, _, _ in c.runners:
        r.close()
    del c, host_csr
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from legion_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    capi.load()  # fails loudly if the CUDA library is missing

    out = run_workload(args, args.workload, rank, world, local, dist, args.steps, args.warmup, main_line=True)

    # --- extra keys at N=1: the products shape (BASELINE.json configs[1], round-1 headline) and the boundary itself ---
    if world == 1 and args.extras and args.workload != "products" and args.scale == 1.0 and args.cache_ratio >= 1.0 and args.topo == "hbm":
        try:
            sub = argparse.Namespace(**vars(args))
            sub.no_cpu_baseline = True
            p = run_workload(sub, "products", rank, world, local, dist, args.steps, args.warmup, main_line=False)
            out["products_configs1"] = {k: p[k] for k in ("value", "ms_per_step", "e2e", "roofline", "breakdown_ms", "parity_selfcheck",
                                                          "tier_rows", "gpu_launches")}
            out["products_configs1"]["config"] = {k: p["config"][k] for k in ("workload", "num_nodes", "num_edges", "feature_dim", "cache", "position_map")}
        except Exception as ex:  # noqa: BLE001  (never lose the main line to an extra key)
            out["products_configs1"] = {"unavailable": repr(ex)[:300]}
    if rank == 0 and world == 1 and args.server_e2e and args.extras:
        # the drop-in boundary itself: sampling_server binary -> shm/semaphores/CUDA IPC -> ipc_service (products shape:
        # the dataset must exist as files)
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "server_e2e.py"), "--epochs", "10"],
                               capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                j = json.loads(line[-1])
                out["e2e_server"] = {"value": j["seeds_per_s"], "unit": "seeds/s", "ms_per_batch": j["ms_per_batch"],
                                     "workload": workload_label("products"),
                                     "steps": j["train_steps_per_epoch"] * j["epochs"], "server_says": j["server_says"],
                                     "tier_telemetry": j.get("tier_telemetry"),
                                     "note": "C++ sampling_server binary (dataset files, meta_config, presampling, cost model, cache fill) -> "
                                             "simpleIPCshm + semaphores + CUDA-IPC buffers -> ipc_service.get_next/get_block_size/"
                                             "synchronize consumer, wall clock over the training steps of 10 epochs"}
            else:
                out["e2e_server"] = {"unavailable": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            out["e2e_server"] = {"unavailable": repr(ex)[:300]}
    if rank == 0 and world == 1 and args.ref_gpu and args.extras:
        # the reference's own CUDA kernels (oracle/_ref, compiled in place for sm_100a) on the same kind of batches:
        # the one baseline that is not a CPU
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ref_gpu_baseline.py")], capture_output=True,
                               text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out["ref_gpu_baseline"] = json.loads(line[-1]) if (r.returncode == 0 and line) else {"unavailable": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            out["ref_gpu_baseline"] = {"unavailable": repr(ex)[:300]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stderr.flush()
        print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port of the baseline BASELINE.json names): DGL-semantics sampler + index_select
# ----------------------------------------------------------------------------------------------
def host_memory_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 1e9
    except Exception:
        return 64.0


def cpu_arm(shape, indptr, indices, train, steps, warmup):
    """The CPU path timed on a bounded sample: `steps` batches.  The feature matrix is generated on the host cores by the
    oracle's twin of the feature function (oracle/synth_oracle.c); when [N x D] does not fit the host, the rows are
    folded onto the first N' vertices (index_select over ids % N') and the sample string says so."""
    from oracle import oracle as O
    B, fanout, D, N = shape["batch"], shape["fanout"], shape["D"], shape["N"]
    base = O.DGLBaseline(indptr, indices, fanout, B)
    base.use_all_cores()  # torchrun sets OMP_NUM_THREADS=1; the CPU arm gets every host core
    avail = host_memory_gb()
    n_feat = N
    if N * D * 4 / 1e9 > 0.8 * avail:
        n_feat = int(0.5 * avail * 1e9 / (D * 4))
    t0 = time.time()
    feat = np.empty((n_feat, D), np.float32)
    for r0 in range(0, n_feat, 1 << 24):
        O.synth_features(r0, min(1 << 24, n_feat - r0), D, SEED, out=feat[r0:r0 + (1 << 24)])
    log(f"[cpu arm] feature matrix {n_feat} x {D} on the host in {time.time() - t0:.1f}s ({base.threads()} threads)")
    out = np.empty((O.num_ids(B, fanout), D), np.float32)
    n_batches = max(1, (len(train) - 1) // B)
    times, rows = [], 0
    for s in range(warmup + steps):
        seeds = train[(s % n_batches) * B:(s % n_batches + 1) * B]
        t = time.perf_counter()
        n = base.sample(seeds, rng_seed=SEED + s)
        if n_feat != N:
            base.ids[:n] %= n_feat
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
                                       f"unique frontier) + index_select over {rows // max(steps, 1)} rows/batch, OpenMP {base.threads()} threads"
                                       + (f"; feature rows folded onto the first {n_feat} vertices (host RAM)" if n_feat != N else ""),
                             "gather_GBps": rows * (8 * D + 8) / tot / 1e9}}


def run_reference(args):
    """CPU arm only: the dataset is built by the oracle's host-side generator (no product code, no GPU)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    shape = shape_of(args)
    t0 = time.time()
    O.lib().lgo_set_num_threads(os.cpu_count() or 1)
    indptr, indices = O.synth_graph(shape["N"], shape["dmin"], shape["dmax"], SEED)
    train = train_split(shape, 1)[0]
    log(f"[reference] graph N={shape['N']} E={int(indptr[-1])} ready in {time.time() - t0:.1f}s")
    r = cpu_arm(shape, indptr, indices, train, steps=args.steps, warmup=args.warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {"impl": "reference", "metric": "sampled+gathered seeds/sec", "value": r["value"], "unit": "seeds/s",
           "n_gpus": max(args.gpus, world), "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 ids / fp32 rows", "data": "synthetic",
           "config": {"workload": workload_label(shape["name"]),
                      "num_nodes": shape["N"], "num_edges": int(indptr[-1]), "feature_dim": shape["D"],
                      "fanout": shape["fanout"], "batch": shape["batch"], "scale": args.scale,
                      "implementation": "CPU: DGL-semantics NeighborSampler + index_select (oracle port, OpenMP); dataset from oracle/synth_oracle.c"},
           "cpu_baseline": r["cpu_baseline"],
           "e2e": {"value": r["value"], "unit": "seeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ukunion", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale N and E of the named shape (1.0 = paper size)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--presample", type=int, default=20, help="presampling batches used for the hotness ranking")
    ap.add_argument("--gather", default="auto", choices=["auto", "ldg", "tma"])
    ap.add_argument("--fuse", type=int, default=2, choices=[0, 1, 2],
                    help="gather launches per batch: 0 one per lookup op, 1 seeds ride with hop 1, 2 single gather")
    ap.add_argument("--kg", type=int, default=0,
                    help="GPUs sharing one partitioned cache; 0 = all N GPUs (Kc=1, Kg=N: one NVSwitch domain, the reference's "
                         "cache_agg_mode on such a box), 1 = replicated")
    ap.add_argument("--replicate-ratio", type=float, default=0.0,
                    help="with Kg > 1: fraction of the rows (hottest first) stored on every GPU instead of partitioned")
    ap.add_argument("--extra-replicate-ratio", type=float, default=0.2, help="replicate ratio of the hybrid layout reported as an extra key")
    ap.add_argument("--cache-ratio", type=float, default=1.0,
                    help="fraction of the feature rows cached in HBM across the clique; < 1 puts the backing matrix in "
                         "pinned host memory and misses are read over PCIe (UVA)")
    ap.add_argument("--topo", default="hbm", choices=["hbm", "host"],
                    help="where the full CSR lives: replicated in HBM, or pinned host memory read through UVA")
    ap.add_argument("--topo-cache-ratio", type=float, default=0.0,
                    help="with --topo host: fraction of the vertices whose adjacency lists are cached in HBM")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight per GPU (own scratch + stream each)")
    ap.add_argument("--overlap", type=int, default=2, choices=[0, 1, 2],
                    help="0 one stream; 1 gathers overlap the next hop; 2 pipelined across the two batch slots")
    ap.add_argument("--no-identity", dest="identity", action="store_false",
                    help="Kg=1 with every row cached: keep the reference's hotness-rank placement + directory instead of storing "
                         "the rows at row index = vertex id")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", dest="parity_check", action="store_false",
                    help="skip the untimed per-rank comparison of one batch with the CPU oracle")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="only the main line: no other layouts (N>1), no products shape / server / reference-kernel legs (N=1)")
    ap.add_argument("--no-server-e2e", dest="server_e2e", action="store_false",
                    help="skip the e2e_server leg (scripts/server_e2e.py: the sampling_server binary feeding an ipc_service "
                         "consumer over shm/semaphores/CUDA IPC; N=1, products shape, ~30 s)")
    ap.add_argument("--no-ref-gpu", dest="ref_gpu", action="store_false",
                    help="skip the ref_gpu_baseline leg (the reference's own kernels from oracle/_ref, device time)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
