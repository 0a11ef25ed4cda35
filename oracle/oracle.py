"""ctypes/numpy front-end of the CPU oracle (oracle/legion_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by legion_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblegion_oracle.so")
_lib = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("legion_oracle.c", "synth_oracle.c")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(_LIB_PATH) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, _LIB_PATH])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    L.lgo_minstd_x.restype = C.c_uint32
    L.lgo_minstd_x.argtypes = [C.c_uint64]
    L.lgo_pick_minstd.restype = C.c_int32
    L.lgo_pick_minstd.argtypes = [C.c_uint64, C.c_int32]
    L.lgo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    L.lgo_pick_philox.restype = C.c_int32
    L.lgo_pick_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32]
    L.lgo_counter_update.argtypes = [i32p, i32p, C.c_int32, C.c_int32, C.c_int32]
    L.lgo_state_create.restype = C.c_void_p
    L.lgo_state_create.argtypes = [C.c_int64]
    L.lgo_state_destroy.argtypes = [C.c_void_p]
    L.lgo_state_load.argtypes = [C.c_void_p, i32p, C.c_int32]
    L.lgo_batch_generate.argtypes = [C.c_void_p, i32p, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     i32p, i32p, i32p, i32p]
    L.lgo_random_sample.argtypes = [C.c_void_p, i64p, i32p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64,
                                    C.c_uint32, C.c_uint32, i32p, i32p, i32p, i32p, i32p, i32p, i32p,
                                    C.c_void_p]
    L.lgo_hotness_measure.restype = C.c_int32
    L.lgo_hotness_measure.argtypes = [i32p, i32p, u64p]
    L.lgo_hotness_rank.argtypes = [u64p, C.c_int64, i32p, C.c_void_p]
    L.lgo_place_features.argtypes = [i32p, C.c_int32, C.c_int32, C.c_int64, i32p]
    L.lgo_place_topology.argtypes = [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, i32p]
    L.lgo_fill_feature_shard.argtypes = [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, f32p, f32p]
    L.lgo_place_features_hybrid.argtypes = [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, i32p]
    L.lgo_fill_feature_shard_hybrid.argtypes = [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, f32p, f32p]
    L.lgo_fill_topo_shard.argtypes = [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, i64p, i32p, i64p,
                                      C.c_void_p]
    L.lgo_feature_lookup.argtypes = [i32p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32,
                                     f32p, C.c_int64, C.c_int32, f32p]
    L.lgo_cost_model.argtypes = [u64p, u64p, i32p, i64p, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                 C.c_uint64, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_double)]
    L.lgo_cost_model_saturating.argtypes = L.lgo_cost_model.argtypes
    L.lgo_coordinate.argtypes = [i32p, i32p, i32p, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, i32p,
                                 C.POINTER(C.c_int32)]
    L.lgo_mode_of.argtypes = [C.c_int32, i32p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.lgo_dgl_sample.restype = C.c_int64
    L.lgo_dgl_sample.argtypes = [i64p, i32p, i32p, C.c_int32, i32p, C.c_int32, C.c_uint64, i32p, i32p, i32p,
                                 i32p, i64p, i64p, i64p, i32p]
    L.lgo_index_select.argtypes = [f32p, C.c_int32, i32p, C.c_int64, f32p]
    L.lgo_block_csc.argtypes = [i32p, i32p, C.c_int64, C.c_int32, i32p, i32p, i32p]
    L.lgo_synth_indptr.restype = C.c_int64
    L.lgo_synth_indptr.argtypes = [C.c_int64, C.c_double, C.c_int32, C.c_uint64, i64p]
    L.lgo_synth_indices.argtypes = [C.c_int64, i64p, C.c_uint64, i32p]
    L.lgo_synth_features.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_uint64, f32p]
    L.lgo_synth_feature_rows.argtypes = [i32p, C.c_int64, C.c_int32, C.c_uint64, f32p]
    L.lgo_set_tail_exact.argtypes = [C.c_int32]
    L.lgo_num_threads.restype = C.c_int32
    L.lgo_set_num_threads.argtypes = [C.c_int32]
    _lib = L
    return L


RNG_MINSTD, RNG_PHILOX = 0, 1


def num_ids(batch, fanout):
    """engine/server.cu:187-199"""
    tot, per = batch, batch
    for f in fanout:
        per *= f
        tot += per
    return tot


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().lgo_philox4x32_10(c, k, o)
    return [int(x) for x in o]


class Oracle:
    """Replays GPURunner::RunOnce (engine/server.cu:302-332) for one GPU on the CPU."""

    def __init__(self, indptr, indices, fanout, max_batch):
        self.L = lib()
        self.indptr = np.ascontiguousarray(indptr, np.int64)
        self.indices = np.ascontiguousarray(indices, np.int32)
        self.N = len(self.indptr) - 1
        self.fanout = list(fanout)
        self.hops = len(self.fanout)
        self.num_ids = num_ids(max_batch, self.fanout)
        self.st = self.L.lgo_state_create(self.N)

    def __del__(self):
        try:
            self.L.lgo_state_destroy(self.st)
        except Exception:
            pass

    def run_batch(self, all_ids, all_labels, batch_size, counter, rng_kind=RNG_PHILOX, seed=0, batch_id=0,
                  stream_id=0, edge_hot=None, node_hot=None, per_hop=False, tail_exact=True):
        """tail_exact: the clipped tail batch starts at batch_size*counter (LG_TAIL_EXACT, the C ABI's default);
        False = the reference's literal stride (engine/operator_impl.cu:159-162)"""
        L = self.L
        L.lgo_set_tail_exact(1 if tail_exact else 0)
        all_ids = np.ascontiguousarray(all_ids, np.int32)
        all_labels = np.ascontiguousarray(all_labels, np.int32)
        n = self.num_ids
        ids = np.full(n, -7, np.int32)
        labels = np.full(max(batch_size, 1), -7, np.int32)
        gid_src = np.full(n, -7, np.int32)
        gid_dst = np.full(n, -7, np.int32)
        agg_src = np.full(n, -7, np.int32)
        agg_dst = np.full(n, -7, np.int32)
        nc = np.zeros(16, np.int32)
        ec = np.zeros(16, np.int32)
        trace = []
        L.lgo_batch_generate(self.st, all_ids, all_labels, len(all_ids), batch_size, counter, self.hops, ids,
                             labels, nc, ec)
        trace.append((0, nc.copy(), ec.copy()))
        L.lgo_counter_update(nc, ec, 1, 0, 0)  # op1 CacheLookup snapshot (engine/operator_impl.cu:515)
        trace.append((1, nc.copy(), ec.copy()))
        eh = edge_hot.ctypes.data_as(C.c_void_p) if edge_hot is not None else None
        for h in range(1, self.hops + 1):
            L.lgo_random_sample(self.st, self.indptr, self.indices, h, self.fanout[h - 1], rng_kind, seed,
                                batch_id, stream_id, ids, gid_src, gid_dst, agg_src, agg_dst, nc, ec, eh)
            trace.append((3 * h, nc.copy(), ec.copy()))
            L.lgo_counter_update(nc, ec, 3 * h + 1, 0, 0)
            trace.append((3 * h + 1, nc.copy(), ec.copy()))
        max_ids = 0
        if node_hot is not None:
            max_ids = L.lgo_hotness_measure(ids, nc, node_hot)
        H = self.hops
        out = dict(ids=ids, labels=labels, agg_src=agg_src, agg_dst=agg_dst, gid_src=gid_src, gid_dst=gid_dst,
                   nc=nc, ec=ec, total_nodes=int(nc[9 + H]), total_edges=int(ec[9 + H]), max_ids=max_ids)
        if per_hop:
            out["trace"] = trace
        return out


def feature_lookup(ids, node_off, cnt, directory, shards, cap, backing, dim, dst):
    L = lib()
    N = backing.shape[0]
    arr = (C.c_void_p * max(len(shards), 1))(*[s.ctypes.data for s in shards])
    d = directory.ctypes.data_as(C.c_void_p) if directory is not None else None
    L.lgo_feature_lookup(np.ascontiguousarray(ids, np.int32), node_off, cnt, d, arr, cap,
                         backing.reshape(-1), N, dim, dst.reshape(-1))
    return dst


def block_csc(src, dst, num_dst):
    """COO -> CSC of one block, in-edges in COO order (lgo_block_csc)"""
    src, dst = np.ascontiguousarray(src, np.int32), np.ascontiguousarray(dst, np.int32)
    e = len(src)
    indptr, indices, eids = np.zeros(num_dst + 1, np.int32), np.zeros(max(e, 1), np.int32), np.zeros(max(e, 1), np.int32)
    lib().lgo_block_csc(src if e else np.zeros(1, np.int32), dst if e else np.zeros(1, np.int32), e, num_dst, indptr, indices, eids)
    return indptr, indices[:e], eids[:e]


def hotness_rank(hot):
    hot = np.ascontiguousarray(hot, np.uint64)
    order = np.empty(len(hot), np.int32)
    sh = np.empty(len(hot), np.uint64)
    lib().lgo_hotness_rank(hot, len(hot), order, sh.ctypes.data_as(C.c_void_p))
    return order, sh


def place_features(order, cap, kg, N):
    d = np.empty(N, np.int32)
    lib().lgo_place_features(np.ascontiguousarray(order, np.int32), cap, kg, N, d)
    return d


def place_features_hybrid(order, cap, kg, rep, j, N):
    """lg_place_features_hybrid: the rep hottest ranks replicated on every part, the rest interleaved below them"""
    d = np.empty(N, np.int32)
    lib().lgo_place_features_hybrid(np.ascontiguousarray(order, np.int32), cap, kg, rep, j, N, d)
    return d


def fill_feature_shard_hybrid(order, cap, kg, rep, j, backing):
    N, dim = backing.shape
    sh = np.empty((cap, dim), np.float32)
    lib().lgo_fill_feature_shard_hybrid(np.ascontiguousarray(order, np.int32), cap, kg, rep, j, dim, N,
                                        np.ascontiguousarray(backing, np.float32).reshape(-1), sh.reshape(-1))
    return sh


def place_topology(order, cap, kg, ki, N):
    d = np.empty(N, np.int32)
    lib().lgo_place_topology(np.ascontiguousarray(order, np.int32), cap, kg, ki, N, d)
    return d


def fill_feature_shard(order, cap, kg, j, backing):
    N, dim = backing.shape
    sh = np.empty((cap, dim), np.float32)
    lib().lgo_fill_feature_shard(np.ascontiguousarray(order, np.int32), cap, kg, j, dim, N,
                                 backing.reshape(-1), sh.reshape(-1))
    return sh


def fill_topo_shard(order, cap, kg, j, indptr, indices):
    N = len(indptr) - 1
    order = np.ascontiguousarray(order, np.int32)
    sip = np.empty(cap + 1, np.int64)
    lib().lgo_fill_topo_shard(order, cap, kg, j, N, indptr, indices, sip, None)
    sidx = np.empty(max(int(sip[-1]), 1), np.int32)
    lib().lgo_fill_topo_shard(order, cap, kg, j, N, indptr, indices, sip, sidx.ctypes.data_as(C.c_void_p))
    return sip, sidx[: int(sip[-1])]


def cost_model(sorted_node_hot, sorted_edge_hot, topo_order, indptr, dim, cache_bytes, kg, topo_trans, feat_trans,
               saturating=False):
    nc_, ec_, al = C.c_int32(), C.c_int32(), C.c_double()
    n = len(topo_order)
    fn = lib().lgo_cost_model_saturating if saturating else lib().lgo_cost_model
    fn(np.ascontiguousarray(sorted_node_hot, np.uint64),
                         np.ascontiguousarray(sorted_edge_hot, np.uint64),
                         np.ascontiguousarray(topo_order, np.int32), np.ascontiguousarray(indptr, np.int64),
                         n, dim, cache_bytes, kg, topo_trans, feat_trans, C.byref(nc_), C.byref(ec_), C.byref(al))
    return nc_.value, ec_.value, al.value


def coordinate(train_num, valid_num, test_num, raw_batch, epoch):
    P = len(train_num)
    steps = np.zeros(3, np.int32)
    vb = np.zeros(P, np.int32)
    tb = np.zeros(P, np.int32)
    ms = C.c_int32()
    lib().lgo_coordinate(np.asarray(train_num, np.int32), np.asarray(valid_num, np.int32),
                         np.asarray(test_num, np.int32), P, raw_batch, epoch, steps, vb, tb, C.byref(ms))
    return steps, vb, tb, ms.value


def mode_of(gb, steps, epoch):
    m, l = C.c_int32(), C.c_int32()
    lib().lgo_mode_of(gb, np.asarray(steps, np.int32), epoch, C.byref(m), C.byref(l))
    return m.value, l.value


class DGLBaseline:
    """DGL NeighborSampler semantics + index_select on the host cores."""

    def __init__(self, indptr, indices, fanout, max_batch):
        self.L = lib()
        self.indptr = np.ascontiguousarray(indptr, np.int64)
        self.indices = np.ascontiguousarray(indices, np.int32)
        self.N = len(self.indptr) - 1
        self.fanout = np.asarray(fanout, np.int32)
        n = num_ids(max_batch, list(fanout))
        self.pos = np.full(self.N, -1, np.int32)
        self.ids = np.empty(n, np.int32)
        self.src = np.empty(n, np.int32)
        self.dst = np.empty(n, np.int32)
        self.eoff = np.empty(n + 1, np.int64)
        self.gid = np.empty(n, np.int32)
        self.ln = np.zeros(len(fanout) + 1, np.int64)
        self.le = np.zeros(len(fanout) + 1, np.int64)

    def sample(self, seeds, rng_seed=1):
        seeds = np.ascontiguousarray(seeds, np.int32)
        n = self.L.lgo_dgl_sample(self.indptr, self.indices, seeds, len(seeds), self.fanout, len(self.fanout),
                                  rng_seed, self.pos, self.ids, self.src, self.dst, self.ln, self.le, self.eoff,
                                  self.gid)
        return int(n)

    def gather(self, feat, n, out):
        self.L.lgo_index_select(feat.reshape(-1), feat.shape[1], self.ids, n, out.reshape(-1))
        return out

    def threads(self):
        return int(self.L.lgo_num_threads())

    def use_all_cores(self):
        self.L.lgo_set_num_threads(os.cpu_count() or 1)
        return self.threads()


# ---- host-side twin of the synthetic dataset generator (oracle/synth_oracle.c) ----
def synth_graph(n, dmin, dmax, seed):
    """CSR of the synthetic graph, generated on the host cores (OpenMP); equals legion_b200.synth.graph bit for bit"""
    L = lib()
    indptr = np.empty(n + 1, np.int64)
    e = int(L.lgo_synth_indptr(n, float(dmin), int(dmax), int(seed), indptr))
    indices = np.empty(max(e, 1), np.int32)
    L.lgo_synth_indices(n, indptr, int(seed), indices)
    return indptr, indices[:e]


def synth_features(row0, rows, dim, seed, out=None):
    if out is None:
        out = np.empty((rows, dim), np.float32)
    lib().lgo_synth_features(int(row0), int(rows), int(dim), int(seed), out.reshape(-1))
    return out


def synth_feature_rows(ids, dim, seed):
    ids = np.ascontiguousarray(ids, np.int32)
    out = np.empty((len(ids), dim), np.float32)
    lib().lgo_synth_feature_rows(ids, len(ids), int(dim), int(seed), out.reshape(-1))
    return out
