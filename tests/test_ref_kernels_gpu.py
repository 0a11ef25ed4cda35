"""-m gpu: the REFERENCE's own kernels (oracle/_ref/libref_ops.so: /root/reference sources compiled in place for
sm_100a, see oracle/ref_ops.cu) run beside liblegion_b200.so and beside the oracle.  This is what pins parity on
the reference itself: deterministic kernels are compared bit-for-bit, the atomics-ordered sampler as multisets,
and the oracle's hop-2 replay is validated from the reference's own hop-1 order."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import make_sets, small_graph  # noqa: E402
from gpu_util import Rig  # noqa: E402
from legion_b200 import capi, synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref_ops.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_ops.so not built (needs /root/reference at build time)")
    return C.CDLL(REF)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to("cuda:0") if dtype is None else t.to("cuda:0").to(dtype)


def P(t):
    return C.c_void_p(t.data_ptr())


class RefState:
    """the reference's MemoryPool scratch (engine/server.cu:221-234) as torch tensors"""

    def __init__(self, N, num_ids, B):
        z = lambda n, dt=torch.int32: torch.zeros(n, dtype=dt, device="cuda:0")  # noqa: E731
        self.ids, self.labels = z(num_ids), z(B)
        self.agg_src_ids, self.agg_dst_ids = z(num_ids), z(num_ids)
        self.agg_src_off, self.agg_dst_off = z(num_ids), z(num_ids)
        self.accessed = z(N // 32 + 1)
        self.position_map = z(N)
        self.nc, self.ec = z(16), z(16)
        self.part_idx = torch.full((num_ids,), -2, dtype=torch.int8, device="cuda:0")  # every lookup misses
        self.part_off = z(num_ids)


def _ref_run_hops(ref, st, rs, d_ip, d_ix, fanout, d_ids, d_lab, B, counter):
    tab_ip = torch.tensor([d_ip.data_ptr()], dtype=torch.int64, device="cuda:0")  # P+1 = 1 slot: the full CSR
    tab_ix = torch.tensor([d_ix.data_ptr()], dtype=torch.int64, device="cuda:0")
    assert ref.ref_batch_generate(st, P(rs.ids), P(rs.labels), B, counter, P(d_ids), P(d_lab), d_ids.numel(),
                                  P(rs.position_map), P(rs.accessed), rs.position_map.numel(), P(rs.nc), P(rs.ec),
                                  len(fanout)) == 0
    snaps = []
    torch.cuda.synchronize()
    snaps.append((rs.nc.cpu().numpy().copy(), rs.ec.cpu().numpy().copy()))
    assert ref.ref_counter_update(st, P(rs.nc), P(rs.ec), 1) == 0  # op 1 (CacheLookup) snapshot
    for h, c in enumerate(fanout, 1):
        assert ref.ref_random_sample(st, P(rs.ids), 3 * h, P(tab_ip), P(tab_ix), P(rs.part_idx), P(rs.part_off), c, 0,
                                     P(rs.agg_src_ids), P(rs.agg_dst_ids), P(rs.agg_src_off), P(rs.agg_dst_off),
                                     P(rs.accessed), P(rs.position_map), P(rs.nc), P(rs.ec)) == 0
        torch.cuda.synchronize()
        snaps.append((rs.nc.cpu().numpy().copy(), rs.ec.cpu().numpy().copy()))
        assert ref.ref_counter_update(st, P(rs.nc), P(rs.ec), 3 * h + 1) == 0  # op 3h+1 snapshot
    return snaps


def _edge_multiset(src, dst):
    return np.sort(src.astype(np.int64) * (1 << 32) + dst.astype(np.int64))


def test_reference_sampler_vs_ours_and_oracle(ref, oracle):
    indptr, indices = small_graph(4000, 16.0, 500)
    N = len(indptr) - 1
    ids, labels = make_sets(N)
    fanout, B, counter = [25, 10], 300, 2
    rig = Rig(indptr, indices, synth.features(0, N, 8, 1), fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    st = rig.dp._stream()
    rs = RefState(N, rig.dp.num_ids, B)
    snaps = _ref_run_hops(ref, st, rs, rig.d_ip, rig.d_ix, fanout, d_ids, d_lab, B, counter)

    # ours with the reference's RNG stream (minstd) + the oracle
    buf = rig.dp.alloc_batch(feature_rows=1)
    p = rig.dp.params(d_ids, d_lab, B, counter, rng_kind=capi.RNG_MINSTD)
    rig.dp.run_once(p, buf, gather=False)
    torch.cuda.synchronize()
    mine = buf.to_host(2)
    orc = oracle.Oracle(indptr, indices, fanout, B)
    want = orc.run_batch(ids, labels, B, counter, rng_kind=oracle.RNG_MINSTD, per_hop=True)

    # op 0 is deterministic: bit-exact against the reference kernels
    r_nc0, r_ec0 = snaps[0]
    tr = {op: (nc, ec) for op, nc, ec in want["trace"]}
    assert np.array_equal(r_nc0, tr[0][0]) and np.array_equal(r_ec0, tr[0][1])
    r_ids = rs.ids.cpu().numpy()
    assert np.array_equal(r_ids[:B], mine["ids"][:B]) and np.array_equal(rs.labels.cpu().numpy()[:B], mine["labels"])

    # hop 1: frontier order is fixed (the seeds), so the reference's edges/new nodes equal ours as multisets/sets
    r_nc1, r_ec1 = snaps[1]
    assert np.array_equal(r_nc1, tr[3][0]) and np.array_equal(r_ec1, tr[3][1])  # all 16+16 counter slots
    e1, n1 = int(r_ec1[10]), int(r_nc1[10])
    r_src, r_dst = rs.agg_src_ids.cpu().numpy(), rs.agg_dst_ids.cpu().numpy()
    m_gsrc, m_gdst = mine["ids"][mine["agg_src"][:e1]], mine["ids"][mine["agg_dst"][:e1]]
    assert np.array_equal(_edge_multiset(r_src[:e1], r_dst[:e1]), _edge_multiset(m_gsrc, m_gdst))
    assert np.array_equal(np.sort(r_ids[B:n1]), np.sort(mine["ids"][B:n1]))
    # the reference's construct_graph output is consistent with its own ids (local index semantics)
    r_so, r_do = rs.agg_src_off.cpu().numpy(), rs.agg_dst_off.cpu().numpy()
    assert np.array_equal(r_ids[r_so[:e1]], r_src[:e1]) and np.array_equal(r_ids[r_do[:e1]], r_dst[:e1])

    # hop 2: replay the ORACLE from the reference's own hop-1 order -> must reproduce the reference's hop 2
    r_nc2, r_ec2 = snaps[2]
    L = oracle.lib()
    s2 = L.lgo_state_create(N)
    L.lgo_state_load(s2, r_ids, n1)
    o_ids, o_gsrc, o_gdst = r_ids.copy(), r_src.copy(), r_dst.copy()
    o_ids[n1:] = -7
    o_gsrc[e1:] = -7
    o_asrc, o_adst = np.full_like(r_src, -7), np.full_like(r_src, -7)
    nc, ec = r_nc1.copy(), r_ec1.copy()
    L.lgo_counter_update(nc, ec, 4, 0, 0)
    L.lgo_random_sample(s2, indptr, indices, 2, fanout[1], oracle.RNG_MINSTD, 0, 0, 0, o_ids, o_gsrc, o_gdst, o_asrc,
                        o_adst, nc, ec, None)
    L.lgo_state_destroy(s2)
    assert np.array_equal(nc, r_nc2) and np.array_equal(ec, r_ec2)
    e2, n2 = int(r_ec2[11]), int(r_nc2[11])
    assert np.array_equal(_edge_multiset(o_gsrc[e1:e2], o_gdst[e1:e2]), _edge_multiset(r_src[e1:e2], r_dst[e1:e2]))
    assert np.array_equal(np.sort(o_ids[n1:n2]), np.sort(r_ids[n1:n2]))
    assert np.array_equal(r_ids[r_so[e1:e2]], r_src[e1:e2]) and np.array_equal(r_ids[r_do[e1:e2]], r_dst[e1:e2])
    # and our own hop 2 has the same size (sum of min(deg, fanout) over the hop-1 edge sources, order-independent)
    assert mine["ec"][11] == r_ec2[11]


def test_reference_presample_hotness_vs_ours(ref, oracle):
    indptr, indices = small_graph(3000, 12.0, 300)
    N = len(indptr) - 1
    ids, labels = make_sets(N)
    fanout, B = [6, 4], 128
    rig = Rig(indptr, indices, synth.features(0, N, 8, 1), fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    st = rig.dp._stream()
    rs = RefState(N, rig.dp.num_ids, B)
    r_eh = torch.zeros(N, dtype=torch.int64, device="cuda:0")
    r_nh = torch.zeros(N, dtype=torch.int64, device="cuda:0")
    assert ref.ref_batch_generate(st, P(rs.ids), P(rs.labels), B, 1, P(d_ids), P(d_lab), d_ids.numel(), P(rs.position_map),
                                  P(rs.accessed), N, P(rs.nc), P(rs.ec), 2) == 0
    assert ref.ref_pre_sample(st, P(rs.ids), 3, P(rig.d_ip), P(rig.d_ix), fanout[0], P(rs.agg_src_ids), P(rs.agg_dst_ids),
                              P(rs.agg_src_off), P(rs.agg_dst_off), P(rs.accessed), P(rs.position_map), P(rs.nc), P(rs.ec),
                              P(r_eh)) == 0
    assert ref.ref_hotness_measure(st, P(rs.ids), P(rs.nc), P(r_nh)) == 0
    torch.cuda.synchronize()
    # ours: one hop of presampling with the minstd stream
    rig1 = Rig(indptr, indices, synth.features(0, N, 8, 1), fanout[:1], B)
    d_ids1, d_lab1 = rig1.sets(ids, labels)
    buf = rig1.dp.alloc_batch(feature_rows=1)
    eh = torch.zeros(N, dtype=torch.int64, device="cuda:0")
    nh = torch.zeros(N, dtype=torch.int64, device="cuda:0")
    mx = torch.zeros(1, dtype=torch.int32, device="cuda:0")
    rig1.dp.run_presc(rig1.dp.params(d_ids1, d_lab1, B, 1, rng_kind=capi.RNG_MINSTD), buf, eh, nh, mx)
    torch.cuda.synchronize()
    assert torch.equal(eh, r_eh) and torch.equal(nh, r_nh)
    assert int(mx.item()) == int(rs.nc[7].item())


def test_reference_placement_fill_and_gather_vs_ours(ref, oracle):
    indptr, indices = small_graph(2500, 9.0, 120)
    N, dim, kg, cap = len(indptr) - 1, 100, 4, 300
    feat = synth.features(0, N, dim, 8)
    rig = Rig(indptr, indices, feat, [2], 8)
    st = rig.dp._stream()
    rng = np.random.default_rng(3)
    hot = dev(rng.integers(0, 40, N).astype(np.int64))
    order, _ = rig.dp.rank_hotness(hot)
    L = rig.dp.L
    # --- InitPair (cache/cache_impl.cuh:104-109) + bcht insert/find  ==  our dense directory ---
    sizes = (C.c_int32 * 3)()
    ref.ref_pair_sizes(sizes)
    assert list(sizes) == [8, 8, 8]
    pairs = torch.empty(2 * cap * kg, dtype=torch.int32, device="cuda:0")
    assert ref.ref_init_pair(st, P(pairs), P(order), cap, kg, kg) == 0
    d = torch.empty(N, dtype=torch.int32, device="cuda:0")
    capi.check(L.lg_fill_i32(st, d.data_ptr(), -2, N))
    capi.check(L.lg_place_features(st, order.data_ptr(), cap, kg, N, d.data_ptr()))
    torch.cuda.synchronize()
    pk = pairs.cpu().numpy().reshape(-1, 2)
    dn = d.cpu().numpy()
    assert np.array_equal(dn[pk[:, 0]], pk[:, 1]) and (dn >= 0).sum() == cap * kg
    keys = dev(rng.integers(0, N, 5000).astype(np.int32))
    found = torch.empty(5000, dtype=torch.int32, device="cuda:0")
    assert ref.ref_bcht_build_and_find(st, P(pairs), cap * kg, C.c_int64(2 * cap * kg), P(keys), 5000, P(found)) == 0
    assert torch.equal(found, d[keys.long()])  # find => value or CACHEMISS_FLAG (-2)
    # --- InitIndexPair / InitOffsetPair (cache_impl.cuh:89-101) == our packed topology directory ---
    ipair = torch.empty(2 * cap * kg, dtype=torch.int32, device="cuda:0")  # {int32 key, char value} padded to 8 bytes
    opair = torch.empty(2 * cap * kg, dtype=torch.int32, device="cuda:0")
    assert ref.ref_init_topo_pairs(st, P(ipair), P(opair), P(order), cap, kg, kg, 0) == 0
    td = torch.empty(N, dtype=torch.int32, device="cuda:0")
    capi.check(L.lg_fill_i32(st, td.data_ptr(), -2, N))
    capi.check(L.lg_place_topology(st, order.data_ptr(), cap, kg, 0, N, td.data_ptr()))
    torch.cuda.synchronize()
    ik, ok_ = ipair.cpu().numpy().reshape(-1, 2), opair.cpu().numpy().reshape(-1, 2)
    part = ik[:, 1].astype(np.int32).view(np.int8)[::4].astype(np.int32)  # low byte = char part_id
    tdn = td.cpu().numpy()
    assert np.array_equal(tdn[ik[:, 0]], part * cap + ok_[:, 1]) and np.array_equal(ik[:, 0], ok_[:, 0])
    # --- FeatFillUp (cache_impl.cuh:183-188) == lg_fill_feature_shard, bit-exact; then the gather ---
    shards_ref, shards_mine = [], []
    for j in range(kg):
        a = torch.empty((cap, dim), dtype=torch.float32, device="cuda:0")
        b = torch.empty((cap, dim), dtype=torch.float32, device="cuda:0")
        assert ref.ref_feat_fill_up(st, cap, dim, P(a), C.c_void_p(rig.dp._backing), P(order), kg, j) == 0
        capi.check(L.lg_fill_feature_shard(st, order.data_ptr(), cap, kg, j, dim, N, rig.dp._backing, b.data_ptr()))
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
        shards_ref.append(a)
        shards_mine.append(b)
    # multiGPU_feat_cache_lookup (cache_impl.cuh:239-272) vs both movers of ours: rows [off, off+cnt) of ids
    n_ids, off, cnt = 6000, 700, 5000
    ids = rng.integers(0, N, n_ids).astype(np.int32)
    ids[off + 3] = -1
    d_ids = dev(ids)
    nc = torch.zeros(16, dtype=torch.int32, device="cuda:0")
    nc[2], nc[3] = off, cnt  # op_id % 3 == 1 reads slots 2,3
    cache_index = d[d_ids[off:off + cnt].clamp(min=0).long()].contiguous()
    cache_index[3] = -2
    ptrs = torch.tensor([s.data_ptr() for s in shards_ref], dtype=torch.int64, device="cuda:0")
    out_ref = torch.full((n_ids, dim), 777.0, dtype=torch.float32, device="cuda:0")
    assert ref.ref_feat_cache_lookup(st, C.c_void_p(rig.dp._backing), P(ptrs), dim, P(d_ids), P(cache_index), cap, P(nc),
                                     P(out_ref), N, 4) == 0
    torch.cuda.synchronize()
    cache = capi.FeatureCache()
    cache.n_parts, cache.shard_rows, cache.dim, cache.num_nodes = kg, cap, dim, N
    for j in range(kg):
        cache.shard[j] = shards_mine[j].data_ptr()
    cache.backing, cache.directory = rig.dp._backing, d.data_ptr()
    for variant in (capi.GATHER_LDG, capi.GATHER_TMA):
        out = torch.full((n_ids, dim), 777.0, dtype=torch.float32, device="cuda:0")
        capi.check(L.lg_gather_rows(st, C.byref(cache), d_ids[off:].data_ptr(), cnt, out[off:].data_ptr(), 0, variant, None))
        torch.cuda.synchronize()
        assert torch.equal(out.view(torch.int32), out_ref.view(torch.int32)), variant


def test_reference_tail_batch_stride(ref, oracle):
    """The clipped tail batch: the reference's own batch_generate (engine/operator_impl.cu:27-55 launched as :159-165)
    equals ours in LG_TAIL_REFERENCE mode and the oracle's literal mode; LG_TAIL_EXACT (default) serves the true tail."""
    indptr, indices = small_graph(3000, 10.0, 200)
    N = len(indptr) - 1
    ids, labels = make_sets(N, 0.1)  # 300 ids
    fanout, B, counter = [5, 3], 128, 2  # 300 - 256 = 44 seeds left
    rig = Rig(indptr, indices, synth.features(0, N, 8, 1), fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    st = rig.dp._stream()
    rs = RefState(N, rig.dp.num_ids, B)
    tab = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    assert ref.ref_batch_generate(st, P(rs.ids), P(rs.labels), B, counter, P(d_ids), P(d_lab), d_ids.numel(),
                                  P(rs.position_map), P(rs.accessed), N, P(rs.nc), P(rs.ec), len(fanout)) == 0
    torch.cuda.synchronize()
    r_nc, r_ids, r_lab = rs.nc.cpu().numpy(), rs.ids.cpu().numpy(), rs.labels.cpu().numpy()
    assert r_nc[9] == 44
    buf = rig.dp.alloc_batch(feature_rows=1)
    orc = oracle.Oracle(indptr, indices, fanout, B)
    for mode, exact in ((capi.TAIL_REFERENCE, False), (capi.TAIL_EXACT, True)):
        rig.dp.set_tail_mode(mode)
        rig.dp.run_once(rig.dp.params(d_ids, d_lab, B, counter, seed=1, batch_id=counter), buf, gather=False)
        torch.cuda.synchronize()
        mine = buf.to_host(2)
        want = orc.run_batch(ids, labels, B, counter, seed=1, batch_id=counter, tail_exact=exact)
        assert np.array_equal(mine["ids"][:44], want["ids"][:44]) and np.array_equal(mine["labels"], want["labels"][:44])
        if not exact:  # the reference kernel itself: seeds from 44 * 2 = 88
            assert np.array_equal(mine["ids"][:44], r_ids[:44]) and np.array_equal(mine["labels"], r_lab[:44])
            assert np.array_equal(r_ids[:44], ids[88:132])
        else:
            assert np.array_equal(mine["ids"][:44], ids[256:300])
    del tab


COST_CASES = [
    # N, avg deg, dim, cache bytes per GPU, Kg, train_step, pcie counters (topo transactions)
    (6000, 10.0, 16, 200_000, 1, 30, (0, 0)),            # the reference's normal case: PCM counters are 0
    (6000, 10.0, 16, 60_000, 2, 30, (0, 0)),
    (6000, 10.0, 100, 300_000, 4, 12, (40_000, 2_000)),   # with topology transactions
    (6000, 10.0, 128, 123_457, 8, 7, (5_000_000, 0)),
    (3000, 6.0, 16, 40_000_000, 1, 30, (1000, 0)),        # degenerate: cache > dataset (cache/cache.cu:528-549)
    (3000, 6.0, 16, 40_000_000, 4, 30, (0, 0)),
]


@pytest.mark.parametrize("case", COST_CASES)
def test_cost_model_equals_reference(ref, oracle, case):
    """UnifiedCache::CandidateSelection + CostModel (cache/cache.cu:360-551), run through the reference class itself
    (oracle/ref_ops.cu: ref_cost_model), against lgo_cost_model (oracle) and lg_cost_model (the library)."""
    N, avg, dim, cache_bytes, kg, train_step, counters = case
    indptr, indices = small_graph(N, avg, 300)
    N = len(indptr) - 1
    rng = np.random.default_rng(N + kg)
    # distinct aggregate hotness per vertex, so that the ranking does not depend on how ties are broken
    # (thrust's sort is unstable, cache/cache.cu:415; ours breaks ties by vertex id)
    base_n = rng.permutation(N).astype(np.uint64) * np.uint64(kg) * np.uint64(3)
    base_e = rng.permutation(N).astype(np.uint64) * np.uint64(kg) * np.uint64(5)
    node_hot = [(base_n // np.uint64(kg)) + np.uint64(j == 0) * (base_n % np.uint64(kg)) for j in range(kg)]
    edge_hot = [(base_e // np.uint64(kg)) + np.uint64(j == 0) * (base_e % np.uint64(kg)) for j in range(kg)]
    node_hot = [np.ascontiguousarray(x, np.uint64) for x in node_hot]
    edge_hot = [np.ascontiguousarray(x, np.uint64) for x in edge_hot]
    assert np.array_equal(sum(node_hot), base_n) and np.array_equal(sum(edge_hot), base_e)
    max_ids = np.asarray(rng.integers(1000, 5000, kg), np.int32)
    arr = lambda xs: (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])  # noqa: E731
    ncap, ecap = C.c_int32(), C.c_int32()
    QF, QT = np.zeros(N, np.int32), np.zeros(N, np.int32)
    AF, AT = np.zeros(N, np.uint64), np.zeros(N, np.uint64)
    ref.ref_cost_model.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    rc = ref.ref_cost_model(N, dim, cache_bytes, kg, train_step, arr(node_hot), arr(edge_hot), max_ids.ctypes.data,
                            indptr.ctypes.data, counters[0], counters[1], C.addressof(ncap), C.addressof(ecap),
                            QF.ctypes.data, QT.ctypes.data, AF.ctypes.data, AT.ctypes.data)
    assert rc == 0
    # CandidateSelection: the reference's ranking equals ours (hotness descending; no ties here)
    o_order_n, o_sorted_n = oracle.hotness_rank(base_n)
    o_order_e, o_sorted_e = oracle.hotness_rank(base_e)
    assert np.array_equal(QF, o_order_n) and np.array_equal(AF, o_sorted_n)
    assert np.array_equal(QT, o_order_e) and np.array_equal(AT, o_sorted_e)
    # CostModel inputs as the reference forms them (cache/cache.cu:459-463)
    topo_trans = counters[0] + counters[1]
    feat_trans = sum((int(m) * train_step * dim * 4) // 64 for m in max_ids)
    want = (ncap.value, ecap.value)
    got_o = oracle.cost_model(o_sorted_n, o_sorted_e, o_order_e, indptr, dim, cache_bytes, kg, topo_trans, feat_trans)
    assert got_o[:2] == want, (got_o, want)
    L = capi.load()
    a, b, al = C.c_int32(), C.c_int32(), C.c_double()
    assert L.lg_cost_model(o_sorted_n.ctypes.data, o_sorted_e.ctypes.data, o_order_e.ctypes.data, indptr.ctypes.data, N, dim,
                           cache_bytes, kg, topo_trans, feat_trans, C.byref(a), C.byref(b), C.byref(al)) == 0
    assert (a.value, b.value) == want
