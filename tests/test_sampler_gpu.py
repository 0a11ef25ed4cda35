"""-m gpu: the CUDA sampler (through the C ABI) against the CPU oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["dense", "hash"])
def dedup_layout(request, monkeypatch):
    """every test runs with both layouts of the position map (csrc/sampler.cu: DENSE word per vertex / HASHED table)"""
    monkeypatch.setenv("LG_DEDUP", request.param)
    return request.param
torch = pytest.importorskip("torch")

from conftest import make_sets, small_graph  # noqa: E402
from gpu_util import Rig, assert_batch_equal  # noqa: E402
from legion_b200 import capi, synth  # noqa: E402


def _feat(n, d=8):
    return synth.features(0, n, d, 5)


@pytest.mark.parametrize("rng", [capi.RNG_MINSTD, capi.RNG_PHILOX])
@pytest.mark.parametrize("fanout,batch", [([5, 3], 64), ([25, 10], 256), ([15, 10, 5], 100), ([3], 17), ([40, 2], 33)])
def test_batch_matches_oracle(oracle, rng, fanout, batch):
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    feat = _feat(N)
    ids, labels = make_sets(N)
    rig = Rig(indptr, indices, feat, fanout, batch)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    orc = oracle.Oracle(indptr, indices, fanout, batch)
    for counter in (0, 3):
        p = rig.dp.params(d_ids, d_lab, batch, counter, rng_kind=rng, seed=0xABCDEF0123, batch_id=counter, stream_id=2)
        rig.dp.run_once(p, buf)
        torch.cuda.synchronize()
        want = orc.run_batch(ids, labels, batch, counter, rng_kind=rng, seed=0xABCDEF0123, batch_id=counter, stream_id=2)
        assert_batch_equal(buf.to_host(len(fanout)), want, len(fanout), feat)
    assert rig.dp.status() == 0


def test_per_op_counter_trace(oracle):
    """counters after every op equal the reference state machine (engine/operator_impl.cu:57-89)"""
    indptr, indices = small_graph()
    N = len(indptr) - 1
    ids, labels = make_sets(N)
    fanout, B = [6, 4], 50
    rig = Rig(indptr, indices, _feat(N), fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    L, dp = rig.dp.L, rig.dp
    st = dp._stream()
    want = oracle.Oracle(indptr, indices, fanout, B).run_batch(ids, labels, B, 1, seed=5, batch_id=1, per_hop=True)
    trace = {op: (nc, ec) for op, nc, ec in want["trace"]}

    def snap():
        torch.cuda.synchronize()
        return buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()

    capi.check(L.lg_batch_generate(dp.sampler, st, d_ids.data_ptr(), d_lab.data_ptr(), len(ids), B, 1, C.byref(buf.c)))
    nc, ec = snap()
    assert np.array_equal(nc, trace[0][0]) and np.array_equal(ec, trace[0][1])
    capi.check(L.lg_feature_cache_lookup(dp.sampler, st, C.byref(dp.cache), 1, 0, C.byref(buf.c), None))
    nc, ec = snap()
    assert np.array_equal(nc, trace[1][0]) and np.array_equal(ec, trace[1][1])
    for hop in (1, 2):
        capi.check(L.lg_random_sample(dp.sampler, st, C.byref(dp.topo), hop, capi.RNG_PHILOX, 5, 1, 0, C.byref(buf.c), None))
        nc, ec = snap()
        assert np.array_equal(nc, trace[3 * hop][0]) and np.array_equal(ec, trace[3 * hop][1]), hop
        capi.check(L.lg_feature_cache_lookup(dp.sampler, st, C.byref(dp.cache), 3 * hop + 1, 0, C.byref(buf.c), None))
        nc, ec = snap()
        assert np.array_equal(nc, trace[3 * hop + 1][0]) and np.array_equal(ec, trace[3 * hop + 1][1]), hop


@pytest.mark.parametrize("tail", ["exact", "reference"])
def test_edge_cases(oracle, tail):
    # isolated vertices (deg 0), deg < fanout, duplicate seeds, clipped tail (both strides), empty batch
    rng = np.random.default_rng(11)
    N = 400
    deg = rng.integers(0, 9, N)
    deg[::5] = 0
    indptr = np.zeros(N + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, N, int(indptr[-1])).astype(np.int32)
    feat = _feat(N)
    fanout, B = [4, 3], 40
    ids = rng.permutation(N).astype(np.int32)[:130]
    ids[5] = ids[3]  # duplicate seed: first position wins in oracle and kernel
    labels = (ids % 3).astype(np.int32)
    rig = Rig(indptr, indices, feat, fanout, B)
    rig.dp.set_tail_mode(capi.TAIL_REFERENCE if tail == "reference" else capi.TAIL_EXACT)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    orc = oracle.Oracle(indptr, indices, fanout, B)
    for counter in (0, 1, 2, 3, 4):  # 3 = clipped tail (130 - 120 = 10 seeds), 4 = empty
        buf.features.fill_(777.0)
        p = rig.dp.params(d_ids, d_lab, B, counter, seed=1, batch_id=counter)
        rig.dp.run_once(p, buf)
        torch.cuda.synchronize()
        want = orc.run_batch(ids, labels, B, counter, seed=1, batch_id=counter, tail_exact=(tail == "exact"))
        got = buf.to_host(2)
        assert_batch_equal(got, want, 2, feat)
        if counter == 3:  # exact: the true tail of the set; reference: clipped size x counter (operator_impl.cu:159-162)
            first = 120 if tail == "exact" else 30
            assert np.array_equal(got["ids"][:10], ids[first:first + 10])
    assert want["total_nodes"] == 0 and want["total_edges"] == 0


@pytest.mark.parametrize("tail", ["exact", "reference"])
def test_tail_padding_minus_one(oracle, tail):
    """the clipped tail batch in both stride modes (engine/operator_impl.cu:40-44,159-162)"""
    indptr, indices = small_graph(500, 8.0, 60)
    N = len(indptr) - 1
    ids, labels = make_sets(N, 0.2)  # 100 ids
    rig = Rig(indptr, indices, _feat(N), [3, 2], 64)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    orc = oracle.Oracle(indptr, indices, [3, 2], 64)
    rig.dp.set_tail_mode(capi.TAIL_REFERENCE if tail == "reference" else capi.TAIL_EXACT)
    p = rig.dp.params(d_ids, d_lab, 64, 1, seed=3, batch_id=1)  # clipped: size = 36; stride 36 (reference quirk) or 64
    rig.dp.run_once(p, buf)
    torch.cuda.synchronize()
    want = orc.run_batch(ids, labels, 64, 1, seed=3, batch_id=1, tail_exact=(tail == "exact"))
    assert want["nc"][9] == 36
    assert np.array_equal(want["ids"][:36], ids[64:100] if tail == "exact" else ids[36:72])
    assert_batch_equal(buf.to_host(2), want, 2)


@pytest.mark.parametrize("release", ["fill", "scatter"])
def test_position_map_is_released_between_batches(oracle, monkeypatch, release):
    """every batch must leave the position map all-absent (ClearPosMap, engine/operator_impl.cu:542-548), also when a
    batch is abandoned after batch_generate or mid-way; both release paths of the dense layout: the streaming fill
    (maps of a few MB) and the O(batch) scatter over the batch's vertices (LG_PM_FILL_MB=0: what larger maps use)"""
    monkeypatch.setenv("LG_PM_FILL_MB", "16" if release == "fill" else "0")
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    ids, labels = make_sets(N)
    fanout, B = [10, 5], 128
    rig = Rig(indptr, indices, _feat(N), fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    L, dp = rig.dp.L, rig.dp
    st = dp._stream()
    orc = oracle.Oracle(indptr, indices, fanout, B)
    # abandoned batches: only op 0, then op 0 + hop 1 (no lg_io_complete)
    capi.check(L.lg_batch_generate(dp.sampler, st, d_ids.data_ptr(), d_lab.data_ptr(), len(ids), B, 5, C.byref(buf.c)))
    capi.check(L.lg_batch_generate(dp.sampler, st, d_ids.data_ptr(), d_lab.data_ptr(), len(ids), B, 6, C.byref(buf.c)))
    capi.check(L.lg_random_sample(dp.sampler, st, C.byref(dp.topo), 1, capi.RNG_PHILOX, 9, 6, 0, C.byref(buf.c), None))
    for counter in (2, 2, 7):  # same batch twice: identical output only if nothing leaked from the first run
        p = dp.params(d_ids, d_lab, B, counter, seed=9, batch_id=counter)
        dp.run_once(p, buf)
        torch.cuda.synchronize()
        want = orc.run_batch(ids, labels, B, counter, seed=9, batch_id=counter)
        assert_batch_equal(buf.to_host(2), want, 2)
    capi.check(L.lg_sampler_reset(dp.sampler, st))
    p = dp.params(d_ids, d_lab, B, 1, seed=9, batch_id=1)
    dp.run_once(p, buf)
    torch.cuda.synchronize()
    assert_batch_equal(buf.to_host(2), orc.run_batch(ids, labels, B, 1, seed=9, batch_id=1), 2)
    assert dp.status() == 0


@pytest.mark.parametrize("host_topology", [False, True])
def test_topology_cache_tiers_do_not_change_the_sample(oracle, host_topology):
    """hot rows from an HBM shard, the rest from the full CSR (HBM or host UVA): identical output
    (engine/operator_impl.cu:224-243)"""
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    feat = _feat(N)
    ids, labels = make_sets(N)
    fanout, B = [8, 4], 100
    rig = Rig(indptr, indices, feat, fanout, B, host_topology=host_topology, host_features=host_topology)
    hot = torch.from_numpy(np.bincount(indices, minlength=N).astype(np.int64)).to(rig.dev)
    order, _ = rig.dp.rank_hotness(hot)
    rig.dp.build_topology_cache(order, cap=700)
    rig.dp.build_feature_cache(order, cap=500)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    p = rig.dp.params(d_ids, d_lab, B, 1, seed=4, batch_id=1)
    rig.dp.run_once(p, buf, tier=True)
    torch.cuda.synchronize()
    want = oracle.Oracle(indptr, indices, fanout, B).run_batch(ids, labels, B, 1, seed=4, batch_id=1)
    assert_batch_equal(buf.to_host(2), want, 2, feat)
    tiers = rig.dp.tier_rows.cpu().numpy()
    assert tiers.sum() == want["total_nodes"] and tiers[0] > 0 and tiers[2] > 0 and tiers[1] == 0


def test_presampling_hotness(oracle):
    """pre_sample edge counts + HotnessMeasure + max ids (operator_impl.cu:358; cache_impl.cuh:190-198; cache.cu:59-61)"""
    indptr, indices = small_graph()
    N = len(indptr) - 1
    ids, labels = make_sets(N)
    fanout, B = [5, 3], 64
    rig = Rig(indptr, indices, _feat(N), fanout, B, host_topology=True)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch(feature_rows=1)
    eh = torch.zeros(N, dtype=torch.int64, device=rig.dev)
    nh = torch.zeros(N, dtype=torch.int64, device=rig.dev)
    mx = torch.zeros(1, dtype=torch.int32, device=rig.dev)
    w_eh, w_nh, w_mx = np.zeros(N, np.uint64), np.zeros(N, np.uint64), 0
    orc = oracle.Oracle(indptr, indices, fanout, B)
    for it in range(6):
        p = rig.dp.params(d_ids, d_lab, B, it, seed=2, batch_id=it)
        rig.dp.run_presc(p, buf, eh, nh, mx)
        o = orc.run_batch(ids, labels, B, it, seed=2, batch_id=it, edge_hot=w_eh, node_hot=w_nh)
        w_mx = max(w_mx, o["max_ids"])
    torch.cuda.synchronize()
    assert np.array_equal(eh.cpu().numpy().astype(np.uint64), w_eh)
    assert np.array_equal(nh.cpu().numpy().astype(np.uint64), w_nh)
    assert int(mx.item()) == w_mx
    # ranking parity (ties by ascending id)
    order, sh = rig.dp.rank_hotness(nh)
    w_order, w_sh = oracle.hotness_rank(w_nh)
    assert np.array_equal(order.cpu().numpy(), w_order) and np.array_equal(sh.cpu().numpy().astype(np.uint64), w_sh)


def test_e2e_host_call_matches_device_call(oracle):
    indptr, indices = small_graph()
    N = len(indptr) - 1
    feat = _feat(N)
    ids, labels = make_sets(N)
    fanout, B = [5, 3], 64
    rig = Rig(indptr, indices, feat, fanout, B)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    h_ids = torch.from_numpy(ids[128:192].copy()).pin_memory()
    h_lab = torch.from_numpy(labels[128:192].copy()).pin_memory()
    nc, ec = np.zeros(16, np.int32), np.zeros(16, np.int32)
    p = rig.dp.params(d_ids, d_lab, B, 2, seed=6, batch_id=2)
    rig.dp.run_once_host(p, h_ids.numpy(), h_lab.numpy(), buf, nc, ec)
    want = oracle.Oracle(indptr, indices, fanout, B).run_batch(ids, labels, B, 2, seed=6, batch_id=2)
    assert np.array_equal(nc, want["nc"]) and np.array_equal(ec, want["ec"])
    assert_batch_equal(buf.to_host(2), want, 2, feat)


def test_distribution_fanout_and_uniformity(oracle):
    """fan-out bound and per-draw uniformity (chi-square) of the with-replacement pick; inclusion frequency
    against the DGL-semantics sampler's c/deg (without replacement) where the two are comparable"""
    N, degv, c = 64, 40, 10
    indptr = np.arange(N + 1, dtype=np.int64) * degv
    indices = np.tile(np.arange(degv, dtype=np.int32), N)  # vertex v's neighbours are 0..39
    feat = _feat(N)
    seeds = np.arange(N, dtype=np.int32)
    rig = Rig(indptr, indices, feat, [c], N)
    d_ids, d_lab = rig.sets(seeds, seeds)
    buf = rig.dp.alloc_batch()
    counts = np.zeros(degv, np.int64)
    n_batches = 200
    for b in range(n_batches):
        p = rig.dp.params(d_ids, d_lab, N, 0, seed=0x5EED, batch_id=b)
        rig.dp.run_once(p, buf, gather=False)
        torch.cuda.synchronize()
        h = buf.to_host(1)
        assert h["total_edges"] == N * c  # min(deg, fanout) draws per seed
        per_dst = np.bincount(h["agg_dst"], minlength=N)
        assert (per_dst == c).all()
        counts += np.bincount(h["ids"][h["agg_src"]], minlength=degv)
    total = counts.sum()
    chi2 = ((counts - total / degv) ** 2 / (total / degv)).sum()
    assert chi2 < 80.0, chi2  # 39 dof: p(chi2 > 80) ~ 1e-4
    # DGL-like baseline: inclusion probability c/deg per neighbour; ours 1-(1-1/deg)^c per seed
    base = oracle.DGLBaseline(indptr, indices, [c], N)
    inc_dgl = np.zeros(degv)
    for b in range(50):
        base.sample(seeds, rng_seed=b + 1)
        e = int(base.le[1])
        assert e == N * c
        inc_dgl += np.bincount(base.ids[base.src[:e]], minlength=degv)
    inc_dgl /= 50 * N
    assert abs(inc_dgl.mean() - c / degv) < 1e-9 and np.abs(inc_dgl - c / degv).max() < 0.05


@pytest.mark.parametrize("fuse", [0, 1, 2])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_stream_schedules_give_identical_batches(oracle, mode, fuse):
    """one stream / gather overlapping the next hop / pipelined over two buffer slots: same bits"""
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    feat = _feat(N, 100)
    ids, labels = make_sets(N)
    fanout, B = [10, 5], 128
    rig = Rig(indptr, indices, feat, fanout, B)
    rig.dp.set_overlap(mode)
    rig.dp.set_gather_fusion(fuse)
    d_ids, d_lab = rig.sets(ids, labels)
    bufs = [rig.dp.alloc_batch(), rig.dp.alloc_batch()]
    orc = oracle.Oracle(indptr, indices, fanout, B)
    n_batches = 8
    for it in range(n_batches):
        rig.dp.run_once(rig.dp.params(d_ids, d_lab, B, it, seed=17, batch_id=it), bufs[it % 2])
        if it >= 1:  # consume the previous batch while this one is in flight (what the trainer does)
            prev = bufs[(it - 1) % 2]
            rig.dp.batch_wait(prev)
            torch.cuda.current_stream().synchronize()
            want = orc.run_batch(ids, labels, B, it - 1, seed=17, batch_id=it - 1)
            assert_batch_equal(prev.to_host(2), want, 2, feat)
    rig.dp.batch_wait(bufs[(n_batches - 1) % 2])
    torch.cuda.synchronize()
    want = orc.run_batch(ids, labels, B, n_batches - 1, seed=17, batch_id=n_batches - 1)
    assert_batch_equal(bufs[(n_batches - 1) % 2].to_host(2), want, 2, feat)


def test_two_runners_in_flight_on_one_gpu(oracle):
    """two sampler handles on two streams sharing the storage descriptors (bench --inflight 2)"""
    from legion_b200.runner import DataPath
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    feat = _feat(N, 100)
    ids, labels = make_sets(N)
    fanout, B = [10, 5], 128
    rig = Rig(indptr, indices, feat, fanout, B)
    hot = torch.from_numpy(np.bincount(indices, minlength=N).astype(np.int64)).to(rig.dev)
    order, _ = rig.dp.rank_hotness(hot)
    rig.dp.build_feature_cache(order, cap=1000)
    d2 = DataPath(0, fanout, B, N, 100)
    d2.share_storage_from(rig.dp)
    runners = [(rig.dp, rig.dp.alloc_batch(), torch.cuda.current_stream()), (d2, d2.alloc_batch(), torch.cuda.Stream())]
    for r, _, _ in runners:
        r.set_overlap(2)
        r.set_gather_fusion(2)
    d_ids, d_lab = rig.sets(ids, labels)
    torch.cuda.synchronize()
    orc = oracle.Oracle(indptr, indices, fanout, B)
    for rnd in range(3):
        for k, (r, buf, st) in enumerate(runners):
            with torch.cuda.stream(st):
                r.run_once(r.params(d_ids, d_lab, B, 2 * rnd + k, seed=3, batch_id=2 * rnd + k), buf)
        for k, (r, buf, st) in enumerate(runners):
            with torch.cuda.stream(st):
                r.batch_wait(buf)
            st.synchronize()
            want = orc.run_batch(ids, labels, B, 2 * rnd + k, seed=3, batch_id=2 * rnd + k)
            assert_batch_equal(buf.to_host(2), want, 2, feat)
    d2.close()


def test_async_host_fed_batches(oracle):
    """lg_run_batch_host_async: several host-fed batches in flight, counters land in pinned memory"""
    indptr, indices = small_graph()
    N = len(indptr) - 1
    feat = _feat(N, 100)
    ids, labels = make_sets(N)
    fanout, B = [5, 3], 64
    rig = Rig(indptr, indices, feat, fanout, B)
    rig.dp.set_overlap(2)
    rig.dp.set_gather_fusion(2)
    d_ids, d_lab = rig.sets(ids, labels)
    bufs = [rig.dp.alloc_batch(), rig.dp.alloc_batch()]
    h_ids = torch.from_numpy(ids.copy()).pin_memory()
    h_lab = torch.from_numpy(labels.copy()).pin_memory()
    cnt = torch.zeros((4, 32), dtype=torch.int32).pin_memory()
    cn = cnt.numpy()
    orc = oracle.Oracle(indptr, indices, fanout, B)
    for it in range(4):
        p = rig.dp.params(d_ids, d_lab, B, it, seed=6, batch_id=it)
        rig.dp.run_once_host_async(p, h_ids.numpy()[it * B:(it + 1) * B], h_lab.numpy()[it * B:(it + 1) * B], bufs[it % 2],
                                   cn[it, :16], cn[it, 16:])
        if it % 2 == 1:  # consume the two slots before they are reused
            torch.cuda.synchronize()
            for k in (it - 1, it):
                want = orc.run_batch(ids, labels, B, k, seed=6, batch_id=k)
                assert np.array_equal(cn[k, :16], want["nc"]) and np.array_equal(cn[k, 16:], want["ec"])
                assert_batch_equal(bufs[k % 2].to_host(2), want, 2, feat)


@pytest.mark.parametrize("rng", [capi.RNG_MINSTD, capi.RNG_PHILOX])
@pytest.mark.parametrize("fanout,batch", [([25, 10], 256), ([15, 10, 5], 100), ([3], 17)])
def test_chain_kernel_matches_oracle(oracle, monkeypatch, dedup_layout, rng, fanout, batch):
    """LG_CHAIN=1 (opt-in): the dense layout's whole sampler chain as one persistent kernel with grid barriers
    (csrc/sampler.cu chain_kernel); same batches, bit for bit, also with two runners in flight on one GPU"""
    if dedup_layout != "dense":
        pytest.skip("chain_kernel exists for the dense position map only")
    monkeypatch.setenv("LG_CHAIN", "1")
    indptr, indices = small_graph(3000, 14.0, 400)
    N = len(indptr) - 1
    feat = _feat(N)
    ids, labels = make_sets(N)
    rigs = [Rig(indptr, indices, feat, fanout, batch) for _ in range(2)]
    for r in rigs:
        r.dp.set_gather_fusion(2)  # one gather per batch: the chain has no hop boundaries to gather at
    streams = [torch.cuda.Stream() for _ in rigs]
    bufs = [r.dp.alloc_batch() for r in rigs]
    sets = [r.sets(ids, labels) for r in rigs]
    orc = oracle.Oracle(indptr, indices, fanout, batch)
    for rnd in range(3):
        for k, rig in enumerate(rigs):  # both runners enqueue before anybody synchronises
            counter = 2 * rnd + k
            with torch.cuda.stream(streams[k]):
                p = rig.dp.params(sets[k][0], sets[k][1], batch, counter, rng_kind=rng, seed=77, batch_id=counter, stream_id=1)
                rig.dp.run_once(p, bufs[k])
        torch.cuda.synchronize()
        for k, rig in enumerate(rigs):
            counter = 2 * rnd + k
            want = orc.run_batch(ids, labels, batch, counter, rng_kind=rng, seed=77, batch_id=counter, stream_id=1)
            assert_batch_equal(bufs[k].to_host(len(fanout)), want, len(fanout), feat)
    assert all(r.dp.status() == 0 for r in rigs)


@pytest.mark.parametrize("fanout,batch", [([6, 4], 50), ([15, 10, 5], 100), ([7], 33)])
def test_lazy_relabel_op_by_op(oracle, fanout, batch):
    """lg_sampler_set_lazy_relabel(1) (what the server runs): hop h's construct_graph is finished by hop h+1's sample
    kernel, the last sampling op also releases the position map, the lookup of all hops is one launch at the end
    (LEGION_FUSE_GATHERS).  Counters after every sampling op and the final batch equal the oracle's; several batches on one
    handle prove the release."""
    indptr, indices = small_graph()
    N = len(indptr) - 1
    feat = _feat(N)
    ids, labels = make_sets(N)
    rig = Rig(indptr, indices, feat, fanout, batch)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    L, dp = rig.dp.L, rig.dp
    st = dp._stream()
    H = len(fanout)
    capi.check(L.lg_sampler_set_lazy_relabel(dp.sampler, 1))
    orc = oracle.Oracle(indptr, indices, fanout, batch)
    for counter in (0, 1, 4):
        want = orc.run_batch(ids, labels, batch, counter, seed=9, batch_id=counter, per_hop=True)
        trace = {op: (nc, ec) for op, nc, ec in want["trace"]}
        capi.check(L.lg_batch_generate(dp.sampler, st, d_ids.data_ptr(), d_lab.data_ptr(), len(ids), batch, counter, C.byref(buf.c)))
        for hop in range(1, H + 1):
            capi.check(L.lg_random_sample(dp.sampler, st, C.byref(dp.topo), hop, capi.RNG_PHILOX, 9, counter, 0, C.byref(buf.c), None))
            torch.cuda.synchronize()
            nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
            # the lookup ops in between are skipped, so only the cumulative per-hop totals are compared here
            assert np.array_equal(nc[9:], trace[3 * hop][0][9:]) and np.array_equal(ec[9:], trace[3 * hop][1][9:]), hop
        capi.check(L.lg_feature_cache_lookup_range(dp.sampler, st, C.byref(dp.cache), 3 * H + 1, 0, 0, C.byref(buf.c), None))
        capi.check(L.lg_io_complete(dp.sampler, st, capi.TRAINMODE, C.byref(buf.c), None, None))
        torch.cuda.synchronize()
        assert_batch_equal(buf.to_host(H), want, H, feat)
    assert dp.status() == 0


@pytest.mark.parametrize("chunk", [1, 4])
def test_dynamic_gather_tiles(oracle, monkeypatch, chunk):
    """LG_GATHER_DYNAMIC: the gather's CTAs claim chunks of tiles from a counter (prefetched one chunk ahead) instead of a
    fixed share; more tiles than first chunks, several launches on one handle (the last CTA re-arms the counter)"""
    monkeypatch.setenv("LG_GATHER_DYNAMIC", str(chunk))
    indptr, indices = small_graph(200000, 10.0, 300, seed=11)
    N = len(indptr) - 1
    feat = _feat(N, 128)  # 512-byte rows: 10 single-warp CTAs per SM, 1480 first chunks
    ids, labels = make_sets(N, frac=0.2)
    fanout, B = [25, 10], 4096
    rig = Rig(indptr, indices, feat, fanout, B)
    rig.dp.set_gather_variant(capi.GATHER_TMA)
    d_ids, d_lab = rig.sets(ids, labels)
    buf = rig.dp.alloc_batch()
    orc = oracle.Oracle(indptr, indices, fanout, B)
    for counter in range(3):
        p = rig.dp.params(d_ids, d_lab, B, counter, seed=99, batch_id=counter)
        rig.dp.run_once(p, buf)
        torch.cuda.synchronize()
        want = orc.run_batch(ids, labels, B, counter, seed=99, batch_id=counter)
        if counter == 0:
            assert want["total_nodes"] > 1.5 * 148 * 10 * 8 * chunk  # later chunks come from the counter
        assert_batch_equal(buf.to_host(2), want, 2, feat)
    assert rig.dp.status() == 0
