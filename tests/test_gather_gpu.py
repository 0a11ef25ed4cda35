"""-m gpu: gather kernels (LDG and TMA movers) bit-exact against the oracle / numpy index_select."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from gpu_util import Rig  # noqa: E402
from conftest import small_graph  # noqa: E402
from legion_b200 import capi, synth  # noqa: E402


def _gather(rig, ids, variant, tier=None):
    d_ids = torch.from_numpy(ids).to(rig.dev)
    out = torch.full((len(ids), rig.D), 777.0, dtype=torch.float32, device=rig.dev)
    st = rig.dp._stream()
    capi.check(rig.dp.L.lg_gather_rows(st, C.byref(rig.dp.cache), d_ids.data_ptr(), len(ids), out.data_ptr(), 0, variant,
                                       tier.data_ptr() if tier is not None else None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.fixture(params=["1", "0"], ids=["gather4", "rowwise"])
def gather4(request, monkeypatch):
    """the TMA mover with and without tile::gather4 tensor copies (LG_GATHER4, read per launch)"""
    monkeypatch.setenv("LG_GATHER4", request.param)
    return request.param


@pytest.mark.parametrize("variant", [capi.GATHER_LDG, capi.GATHER_TMA])
@pytest.mark.parametrize("dim", [100, 128, 256, 4, 3, 130])
@pytest.mark.parametrize("host_features", [False, True])
def test_gather_bit_exact(oracle, variant, dim, host_features, gather4):
    indptr, indices = small_graph(1500, 6.0, 50)
    N = len(indptr) - 1
    feat = synth.features(0, N, dim, 77)
    rig = Rig(indptr, indices, feat, [2], 8, host_features=host_features)
    rng = np.random.default_rng(dim)
    hot = torch.from_numpy(rng.integers(0, 1000, N).astype(np.int64)).to(rig.dev)
    order, _ = rig.dp.rank_hotness(hot)
    directory = rig.dp.build_feature_cache(order, cap=600)
    for n in (1, 31, 32, 33, 1000, 4097):
        ids = rng.integers(0, N, n).astype(np.int32)
        if n > 40:
            ids[5] = -1  # skipped row (cache_impl.cuh:263-264): destination untouched
            ids[37] = -1
        tier = torch.zeros(3, dtype=torch.int64, device=rig.dev)
        got = _gather(rig, ids, variant, tier)
        want = np.full((n, dim), 777.0, np.float32)
        d = directory.cpu().numpy()
        shard = rig.dp.feat_shard.tensor(torch.float32, (600, dim)).cpu().numpy()
        oracle.feature_lookup(ids, 0, n, d, [shard], 600, feat, dim, want)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (n, dim)
        ok = ids >= 0
        assert np.array_equal(got[ok].view(np.uint32), feat[ids[ok]].view(np.uint32))
        t = tier.cpu().numpy()
        assert t[0] == (d[ids[ok]] >= 0).sum() and t[2] == (d[ids[ok]] < 0).sum() and t[1] == 0


@pytest.mark.parametrize("variant", [capi.GATHER_LDG, capi.GATHER_TMA])
def test_gather_without_cache_reads_backing(variant):
    indptr, indices = small_graph(800, 6.0, 50)
    N = len(indptr) - 1
    feat = synth.features(0, N, 100, 3)
    rig = Rig(indptr, indices, feat, [2], 8)
    ids = np.random.default_rng(0).integers(0, N, 5000).astype(np.int32)
    got = _gather(rig, ids, variant)
    assert np.array_equal(got.view(np.uint32), feat[ids].view(np.uint32))


def test_cache_build_matches_oracle(oracle):
    indptr, indices = small_graph(2500, 9.0, 120)
    N = len(indptr) - 1
    dim = 100
    feat = synth.features(0, N, dim, 8)
    rig = Rig(indptr, indices, feat, [2], 8, host_topology=True, host_features=True)
    rng = np.random.default_rng(4)
    hot_np = rng.integers(0, 50, N).astype(np.uint64)  # many ties
    order, sh = rig.dp.rank_hotness(torch.from_numpy(hot_np.astype(np.int64)).to(rig.dev))
    w_order, w_sh = oracle.hotness_rank(hot_np)
    assert np.array_equal(order.cpu().numpy(), w_order)
    L, st = rig.dp.L, rig.dp._stream()
    for kg, cap in [(1, 300), (4, 200), (8, 313), (8, 400)]:  # last: cap*kg > N -> ranks past N skipped
        d = torch.empty(N, dtype=torch.int32, device=rig.dev)
        capi.check(L.lg_fill_i32(st, d.data_ptr(), -2, N))
        capi.check(L.lg_place_features(st, order.data_ptr(), cap, kg, N, d.data_ptr()))
        assert np.array_equal(d.cpu().numpy(), oracle.place_features(w_order, cap, kg, N))
        capi.check(L.lg_fill_i32(st, d.data_ptr(), -2, N))
        capi.check(L.lg_place_topology(st, order.data_ptr(), cap, kg, 0, N, d.data_ptr()))
        assert np.array_equal(d.cpu().numpy(), oracle.place_topology(w_order, cap, kg, 0, N))
        for j in (0, kg - 1):
            sh_t = torch.empty((cap, dim), dtype=torch.float32, device=rig.dev)
            capi.check(L.lg_fill_feature_shard(st, order.data_ptr(), cap, kg, j, dim, N, rig.dp._backing, sh_t.data_ptr()))
            assert np.array_equal(sh_t.cpu().numpy().view(np.uint32),
                                  oracle.fill_feature_shard(w_order, cap, kg, j, feat).view(np.uint32))
            sip = torch.empty(cap + 1, dtype=torch.int64, device=rig.dev)
            capi.check(L.lg_topo_shard_indptr(st, order.data_ptr(), cap, kg, j, N, rig.dp._full[0], sip.data_ptr()))
            w_sip, w_six = oracle.fill_topo_shard(w_order, cap, kg, j, indptr, indices)
            assert np.array_equal(sip.cpu().numpy(), w_sip)
            six = torch.empty(max(int(w_sip[-1]), 1), dtype=torch.int32, device=rig.dev)
            capi.check(L.lg_topo_shard_fill(st, order.data_ptr(), cap, kg, j, N, rig.dp._full[0], rig.dp._full[1],
                                            sip.data_ptr(), six.data_ptr()))
            assert np.array_equal(six.cpu().numpy()[: int(w_sip[-1])], w_six)


@pytest.mark.parametrize("variant", [capi.GATHER_LDG, capi.GATHER_TMA])
def test_hybrid_placement_gather(oracle, variant):
    """hybrid placement (replicated head + interleaved tail): directory and shards equal the oracle's for every part, and
    a gather through part j's directory returns the right rows with the head counted as local reads"""
    indptr, indices = small_graph(2500, 9.0, 120)
    N = len(indptr) - 1
    dim = 100
    feat = synth.features(0, N, dim, 8)
    rig = Rig(indptr, indices, feat, [2], 8)
    rng = np.random.default_rng(5)
    hot_np = rng.integers(0, 1000, N).astype(np.uint64)
    order, _ = rig.dp.rank_hotness(torch.from_numpy(hot_np.astype(np.int64)).to(rig.dev))
    w_order, _ = oracle.hotness_rank(hot_np)
    L, st = rig.dp.L, rig.dp._stream()
    kg, cap, rep = 4, 300, 120
    shards = []
    for j in range(kg):
        sh = torch.empty((cap, dim), dtype=torch.float32, device=rig.dev)
        capi.check(L.lg_fill_feature_shard_hybrid(st, order.data_ptr(), cap, kg, rep, j, dim, N, rig.dp._backing, sh.data_ptr()))
        assert np.array_equal(sh.cpu().numpy().view(np.uint32), oracle.fill_feature_shard_hybrid(w_order, cap, kg, rep, j, feat).view(np.uint32))
        shards.append(sh)
    ids = rng.integers(0, N, 6000).astype(np.int32)
    d_ids = torch.from_numpy(ids).to(rig.dev)
    cached_rank = np.full(N, N, np.int64)
    cached_rank[w_order] = np.arange(N)
    for j in (0, 2):
        d = torch.empty(N, dtype=torch.int32, device=rig.dev)
        capi.check(L.lg_fill_i32(st, d.data_ptr(), -2, N))
        capi.check(L.lg_place_features_hybrid(st, order.data_ptr(), cap, kg, rep, j, N, d.data_ptr()))
        assert np.array_equal(d.cpu().numpy(), oracle.place_features_hybrid(w_order, cap, kg, rep, j, N))
        fc = capi.FeatureCache()
        fc.n_parts, fc.shard_rows, fc.dim, fc.num_nodes = kg, cap, dim, N
        for k in range(kg):
            fc.shard[k] = shards[k].data_ptr()
        fc.backing = rig.dp._backing
        fc.directory = d.data_ptr()
        out = torch.empty((len(ids), dim), dtype=torch.float32, device=rig.dev)
        tiers = torch.zeros(3, dtype=torch.int64, device=rig.dev)
        capi.check(L.lg_gather_rows(st, C.byref(fc), d_ids.data_ptr(), len(ids), out.data_ptr(), j, variant, tiers.data_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), feat[ids].view(np.uint32))
        r = cached_rank[ids]
        tail = (r >= rep) & (r < rep + (cap - rep) * kg)
        want_local = int((r < rep).sum() + (tail & ((r - rep) % kg == j)).sum())
        want_peer = int((tail & ((r - rep) % kg != j)).sum())
        assert tiers.cpu().tolist() == [want_local, want_peer, len(ids) - want_local - want_peer]


def test_synth_device_generator_matches_numpy():
    L = capi.load()
    N, dmin, dmax, seed = 30000, 5.25, 700, 0x1e910
    ip = torch.empty(N + 1, dtype=torch.int64, device="cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(L.lg_synth_indptr(st, N, dmin, dmax, seed, ip.data_ptr()))
    w_ip, w_ix = synth.graph(N, dmin, dmax, seed)
    assert np.array_equal(ip.cpu().numpy(), w_ip)
    ix = torch.empty(int(w_ip[-1]), dtype=torch.int32, device="cuda:0")
    capi.check(L.lg_synth_indices(st, N, ip.data_ptr(), seed, ix.data_ptr()))
    assert np.array_equal(ix.cpu().numpy(), w_ix)
    f = torch.empty((500, 100), dtype=torch.float32, device="cuda:0")
    capi.check(L.lg_synth_features(st, 1234, 500, 100, seed, f.data_ptr()))
    assert np.array_equal(f.cpu().numpy().view(np.uint32), synth.features(1234, 500, 100, seed).view(np.uint32))
    lab = torch.empty(N, dtype=torch.int32, device="cuda:0")
    capi.check(L.lg_synth_labels(st, N, 47, lab.data_ptr()))
    assert np.array_equal(lab.cpu().numpy(), synth.labels(N, 47))


@pytest.mark.parametrize("dim", [100, 128, 7])
def test_identity_cache_needs_no_directory(dim):
    """LG_CACHE_IDENTITY: every row resident at row index = vertex id — both movers, -1 padding, all rows counted local"""
    from legion_b200.runner import DataPath
    N = 50000
    feat = synth.features(0, N, dim, 3)
    d_feat = torch.from_numpy(feat).cuda()
    dp = DataPath(0, [2], 8, N, dim)
    dp.set_backing_features(d_feat.data_ptr(), keep=[d_feat])
    dp.build_feature_cache_identity()
    assert dp.cache.directory is None and dp.cache.flags == capi.CACHE_IDENTITY
    rng = np.random.default_rng(dim)
    ids = rng.integers(0, N, 30000).astype(np.int32)
    ids[::97] = -1
    d_ids = torch.from_numpy(ids).cuda()
    for variant in (capi.GATHER_LDG, capi.GATHER_TMA):
        out = torch.full((len(ids), dim), 7.0, dtype=torch.float32, device="cuda")
        tiers = torch.zeros(3, dtype=torch.int64, device="cuda")
        capi.check(dp.L.lg_gather_rows(dp._stream(), C.byref(dp.cache), C.c_void_p(d_ids.data_ptr()), len(ids),
                                       C.c_void_p(out.data_ptr()), 0, variant, C.c_void_p(tiers.data_ptr())))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        ok = ids >= 0
        assert np.array_equal(got[ok].view(np.uint32), feat[ids[ok]].view(np.uint32))
        assert (got[~ok] == 7.0).all()
        assert tiers.cpu().tolist() == [int(ok.sum()), 0, 0]
    dp.close()


@pytest.mark.parametrize("promo", [0, 2, 3])
@pytest.mark.parametrize("dim", [128, 256, 8, 104, 100])
def test_gather4_tensor_copies(monkeypatch, dim, promo):
    """LG_GATHER4=1: groups of 4 rows that all live in the local shard move as one cp.async.bulk.tensor tile::gather4
    (identity placement and a fully cached directory placement); -1 holes, a ragged last tile and dims the tensor path
    does not take (100: 4 rows are not a multiple of 128 bytes) fall back to per-row copies inside the same launch"""
    from legion_b200.runner import DataPath
    monkeypatch.setenv("LG_GATHER4", "1")
    monkeypatch.setenv("LG_GATHER4_PROMO", str(promo))
    N = 40000
    feat = synth.features(0, N, dim, 9)
    d_feat = torch.from_numpy(feat).cuda()
    rng = np.random.default_rng(dim + promo)
    for placement in ("identity", "directory"):
        dp = DataPath(0, [2], 8, N, dim)
        dp.set_backing_features(d_feat.data_ptr(), keep=[d_feat])
        if placement == "identity":
            dp.build_feature_cache_identity()
        else:
            hot = torch.from_numpy(rng.integers(0, 1000, N).astype(np.int64)).cuda()
            order, _ = dp.rank_hotness(hot)
            dp.build_feature_cache(order, cap=N - 5000)  # 12 % of the rows miss: mixed groups
        for n in (3, 8, 33, 20001):
            ids = rng.integers(0, N, n).astype(np.int32)
            if n > 40:
                ids[::61] = -1
            d_ids = torch.from_numpy(ids).cuda()
            out = torch.full((n, dim), 7.0, dtype=torch.float32, device="cuda")
            tiers = torch.zeros(3, dtype=torch.int64, device="cuda")
            capi.check(dp.L.lg_gather_rows(dp._stream(), C.byref(dp.cache), C.c_void_p(d_ids.data_ptr()), n,
                                           C.c_void_p(out.data_ptr()), 0, capi.GATHER_TMA, C.c_void_p(tiers.data_ptr())))
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            ok = ids >= 0
            assert np.array_equal(got[ok].view(np.uint32), feat[ids[ok]].view(np.uint32)), (placement, n)
            assert (got[~ok] == 7.0).all()
            assert int(tiers.sum().item()) == int(ok.sum())
        dp.close()
