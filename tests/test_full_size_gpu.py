"""-m gpu: BASELINE.json configs[1] at full size (products shape, B=8000, fanout [25,10]).
Direct oracle comparison (the oracle needs ~0.1 s per batch at this size) plus size-independent
properties: uniqueness, index bounds, E_h = sum min(deg, fanout), gather == index_select,
run-to-run determinism, both gather movers."""
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

from legion_b200 import capi, synth  # noqa: E402
from legion_b200.runner import DataPath  # noqa: E402

SEED = 0x1E910


@pytest.fixture(scope="module")
def products():
    L = capi.load()
    N, E_t, D, classes = synth.SHAPES["products"]
    dev = "cuda:0"
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ip = torch.empty(N + 1, dtype=torch.int64, device=dev)
    capi.check(L.lg_synth_indptr(st, N, synth.dmin_for(N, E_t), 20000, SEED, ip.data_ptr()))
    E = int(ip[N].item())
    ix = torch.empty(E, dtype=torch.int32, device=dev)
    capi.check(L.lg_synth_indices(st, N, ip.data_ptr(), SEED, ix.data_ptr()))
    feat = torch.empty((N, D), dtype=torch.float32, device=dev)
    capi.check(L.lg_synth_features(st, 0, N, D, SEED, feat.data_ptr()))
    train = synth.split_sets(N, SEED)[0]
    d_train = torch.from_numpy(train).to(dev)
    d_lab = (d_train % classes).to(torch.int32)
    assert abs(E - E_t) / E_t < 0.02
    return dict(N=N, E=E, D=D, ip=ip, ix=ix, feat=feat, train=train, d_train=d_train, d_lab=d_lab)


def test_full_size_batches(oracle, products):
    g = products
    B, fanout = 8000, [25, 10]
    dp = DataPath(0, fanout, B, g["N"], g["D"])
    dp.set_full_graph(g["ip"].data_ptr(), g["ix"].data_ptr())
    dp.set_backing_features(g["feat"].data_ptr())
    hot = torch.bincount(g["ix"].long(), minlength=g["N"])
    order, _ = dp.rank_hotness(hot)
    dp.build_feature_cache(order, cap=g["N"])  # fully HBM-cached, rows permuted by hotness rank
    buf = dp.alloc_batch()
    h_ip, h_ix = g["ip"].cpu().numpy(), g["ix"].cpu().numpy()
    orc = oracle.Oracle(h_ip, h_ix, fanout, B)
    labels = (g["train"] % 47).astype(np.int32)
    deg = (g["ip"][1:] - g["ip"][:-1])
    prev = None
    for variant, counter in [(capi.GATHER_LDG, 0), (capi.GATHER_TMA, 7), (capi.GATHER_TMA, 7)]:
        dp.set_gather_variant(variant)
        p = dp.params(g["d_train"], g["d_lab"], B, counter, seed=SEED, batch_id=counter)
        dp.run_once(p, buf)
        torch.cuda.synchronize()
        assert dp.status() == 0
        nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
        n, e1, e = int(nc[11]), int(ec[10]), int(ec[11])
        ids = buf.ids[:n]
        # properties
        assert nc[8] == 2 and nc[9] == B and torch.unique(ids).numel() == n
        seeds = g["d_train"][counter * B:(counter + 1) * B]
        assert torch.equal(ids[:B], seeds)
        assert e1 == int(torch.clamp(deg[seeds.long()], max=25).sum().item())
        src, dst = buf.agg_src[:e], buf.agg_dst[:e]
        assert int(dst[:e1].max()) < B and int(src[:e1].max()) < nc[10]
        assert int(dst[e1:].max()) < nc[10] and int(src[e1:].max()) < n and int(src.min()) >= 0
        frontier2 = ids[src[:e1].long()]
        assert e - e1 == int(torch.clamp(deg[frontier2.long()], max=10).sum().item())
        assert torch.equal(buf.features[:n].view(torch.int32), g["feat"][ids.long()].view(torch.int32))
        # oracle, bit-exact
        want = orc.run_batch(g["train"], labels, B, counter, seed=SEED, batch_id=counter)
        assert np.array_equal(nc, want["nc"]) and np.array_equal(ec, want["ec"])
        assert np.array_equal(ids.cpu().numpy(), want["ids"][:n])
        assert np.array_equal(src.cpu().numpy(), want["agg_src"][:e]) and np.array_equal(dst.cpu().numpy(), want["agg_dst"][:e])
        snap = (ids.clone(), src.clone(), dst.clone())
        if prev is not None and counter == 7 and prev[3] == 7:
            assert all(torch.equal(a, b) for a, b in zip(prev[:3], snap))  # determinism
        prev = snap + (counter,)
    dp.close()


def test_full_size_three_hops_minstd(oracle, products):
    """BASELINE.json configs[4] fan-out [15,10,5] (on the products-shaped graph) with the reference's own RNG stream"""
    g = products
    B, fanout = 8000, [15, 10, 5]
    dp = DataPath(0, fanout, B, g["N"], g["D"])
    dp.set_full_graph(g["ip"].data_ptr(), g["ix"].data_ptr())
    dp.set_backing_features(g["feat"].data_ptr())
    buf = dp.alloc_batch(feature_rows=1)
    p = dp.params(g["d_train"], g["d_lab"], B, 3, rng_kind=capi.RNG_MINSTD, batch_id=3)
    dp.run_once(p, buf, gather=False)
    torch.cuda.synchronize()
    orc = oracle.Oracle(g["ip"].cpu().numpy(), g["ix"].cpu().numpy(), fanout, B)
    want = orc.run_batch(g["train"], (g["train"] % 47).astype(np.int32), B, 3, rng_kind=oracle.RNG_MINSTD, batch_id=3)
    nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
    assert np.array_equal(nc[[0, 1, 6, 7, 8, 9, 10, 11, 12]], want["nc"][[0, 1, 6, 7, 8, 9, 10, 11, 12]])
    assert np.array_equal(ec, want["ec"])
    n, e = int(nc[12]), int(ec[12])
    assert np.array_equal(buf.ids[:n].cpu().numpy(), want["ids"][:n])
    assert np.array_equal(buf.agg_src[:e].cpu().numpy(), want["agg_src"][:e])
    assert np.array_equal(buf.agg_dst[:e].cpu().numpy(), want["agg_dst"][:e])
    dp.close()


@pytest.mark.parametrize("B,fanout", [(1000, [20, 10]), (2500, [20, 4]), (600, [30, 30])])
def test_mid_size_frontiers(oracle, products, B, fanout):
    """frontier lengths that select the 64- and 128-entry sample tiles and the 8/12/16-edge rank tiles
    (csrc/sampler.cu pick_tile_f / pick_rank_items): 20 k, 50 k and 18 k entries; two batches each, bit-exact"""
    g = products
    dp = DataPath(0, fanout, B, g["N"], g["D"])
    dp.set_full_graph(g["ip"].data_ptr(), g["ix"].data_ptr())
    dp.set_backing_features(g["feat"].data_ptr())
    buf = dp.alloc_batch(feature_rows=1)
    orc = oracle.Oracle(g["ip"].cpu().numpy(), g["ix"].cpu().numpy(), fanout, B)
    labels = (g["train"] % 47).astype(np.int32)
    for counter in (2, 5):
        p = dp.params(g["d_train"], g["d_lab"], B, counter, seed=SEED, batch_id=counter)
        dp.run_once(p, buf, gather=False)
        torch.cuda.synchronize()
        assert dp.status() == 0
        want = orc.run_batch(g["train"], labels, B, counter, seed=SEED, batch_id=counter)
        nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
        H = len(fanout)
        assert np.array_equal(nc[9:], want["nc"][9:]) and np.array_equal(ec, want["ec"])
        n, e = int(nc[9 + H]), int(ec[9 + H])
        assert np.array_equal(buf.ids[:n].cpu().numpy(), want["ids"][:n])
        assert np.array_equal(buf.agg_src[:e].cpu().numpy(), want["agg_src"][:e])
        assert np.array_equal(buf.agg_dst[:e].cpu().numpy(), want["agg_dst"][:e])
    dp.close()
