"""-m gpu: a graph large enough that the library SELECTS the hashed position map by itself (4 N bytes > 48 MB: the layout
every paper-scale shape runs with), with the host tiers of the reference in play at the same time: the full CSR in
pinned host memory read through UVA (storage/storage_management.cu:100-115, engine/operator_impl.cu:224-243), the hottest
adjacency lists in an HBM topology shard (storage/graph_storage.cu:76-111), the hottest feature rows in an HBM shard and
the rest in a pinned host backing matrix (cache/cache_impl.cuh:262-266).  Batches are compared with the oracle, gathered
rows with the feature function, the per-tier row counts with the placement."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from legion_b200 import capi, synth  # noqa: E402
from legion_b200.runner import DataPath, MappedHostBuffer  # noqa: E402

SEED = 77


def test_auto_hashed_map_with_host_tiers(oracle, monkeypatch):
    monkeypatch.delenv("LG_DEDUP", raising=False)
    monkeypatch.delenv("LG_DENSE_MAX_MB", raising=False)
    N, D, B, fanout = 13_000_000, 8, 2000, [10, 5]
    oracle.lib().lgo_set_num_threads(16)
    indptr, indices = oracle.synth_graph(N, synth.dmin_for(N, 4 * N), 2000, SEED)  # ~52 M edges, generated on the host cores
    E = len(indices)
    dev = "cuda:0"
    dp = DataPath(0, fanout, B, N, D)
    assert dp.L.lg_sampler_dedup_layout(dp.sampler) == 1, "4N bytes > 48 MB must select the hashed position map"
    # full CSR in mapped host memory (slot P of the pointer tables); a device copy only for the cache fill
    h_ip, h_ix = MappedHostBuffer((N + 1) * 8), MappedHostBuffer(E * 4)
    h_ip.numpy(np.int64, (N + 1,))[:] = indptr
    h_ix.numpy(np.int32, (E,))[:] = indices
    d_ip, d_ix = torch.from_numpy(indptr).to(dev), torch.from_numpy(indices).to(dev)
    dp.set_full_graph(d_ip.data_ptr(), d_ix.data_ptr(), keep=[d_ip, d_ix])
    # backing feature matrix in mapped host memory
    h_feat = MappedHostBuffer(N * D * 4)
    feat = h_feat.numpy(np.float32, (N, D))
    oracle.synth_features(0, N, D, SEED, out=feat)
    dp.set_backing_features(h_feat.dev_ptr, keep=[h_feat])
    # hotness = in-degree; hottest 20 % of the rows and 10 % of the adjacency lists go to HBM
    hot = torch.bincount(d_ix.long(), minlength=N)
    order, _ = dp.rank_hotness(hot)
    cap_f, cap_t = N // 5, N // 10
    dp.build_feature_cache(order, cap_f)
    dp.build_topology_cache(order, cap_t)
    dp.repoint_full_graph(h_ip.dev_ptr, h_ix.dev_ptr, drop=[d_ip, d_ix])  # misses now read the host CSR through UVA
    del d_ip, d_ix
    torch.cuda.empty_cache()
    train = synth.split_sets(N, SEED, train_frac=0.01)[0]
    labels = (train % 5).astype(np.int32)
    d_train, d_lab = torch.from_numpy(train).to(dev), torch.from_numpy(labels).to(dev)
    buf = dp.alloc_batch()
    orc = oracle.Oracle(indptr, indices, fanout, B)
    h_order = order.cpu().numpy()
    in_feat_cache = np.zeros(N, bool)
    in_feat_cache[h_order[:cap_f]] = True
    for variant, counter in ((capi.GATHER_TMA, 0), (capi.GATHER_LDG, 3), (capi.GATHER_TMA, 4)):
        dp.set_gather_variant(variant)
        dp.tier_rows.zero_()
        p = dp.params(d_train, d_lab, B, counter, seed=SEED, batch_id=counter)
        dp.run_once(p, buf, tier=True)
        dp.batch_wait(buf)
        torch.cuda.synchronize()
        assert dp.status() == 0
        want = orc.run_batch(train, labels, B, counter, seed=SEED, batch_id=counter)
        n, e = want["total_nodes"], want["total_edges"]
        nc, ec = buf.node_counter.cpu().numpy(), buf.edge_counter.cpu().numpy()
        assert np.array_equal(nc, want["nc"]) and np.array_equal(ec, want["ec"])
        ids = buf.ids[:n].cpu().numpy()
        assert np.array_equal(ids, want["ids"][:n])
        assert np.array_equal(buf.agg_src[:e].cpu().numpy(), want["agg_src"][:e])
        assert np.array_equal(buf.agg_dst[:e].cpu().numpy(), want["agg_dst"][:e])
        assert np.array_equal(buf.features[:n].cpu().numpy().view(np.uint32), feat[ids].view(np.uint32))
        tiers = dp.tier_rows.cpu().numpy()
        hits = int(in_feat_cache[ids].sum())
        assert tiers[0] == hits and tiers[1] == 0 and tiers[2] == n - hits and 0 < hits < n, (tiers, hits, n)
    dp.close()
