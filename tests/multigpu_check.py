"""Run under torchrun with N >= 2 ranks (one per GPU): the partitioned feature/topology cache over CUDA IPC.
Every rank samples its own batches; rows are read from the local shard, from peers over NVLink, or from the
backing matrix (miss).  Checks, per rank: batch bit-exact vs the oracle, features bit-exact, tier mix as placed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from legion_b200 import capi, synth  # noqa: E402
from legion_b200.runner import DataPath  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    N, D, B, fanout = 30000, 100, 256, [10, 5]
    indptr, indices = synth.graph(N, 7.0, 400, 5)
    feat = synth.features(0, N, D, 5)
    train = synth.partition_ids(synth.split_sets(N, 5, train_frac=0.3)[0], world)[rank]
    labels = synth.labels(N, 47)[train]
    d_ip, d_ix = torch.from_numpy(indptr).to(dev), torch.from_numpy(indices).to(dev)
    d_feat = torch.from_numpy(feat).to(dev)
    d_train, d_lab = torch.from_numpy(train).to(dev), torch.from_numpy(labels).to(dev)
    dp = DataPath(local, fanout, B, N, D, rank=rank, world=world)
    dp.set_full_graph(d_ip.data_ptr(), d_ix.data_ptr(), keep=[d_ip, d_ix])
    dp.set_backing_features(d_feat.data_ptr(), keep=[d_feat])
    hot = torch.from_numpy(np.bincount(indices, minlength=N).astype(np.int64)).to(dev)
    dist.all_reduce(hot)  # same on every rank anyway; exercises the init-time collective
    hot //= world
    order, _ = dp.rank_hotness(hot)
    cap_f, cap_t = N // (2 * world), N // (4 * world)  # half of the table cached, a quarter of the topology
    fdir = dp.build_feature_cache(order, cap_f, kg=world, j=rank, dist=dist)
    dp.build_topology_cache(order, cap_t, kg=world, j=rank, ki=0, dist=dist)
    buf = dp.alloc_batch()
    orc = O.Oracle(indptr, indices, fanout, B)
    for variant in (capi.GATHER_TMA, capi.GATHER_LDG):
        dp.set_gather_variant(variant)
        dp.tier_rows.zero_()
        for it in range(3):
            p = dp.params(d_train, d_lab, B, it, seed=11, batch_id=it)
            dp.run_once(p, buf, tier=True)
            dp.batch_wait(buf)
            torch.cuda.synchronize()
            got = buf.to_host(2)
            want = orc.run_batch(train, labels, B, it, seed=11, batch_id=it, stream_id=rank)
            n, e = want["total_nodes"], want["total_edges"]
            assert np.array_equal(got["nc"], want["nc"]) and np.array_equal(got["ec"], want["ec"])
            assert np.array_equal(got["ids"], want["ids"][:n])
            assert np.array_equal(got["agg_src"], want["agg_src"][:e]) and np.array_equal(got["agg_dst"], want["agg_dst"][:e])
            assert np.array_equal(got["features"].view(np.uint32), feat[want["ids"][:n]].view(np.uint32))
        tiers = dp.tier_rows.cpu().numpy()
        assert tiers[0] > 0 and tiers[1] > 0 and tiers[2] > 0, tiers  # local, peer (NVLink), miss
    d = fdir.cpu().numpy()
    owner = d[d >= 0] // cap_f
    assert set(owner.tolist()) == set(range(world))
    assert dp.status() == 0
    dist.barrier()
    if rank == 0:
        print(f"multigpu_check ok: world={world} tiers(local,peer,miss)={tiers.tolist()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
