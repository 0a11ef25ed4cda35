"""world_size-2 gloo test (CPU): the host-side logic of the N>1 path — seed sharding, clique handle exchange,
hotness aggregation + ranking, lock-stepped step count."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from legion_b200 import multigpu, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, B = 5000, 100
    train, _, _ = synth.split_sets(N, 7, train_frac=0.3)
    parts = synth.partition_ids(train, world)
    mine = parts[rank]
    steps = multigpu.coordinate_train_steps(dist, len(mine), B)
    # each rank's "IPC handle" is a recognisable blob; the exchange must return the clique in slot order
    got = multigpu.exchange_handles(dist, (bytes([rank]) * 64, 1000 + rank), rank, world)
    # hotness: every rank counts its own seeds, the sum must equal the global count
    hot = torch.zeros(N, dtype=torch.int64)
    hot[torch.from_numpy(mine.astype(np.int64))] += 1
    multigpu.aggregate_hotness(dist, hot, rank, world)
    # descriptor exchange (what carries the VMM shard handles): every rank passes an open pipe's read end
    rd, wr = os.pipe()
    os.write(wr, bytes([65 + rank]) * 4)
    fds = multigpu.exchange_fds(dist, rd, rank, world, tag="t")
    seen = [os.read(fds[p], 4) if p != rank else b"" for p in range(world)]
    for p in range(world):
        if p != rank:
            os.close(fds[p])
    os.close(rd)
    os.close(wr)
    out[rank] = dict(steps=steps, got=got, hot=hot.numpy().copy(), n=len(mine), ids=mine.copy(), seen=seen)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic(oracle):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    from legion_b200 import synth
    train, _, _ = synth.split_sets(5000, 7, train_frac=0.3)
    assert r0["n"] + r1["n"] == len(train)
    assert (r0["ids"] % 2 == 0).all() and (r1["ids"] % 2 == 1).all()  # id % gpus (storage_management.cu:175-179)
    want_steps = (min(r0["n"], r1["n"]) - 1) // 100
    assert r0["steps"] == r1["steps"] == want_steps
    for r in (r0, r1):
        assert [g[0][0] for g in r["got"]] == [0, 1] and [g[1] for g in r["got"]] == [1000, 1001]
    assert r0["seen"] == [b"", b"BBBB"] and r1["seen"] == [b"AAAA", b""]  # each rank read its peer's pipe through the passed fd
    want_hot = np.bincount(train, minlength=5000)
    assert np.array_equal(r0["hot"], want_hot) and np.array_equal(r1["hot"], want_hot)
    # both ranks derive the same ranking -> the same interleaved placement (rank r -> GPU r % Kg, row r / Kg)
    order, _ = oracle.hotness_rank(r0["hot"].astype(np.uint64))
    d = oracle.place_features(order, 1500, 2, 5000)
    owners = d[order[:3000]] // 1500
    assert np.array_equal(owners, np.arange(3000) % 2)


def test_clique_arithmetic():
    from legion_b200 import multigpu
    assert multigpu.clique_of(5, 4) == (1, 1, 4)
    assert multigpu.clique_of(7, 8) == (0, 7, 0)
    assert multigpu.clique_of(2, 1) == (2, 0, 2)


def _csr_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from legion_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    indptr, indices = synth.graph(3000, 6.0, 200, 11)  # every rank holds the identical graph (as it does in HBM)
    csr = bench.HostCSR(torch.from_numpy(indptr), torch.from_numpy(indices), 3000, len(indices), rank, world, dist)
    leftovers = [f for f in os.listdir("/dev/shm") if f.startswith("legion_b200_csr_")]
    out[rank] = dict(ok=bool(np.array_equal(csr.indptr, indptr) and np.array_equal(csr.indices, indices)),
                     shared=isinstance(csr.indptr, np.memmap), leftovers=leftovers)
    dist.barrier()
    dist.destroy_process_group()


def test_host_csr_is_shared_once_per_box():
    """bench.py's parity self-check (N > 1): one host copy of the CSR in /dev/shm, each rank writes its slice, every rank
    sees the whole graph, and the names are unlinked as soon as everybody has them mapped"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_csr_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        assert out[r]["ok"] and out[r]["shared"] and out[r]["leftovers"] == [], out[r]
