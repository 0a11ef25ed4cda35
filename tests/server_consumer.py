"""One trainer process of the server hand-off test (rank == device, like legion_graphsage.py): consumes every batch of
its GPU through the `ipc_service` extension and compares it with the oracle.  argv: gpu gpus data_dir N E D B epochs seed fanout...
Environment: LEGION_IPC_SERVICE_DIR = directory holding the `ipc_service` module to import (default: the repo's
training_backend/; oracle/_ref/ref_trainer = the REFERENCE's extension compiled in place); LEGION_CONSUMER_CHECK = "all"
(default) or k: compare only the first k batches of every mode and the last batch of the run (large datasets);
LEGION_CONSUMER_CSC=1: consume through get_next_csc (B200 extension only) and compare the CSC of every block with the
oracle's; LEGION_EXPECT_HOST_COUNTERS=1/0: assert that the counters did / did not come through the server's side channel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("LEGION_IPC_SERVICE_DIR") or os.path.join(ROOT, "training_backend"))

from legion_b200 import dataset, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    gpu, gpus, data, N, E, D, B, epochs, seed = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], *map(int, sys.argv[4:10])
    fanout = [int(x) for x in sys.argv[10:]]
    import ipc_service
    torch.cuda.set_device(gpu)
    sizes = {k: os.path.getsize(os.path.join(data, k)) // 4 for k in ("trainingset", "validationset", "testingset")}
    check = os.environ.get("LEGION_CONSUMER_CHECK", "all")
    seen = {0: 0, 1: 0, 2: 0}
    print(f"consumer {gpu}: ipc_service from {os.path.dirname(ipc_service.__file__)}", flush=True)
    ds = dataset.read_dataset(data, N, E, D, sizes["trainingset"], sizes["validationset"], sizes["testingset"])
    parts = {k: synth.partition_ids(ds[k], gpus) for k in ("train", "valid", "test")}
    ipc_service.initialize()
    steps = ipc_service.get_steps()
    w_steps, vb, tb, max_step = O.coordinate([len(p) for p in parts["train"]], [len(p) for p in parts["valid"]],
                                             [len(p) for p in parts["test"]], B, epochs)
    assert list(steps) == list(w_steps), (steps, w_steps)
    H = len(fanout)
    sets = {0: (parts["train"][gpu], B), 1: (parts["valid"][gpu], int(vb[gpu])), 2: (parts["test"][gpu], int(tb[gpu]))}
    orc = O.Oracle(ds["indptr"], ds["indices"], fanout, max(B, int(max(vb)), int(max(tb))))
    feat, labels = ds["features"], ds["labels"]
    for g in range(max_step):
        mode, local = O.mode_of(g, w_steps, epochs)
        ids_all, bs = sets[mode]
        use_csc = os.environ.get("LEGION_CONSUMER_CSC") == "1"
        full = ipc_service.get_next_csc(D) if use_csc else ipc_service.get_next(D)
        blk = ipc_service.get_block_size()
        if use_csc:  # [ids, feats, labels, (indptr, indices, eids) per block]: rebuild the COO views for the checks below
            out = list(full[:3])
            for k in range(H):
                e_k = int(full[3 + 3 * k + 1].numel())
                out += [None, None]
        else:
            out = full
        if os.environ.get("LEGION_EXPECT_SERVER_CSC") is not None:
            assert ipc_service.csc_from_server() == (os.environ["LEGION_EXPECT_SERVER_CSC"] == "1")
        exp = os.environ.get("LEGION_EXPECT_HOST_COUNTERS")
        if exp is not None:
            assert ipc_service.counters_from_host() == (exp == "1"), (g, ipc_service.counters_from_host())
        seen[mode] += 1
        if check != "all" and seen[mode] > int(check) and g != max_step - 1:
            assert len(full) == 3 + (3 if use_csc else 2) * H and out[0].numel() == blk[0] and out[1].shape == (blk[0], D)
            ipc_service.synchronize()
            continue
        want = orc.run_batch(ids_all, labels[ids_all], bs, local, seed=seed, batch_id=g, stream_id=gpu)
        n = want["total_nodes"]
        assert len(out) == 3 + 2 * H
        if use_csc:
            torch.cuda.synchronize()
        assert np.array_equal(out[0].cpu().numpy(), want["ids"][:n]), (gpu, g, mode)
        assert np.array_equal(out[1].cpu().numpy().view(np.uint32), feat[want["ids"][:n]].view(np.uint32)), (gpu, g, mode)
        assert np.array_equal(out[2].cpu().numpy(), want["labels"][: want["nc"][9]])
        for k, h in enumerate(range(H, 0, -1)):
            eh = int(want["ec"][9 + h])
            if use_csc:
                w_ip, w_ix, w_eid = O.block_csc(want["agg_src"][:eh], want["agg_dst"][:eh], int(want["nc"][9 + h - 1]))
                ip, ix, eid = (full[3 + 3 * k + j].cpu().numpy() for j in range(3))
                assert np.array_equal(ip, w_ip) and np.array_equal(ix, w_ix) and np.array_equal(eid, w_eid), (gpu, g, h)
            else:
                assert np.array_equal(out[3 + 2 * k].cpu().numpy(), want["agg_src"][:eh])
                assert np.array_equal(out[4 + 2 * k].cpu().numpy(), want["agg_dst"][:eh])
            assert blk[2 * k] == want["nc"][9 + h] and blk[2 * k + 1] == want["nc"][9 + h - 1]
        ipc_service.synchronize()
    ipc_service.finalize()
    print(f"consumer {gpu} ok: {max_step} batches")


if __name__ == "__main__":
    main()
