"""-m gpu: the drop-in boundary end to end.  The C++ sampling_server binary (meta_config, argv, ready line,
simpleIPCshm, semaphores, CUDA-IPC buffers) feeds a consumer that uses the `ipc_service` torch extension
exactly like legion_graphsage.py:74-75,90 does; every batch of every mode is compared with the oracle."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "sampling_server", "build", "bin", "sampling_server")


def _wait_ready(proc, timeout=120):
    t0 = time.time()
    lines = []
    while time.time() - t0 < timeout:
        line = proc.stdout.readline()
        if not line:
            if proc.poll() is not None:
                break
            continue
        lines.append(line)
        if "System is ready for serving" in line:
            return lines
    raise AssertionError("server did not become ready:\n" + "".join(lines))


@pytest.mark.parametrize("fanout,cache_bytes", [([25, 10], 60_000), ([4, 3, 2], 10_000_000)])
def test_server_to_trainer_handoff(oracle, tmp_path, fanout, cache_bytes):
    from legion_b200 import dataset, synth
    assert os.path.exists(BIN), "sampling_server binary not built (run __graft_entry__.build())"
    sys.path.insert(0, os.path.join(ROOT, "training_backend"))
    import ipc_service  # the trainer-side extension

    N, D, B, epochs = 6000, 16, 200, 2
    indptr, indices = synth.graph(N, 5.0, 300, 21)
    feat = synth.features(0, N, D, 21)
    labels = synth.labels(N, 7)
    train, valid, test = synth.split_sets(N, 21, train_frac=0.2, valid=700, test=600)
    data = str(tmp_path / "data") + "/"
    dataset.write_dataset(data, indptr, indices, feat, labels, train, valid, test)
    cwd = str(tmp_path)
    dataset.write_meta_config(cwd, data, B, N, len(indices), D, len(train), len(valid), len(test), cache_bytes, epochs,
                              fanout=fanout)
    for f in os.listdir("/dev/shm"):
        if f.startswith("sem.sem_") or f == "simpleIPCshm":
            os.unlink(os.path.join("/dev/shm", f))
    env = dict(os.environ, LEGION_SEED="12345")
    proc = subprocess.Popen([BIN, "1", "0.0"], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        head = "".join(_wait_ready(proc))
        assert "Train Steps:" in head and "Feat capacity:" in head
        torch.cuda.set_device(0)
        ipc_service.initialize()
        steps = ipc_service.get_steps()
        w_steps, vb, tb, max_step = oracle.coordinate([len(train)], [len(valid)], [len(test)], B, epochs)
        assert list(steps) == list(w_steps)
        H = len(fanout)
        sets = {0: (train, B), 1: (valid, int(vb[0])), 2: (test, int(tb[0]))}
        orc = oracle.Oracle(indptr, indices, fanout, max(B, int(vb[0]), int(tb[0])))
        for g in range(max_step):
            mode, local = oracle.mode_of(g, w_steps, epochs)
            ids_all, bs = sets[mode]
            out = ipc_service.get_next(D)
            sizes = ipc_service.get_block_size()
            want = orc.run_batch(ids_all, labels[ids_all], bs, local, seed=12345, batch_id=g, stream_id=0)
            n, e = want["total_nodes"], want["total_edges"]
            assert len(out) == 3 + 2 * H
            assert np.array_equal(out[0].cpu().numpy(), want["ids"][:n]), (g, mode)
            got_f = out[1].cpu().numpy()
            assert np.array_equal(got_f.view(np.uint32), feat[want["ids"][:n]].view(np.uint32)), (g, mode)
            assert np.array_equal(out[2].cpu().numpy(), want["labels"][: want["nc"][9]])
            for k, h in enumerate(range(H, 0, -1)):  # block h = cumulative edges of hops 1..h
                eh = int(want["ec"][9 + h])
                assert np.array_equal(out[3 + 2 * k].cpu().numpy(), want["agg_src"][:eh])
                assert np.array_equal(out[4 + 2 * k].cpu().numpy(), want["agg_dst"][:eh])
                assert sizes[2 * k] == want["nc"][9 + h] and sizes[2 * k + 1] == want["nc"][9 + h - 1]
            ipc_service.synchronize()
        ipc_service.finalize()
        tail, _ = proc.communicate(timeout=60)
        assert "Server Stopped" in tail
        assert proc.returncode == 0
    finally:
        if proc.poll() is None:
            proc.kill()


@pytest.mark.parametrize("gpus,agg_mode,cache_bytes,replicate", [(2, "1.0", 150_000, "0"), (2, "0.0", 10_000_000, "0"),
                                                                   (2, "1.0", 150_000, "0.4")])
def test_server_multi_gpu_handoff(oracle, tmp_path, gpus, agg_mode, cache_bytes, replicate):
    """one server process, one thread per GPU (engine/server.cu:122-130), partitioned (Kg=2) or replicated (Kg=1) cache
    read over cudaDeviceEnablePeerAccess — also with the hybrid placement (LEGION_REPLICATE_RATIO: the head of every
    shard replicated on both GPUs); one trainer process per GPU (rank == device) checks every batch"""
    from legion_b200 import dataset, synth
    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs >= {gpus} GPUs")
    N, D, B, epochs, seed, fanout = 8000, 16, 200, 2, 4242, [10, 5]
    indptr, indices = synth.graph(N, 5.0, 300, 22)
    feat = synth.features(0, N, D, 22)
    labels = synth.labels(N, 7)
    train, valid, test = synth.split_sets(N, 22, train_frac=0.25, valid=900, test=700)
    data = str(tmp_path / "data") + "/"
    dataset.write_dataset(data, indptr, indices, feat, labels, train, valid, test)
    cwd = str(tmp_path)
    dataset.write_meta_config(cwd, data, B, N, len(indices), D, len(train), len(valid), len(test), cache_bytes, epochs, fanout=fanout)
    for f in os.listdir("/dev/shm"):
        if f.startswith("sem.sem_") or f == "simpleIPCshm":
            os.unlink(os.path.join("/dev/shm", f))
    proc = subprocess.Popen([BIN, str(gpus), agg_mode], cwd=cwd, env=dict(os.environ, LEGION_SEED=str(seed), LEGION_REPLICATE_RATIO=replicate),
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    consumers = []
    try:
        _wait_ready(proc)
        for g in range(gpus):
            cmd = [sys.executable, os.path.join(ROOT, "tests", "server_consumer.py"), str(g), str(gpus), data, str(N), str(len(indices)),
                   str(D), str(B), str(epochs), str(seed)] + [str(f) for f in fanout]
            consumers.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        for g, c in enumerate(consumers):
            out, _ = c.communicate(timeout=300)
            assert c.returncode == 0 and f"consumer {g} ok" in out, out[-3000:]
        tail, _ = proc.communicate(timeout=60)
        assert "Server Stopped" in tail and proc.returncode == 0, tail[-2000:]
    finally:
        for c in consumers:
            if c.poll() is None:
                c.kill()
        if proc.poll() is None:
            proc.kill()
