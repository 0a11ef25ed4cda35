"""-m gpu: the drop-in boundary proven on the REFERENCE's own artefacts (built in place into oracle/_ref/ by
oracle/Makefile; nothing of the reference is copied into the repository):

  * the reference's trainer extension (training_backend/ipc_service.cpp + ipc_cuda_kernel.cu:35-235 +
    helper_multiprocess.cpp, compiled by oracle/build_ref_trainer.py) consuming the repo's sampling_server: every
    batch of every mode equals the oracle's;
  * the reference's launcher (legion_server.py:39-110, byte-compiled to oracle/_ref/legion_server.bin — CPython recognises compiled code by its magic number) run UNMODIFIED
    with `--dataset_name products`: it writes meta_config, detects the NVLink clique and execs
    ./sampling_server/build/bin/sampling_server — the repo's binary — which must come up and serve an epoch;
  * the pybind entry sampling_server.Run(fanout, gpu_number, in_memory_mode, cache_mode)
    (sampling_server/sampling_server.cpp:7-22) serving the same consumer.
"""
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
REF_TRAINER = os.path.join(ROOT, "oracle", "_ref", "ref_trainer")
REF_LAUNCHER = os.path.join(ROOT, "oracle", "_ref", "legion_server.bin")


def _clean_ipc():
    for f in os.listdir("/dev/shm"):
        if f.startswith("sem.sem_") or f in ("simpleIPCshm", "legionB200ext"):
            os.unlink(os.path.join("/dev/shm", f))


def _wait_ready(proc, timeout=300):
    t0, lines = time.time(), []
    while time.time() - t0 < timeout:
        line = proc.stdout.readline()
        if not line:
            if proc.poll() is not None:
                break
            continue
        lines.append(line)
        if "System is ready for serving" in line:
            return lines
    raise AssertionError("server did not become ready:\n" + "".join(lines[-40:]))


def _small_dataset(tmp_path, fanout, seed=31):
    from legion_b200 import dataset, synth
    N, D, B, epochs = 7000, 16, 200, 2
    indptr, indices = synth.graph(N, 5.0, 300, seed)
    feat, labels = synth.features(0, N, D, seed), synth.labels(N, 7)
    train, valid, test = synth.split_sets(N, seed, train_frac=0.2, valid=700, test=600)
    data = str(tmp_path / "data") + "/"
    dataset.write_dataset(data, indptr, indices, feat, labels, train, valid, test)
    dataset.write_meta_config(str(tmp_path), data, B, N, len(indices), D, len(train), len(valid), len(test), 200_000, epochs,
                              fanout=fanout)
    return dict(N=N, E=len(indices), D=D, B=B, epochs=epochs, data=data)


def _consume(ds, fanout, seed, module_dir, check="all", timeout=600, extra_env=None):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "server_consumer.py"), "0", "1", ds["data"], str(ds["N"]), str(ds["E"]),
           str(ds["D"]), str(ds["B"]), str(ds["epochs"]), str(seed)] + [str(f) for f in fanout]
    env = dict(os.environ, LEGION_CONSUMER_CHECK=str(check))
    if module_dir:
        env["LEGION_IPC_SERVICE_DIR"] = module_dir
    env.update(extra_env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "consumer 0 ok" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


@pytest.mark.parametrize("fanout", [[25, 10], [4, 3, 2]])
def test_reference_trainer_extension_consumes_our_server(tmp_path, fanout):
    """repo server  ->  simpleIPCshm / semaphores / CUDA IPC  ->  the REFERENCE's ipc_service (compiled in place)"""
    if not any(f.startswith("ipc_service") for f in (os.listdir(REF_TRAINER) if os.path.isdir(REF_TRAINER) else [])):
        pytest.skip("oracle/_ref/ref_trainer not built (needs /root/reference at build time)")
    ds = _small_dataset(tmp_path, fanout)
    _clean_ipc()
    proc = subprocess.Popen([BIN, "1", "0.0"], cwd=str(tmp_path), env=dict(os.environ, LEGION_SEED="777"),
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        _wait_ready(proc)
        out = _consume(ds, fanout, 777, REF_TRAINER)
        assert "oracle/_ref/ref_trainer" in out  # the module really was the reference's
        tail, _ = proc.communicate(timeout=60)
        assert "Server Stopped" in tail and proc.returncode == 0, tail[-2000:]
        assert '"legion_b200_telemetry"' in tail  # per-tier rows / GB/s line of the server
    finally:
        if proc.poll() is None:
            proc.kill()


def test_pybind_run_serves_the_trainer(tmp_path):
    """sampling_server.Run(fanout, gpu_number, in_memory_mode, cache_mode) — the in-process API of the reference
    (sampling_server/sampling_server.cpp:7-22) — serves the same hand-off, with the fan-out passed from Python"""
    fanout = [6, 4]
    ds = _small_dataset(tmp_path, None)  # no fan-out in meta_config: Run()'s argument decides
    _clean_ipc()
    code = f"import sys; sys.path.insert(0, {os.path.join(ROOT, 'sampling_server')!r}); import sampling_server; " \
           f"sys.exit(sampling_server.Run({fanout!r}, 1, 1, 0))"
    proc = subprocess.Popen([sys.executable, "-c", code], cwd=str(tmp_path), env=dict(os.environ, LEGION_SEED="99"),
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        _wait_ready(proc)
        # the B200 trainer extension on the other end: counters through the side channel, blocks as CSC
        _consume(ds, fanout, 99, None, extra_env={"LEGION_CONSUMER_CSC": "1", "LEGION_EXPECT_HOST_COUNTERS": "1",
                                                  "LEGION_EXPECT_SERVER_CSC": "0"})
        tail, _ = proc.communicate(timeout=60)
        assert "Server Stopped" in tail and proc.returncode == 0, tail[-2000:]
    finally:
        if proc.poll() is None:
            proc.kill()


@pytest.mark.parametrize("fanout", [[25, 10], [4, 3, 2]])
def test_server_emits_blocks_as_csc(tmp_path, fanout):
    """LEGION_EMIT_CSC=1: the server builds the CSC of every block on its third stream (lg_block_csc_batch, sizes read from the
    batch's counters on the device) into extra CUDA-IPC buffers; get_next_csc returns views of them — compared with the
    oracle's stable COO -> CSC for every block of every batch of every mode"""
    ds = _small_dataset(tmp_path, fanout)
    _clean_ipc()
    proc = subprocess.Popen([BIN, "1", "0.0"], cwd=str(tmp_path), env=dict(os.environ, LEGION_SEED="8", LEGION_EMIT_CSC="1"),
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        head = "".join(_wait_ready(proc))
        assert "Blocks are emitted as CSC" in head
        _consume(ds, fanout, 8, None, extra_env={"LEGION_CONSUMER_CSC": "1", "LEGION_EXPECT_HOST_COUNTERS": "1",
                                                 "LEGION_EXPECT_SERVER_CSC": "1"})
        tail, _ = proc.communicate(timeout=60)
        assert "Server Stopped" in tail and proc.returncode == 0, tail[-2000:]
    finally:
        if proc.poll() is None:
            proc.kill()


def test_trainer_extension_falls_back_without_side_channel(tmp_path):
    """a server that offers only the reference wire (LEGION_EXT_SHM=0, i.e. what a reference server looks like to the
    trainer): the B200 trainer extension reads the counters with the reference's two device copies and still agrees"""
    fanout = [5, 5]
    ds = _small_dataset(tmp_path, fanout)
    _clean_ipc()
    proc = subprocess.Popen([BIN, "1", "0.0"], cwd=str(tmp_path), env=dict(os.environ, LEGION_SEED="5", LEGION_EXT_SHM="0"),
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        _wait_ready(proc)
        assert not os.path.exists("/dev/shm/legionB200ext")
        _consume(ds, fanout, 5, None, extra_env={"LEGION_EXPECT_HOST_COUNTERS": "0"})
        tail, _ = proc.communicate(timeout=60)
        assert "Server Stopped" in tail and proc.returncode == 0, tail[-2000:]
    finally:
        if proc.poll() is None:
            proc.kill()


def test_unmodified_reference_launcher_starts_our_server(tmp_path, oracle):
    """python legion_server.py --dataset_name products ... (the reference's launcher, unmodified) from a cwd that holds
    ./sampling_server/build/bin/sampling_server = the repo's binary.  The launcher hard-codes the products shape
    (legion_server.py:41-48: 2,449,029 vertices, 123,718,280 edges, 100-d, 196,615 / 39,323 / 2,213,091 ids), so the
    dataset written here has exactly those sizes."""
    if not os.path.exists(REF_LAUNCHER):
        pytest.skip("oracle/_ref/legion_server.bin not built (needs /root/reference at build time)")
    pytest.importorskip("networkx")
    from legion_b200 import dataset, synth
    N, E_META, D = 2_449_029, 123_718_280, 100
    n_train, n_valid, n_test = 196_615, 39_323, 2_213_091
    work = "/dev/shm/legion_launcher_test"
    subprocess.run(["rm", "-rf", work])
    os.makedirs(os.path.join(work, "dataset"), exist_ok=True)
    try:
        oracle.lib().lgo_set_num_threads(os.cpu_count() or 1)
        indptr, indices = oracle.synth_graph(N, synth.dmin_for(N, E_META), 20000, 5)
        E = len(indices)
        assert E <= E_META
        feat = oracle.synth_features(0, N, D, 5)
        labels = synth.labels(N, 47)
        perm = np.random.default_rng(5).permutation(N).astype(np.int32)
        train, valid, test = perm[:n_train], perm[n_train:n_train + n_valid], perm[n_train + n_valid:n_train + n_valid + n_test]
        data = os.path.join(work, "dataset", "products") + "/"
        dataset.write_dataset(data, indptr, indices, feat, labels, train, valid, test)
        with open(os.path.join(data, "edge_dst"), "ab") as f:  # the launcher's edge count: pad the file, the CSR never points there
            f.write(np.zeros(E_META - E, np.int32).tobytes())
        del feat
        os.makedirs(os.path.join(work, "sampling_server", "build", "bin"))
        os.symlink(BIN, os.path.join(work, "sampling_server", "build", "bin", "sampling_server"))
        _clean_ipc()
        proc = subprocess.Popen([sys.executable, REF_LAUNCHER, "--dataset_name", "products", "--dataset_path", os.path.join(work, "dataset"),
                                 "--gpu_number", "1", "--epoch", "1", "--cache_memory", "40000000000", "--usenvlink", "1"],
                                cwd=work, env=dict(os.environ, LEGION_SEED="4711"), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        try:
            head = "".join(_wait_ready(proc, timeout=600))
            assert "Train Steps: 24" in head  # (196615 - 1) / 8000
            meta = open(os.path.join(work, "meta_config")).read().split()
            assert meta[1:8] == ["8000", str(N), str(E_META), str(D), str(n_train), str(n_valid), str(n_test)]
            ds = dict(N=N, E=E, D=D, B=8000, epochs=1, data=data)
            _consume(ds, [25, 10], 4711, None, check=2, timeout=900)  # fan-out hard-coded in the reference (src/main.cu:9-11)
            tail, _ = proc.communicate(timeout=120)
            assert "Server Stopped" in tail, tail[-2000:]
            assert "NVLink clique size" in head + tail
        finally:
            if proc.poll() is None:
                proc.kill()
    finally:
        subprocess.run(["rm", "-rf", work])
