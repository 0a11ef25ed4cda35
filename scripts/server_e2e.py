#!/usr/bin/env python
"""The drop-in boundary end to end at full size: the C++ sampling_server binary (meta_config, dataset files, presampling,
cost model, cache fill, simpleIPCshm + semaphores + CUDA-IPC buffers) feeding a consumer that calls the `ipc_service`
torch extension exactly like legion_graphsage.py:74-75,90 (get_next, get_block_size, synchronize) and touches nothing
else.  Prints one JSON line: seeds/s over the training steps of the run (wall clock around the consumer loop).
usage: server_e2e.py [--workload products] [--scale 1.0] [--epochs 10] [--cache-gb 100] [--dir /dev/shm/legion_ds]"""
import argparse, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "training_backend"))
import bench
from legion_b200 import dataset, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="products")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--cache-gb", type=float, default=100.0)
ap.add_argument("--dir", default="/dev/shm/legion_ds")
ap.add_argument("--csc", action="store_true", help="consume through get_next_csc: the CSC of every block (lg_block_csc), built by the trainer extension")
ap.add_argument("--server-csc", action="store_true", help="with --csc: the SERVER builds the blocks (LEGION_EMIT_CSC=1) and get_next_csc returns views")
args = ap.parse_args()

import torch
shape = bench.shape_of(args)
N, D, B, fanout = shape["N"], shape["D"], shape["batch"], shape["fanout"]
t0 = time.time()
ip, ix, feat, lab, E = bench.device_dataset(shape, 0)
if feat is None:
    feat = torch.empty((N, D), dtype=torch.float32, device="cuda")
    from legion_b200 import capi
    import ctypes as C
    capi.check(capi.load().lg_synth_features(C.c_void_p(torch.cuda.current_stream().cuda_stream), 0, N, D, bench.SEED, feat.data_ptr()))
train, valid, test = synth.split_sets(N, bench.SEED)
data = args.dir.rstrip("/") + "/"
dataset.write_dataset(data, ip.cpu().numpy(), ix.cpu().numpy(), feat.cpu().numpy(), lab.cpu().numpy(), train, valid, test)
del ip, ix, feat, lab
torch.cuda.empty_cache()
cwd = args.dir.rstrip("/") + "_cwd"
os.makedirs(cwd, exist_ok=True)
dataset.write_meta_config(cwd, data, B, N, E, D, len(train), len(valid), len(test), int(args.cache_gb * 1e9), args.epochs, fanout=fanout)
print(f"[server_e2e] dataset written in {time.time() - t0:.1f}s: N={N} E={E} D={D}", file=sys.stderr)
for f in os.listdir("/dev/shm"):
    if f.startswith("sem.sem_") or f == "simpleIPCshm":
        os.unlink(os.path.join("/dev/shm", f))
BIN = os.path.join(ROOT, "sampling_server", "build", "bin", "sampling_server")
t0 = time.time()
proc = subprocess.Popen([BIN, "1", "0.0"], cwd=cwd, env=dict(os.environ, LEGION_SEED=str(bench.SEED), LEGION_EMIT_CSC="1" if args.server_csc else "0"), stdout=subprocess.PIPE,
                        stderr=subprocess.STDOUT, text=True)
head = []
while True:
    line = proc.stdout.readline()
    if not line:
        raise SystemExit("server died:\n" + "".join(head))
    head.append(line)
    if "System is ready for serving" in line:
        break
t_ready = time.time() - t0
import ipc_service
torch.cuda.set_device(0)
ipc_service.initialize()
steps = ipc_service.get_steps()
train_steps, valid_steps, test_steps = [int(x) for x in steps]
max_step = (train_steps + valid_steps) * args.epochs + test_steps
bb = None
H = len(fanout)
t_train, n_train, rows, edges = 0.0, 0, 0, 0
g = 0
for ep in range(args.epochs):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for s in range(train_steps):
        out = ipc_service.get_next_csc(D) if args.csc else ipc_service.get_next(D)  # csc: (indptr, indices, eids) per block
        sizes = ipc_service.get_block_size()
        if args.csc:
            torch.cuda.synchronize()  # the blocks are built on this process's stream: wait for them like a layer would
        rows += out[0].numel(); edges += out[4 if args.csc else 3].numel()
        ipc_service.synchronize()
        g += 1
    t_train += time.perf_counter() - t1
    n_train += train_steps
    for s in range(valid_steps):
        ipc_service.get_next(D); ipc_service.synchronize(); g += 1
for s in range(test_steps):
    ipc_service.get_next(D); ipc_service.synchronize(); g += 1
from_host = bool(ipc_service.counters_from_host())
csc_server = bool(ipc_service.csc_from_server())
ipc_service.finalize()
tail, _ = proc.communicate(timeout=120)
assert proc.returncode == 0 and "Server Stopped" in tail, tail[-2000:]
caps = [l.strip() for l in head if "capacity" in l or "Alpha" in l or "Preprocessing" in l]
telemetry = [json.loads(l)["legion_b200_telemetry"] for l in tail.splitlines() if l.startswith('{"legion_b200_telemetry"')]
print(json.dumps({"what": "sampling_server binary -> simpleIPCshm/semaphores/CUDA IPC -> ipc_service consumer (1 GPU)",
                  "workload": shape["name"], "num_nodes": N, "num_edges": E, "feature_dim": D, "batch": B, "fanout": fanout,
                  "train_steps_per_epoch": train_steps, "epochs": args.epochs, "seeds_per_s": n_train * B / t_train,
                  "ms_per_batch": 1e3 * t_train / n_train, "rows_per_batch": rows / n_train, "edges_per_batch": edges / n_train,
                  "consumer_builds_csc": bool(args.csc), "counters_from_host": from_host, "csc_built_by_server": csc_server, "server_ready_s": round(t_ready, 1), "server_says": caps, "tier_telemetry": telemetry}))
