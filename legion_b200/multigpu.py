"""Host-side plumbing for the one-process-per-GPU layout (torch.distributed: nccl on the GPU box, gloo in CPU tests).

Nothing here is on the serving path: the exchange happens once at cache-build time, after which every gather kernel
reads its peers' shards with plain loads over NVLink (reference: pointer tables replicated to the clique members,
cache/cache.cu:572-602; hotness summed over the clique, cache/cache.cu:408-411; steps lock-stepped by the smallest
training shard, engine/ipc_service.cu:73-82)."""


def clique_of(rank, kg):
    """(clique index Ki, slot j inside the clique, first rank of the clique) — device = Ki*Kg + j"""
    return rank // kg, rank % kg, (rank // kg) * kg


def exchange_handles(dist, payload, rank, kg):
    """all-gather one picklable payload per rank (a CUDA IPC handle + size); returns the kg payloads of this
    rank's clique in slot order, own slot included."""
    world = dist.get_world_size()
    assert world % kg == 0, "world size must be a multiple of the clique size"
    everyone = [None] * world
    dist.all_gather_object(everyone, payload)
    _, _, base = clique_of(rank, kg)
    return everyone[base:base + kg]


def aggregate_hotness(dist, hot, rank, kg, group=None):
    """sum of the per-GPU access counters over the clique (aggregate_access); in place"""
    dist.all_reduce(hot, op=dist.ReduceOp.SUM, group=group)
    return hot


def coordinate_train_steps(dist, n_train_local, batch, device=None):
    """train_step = (min over GPUs of the training-shard size - 1) / batch"""
    import torch
    t = torch.tensor([n_train_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return (int(t.item()) - 1) // batch


def exchange_fds(dist, fd, rank, kg, tag="0"):
    """Every rank of a clique owns one file descriptor (the POSIX handle of its VMM cache shard, lg_vmm_alloc) and
    needs its peers' descriptors.  Descriptors only travel over AF_UNIX sockets (SCM_RIGHTS), so each rank listens on
    an abstract-namespace socket, serves its descriptor to the kg-1 peers from a helper thread and fetches theirs.
    Returns the kg descriptors in slot order; the own slot holds `fd` itself.  Received descriptors belong to the
    caller (close them once imported)."""
    import os
    import socket
    import threading
    _, j, base = clique_of(rank, kg)
    port = os.environ.get("MASTER_PORT", "0")
    name = lambda r: f"\0legion_b200_vmm_{port}_{tag}_{r}"  # noqa: E731  (abstract namespace: nothing to unlink)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name(rank))
    srv.listen(kg)

    def serve():
        for _ in range(kg - 1):
            c, _a = srv.accept()
            socket.send_fds(c, [b"fd"], [fd])
            c.close()

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    dist.barrier()  # every listener is up
    got = [None] * kg
    got[j] = fd
    for slot in range(kg):
        if slot == j:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(name(base + slot))
        _msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
        c.close()
        assert len(fds) == 1, "peer sent no descriptor"
        got[slot] = fds[0]
    th.join()
    srv.close()
    dist.barrier()
    return got
