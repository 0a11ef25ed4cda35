"""Host-side mirror of the reference's per-GPU runner for the data path.

Mirrors GPURunner (sampling_server/src/engine/server.cu:170-365) and the pieces of
StorageManagement / UnifiedCache it drives, on top of the C ABI (include/legion_b200.h):

    GPURunner.Initialize            -> DataPath.__init__  (scratch, batch buffers)
    GPURunner.RunPreSc              -> DataPath.run_presc (sampling-only ops + hotness)
    UnifiedCache.CandidateSelection -> DataPath.rank_hotness
    UnifiedCache.FillUp             -> DataPath.build_feature_cache / build_topology_cache
    GPURunner.RunOnce               -> DataPath.run_once  (ops 0..3*hops+3 on one stream)

PyTorch is used for device memory and streams only.  Multi-GPU follows the launch contract of
bench.py: one process per GPU; peers' cache shards are mapped with CUDA IPC handles exchanged
through torch.distributed (plumbing) and read with plain peer loads inside the kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import Batch, BatchParams, FeatureCache, Topology, check

I32 = torch.int32


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class RawDeviceBuffer:
    """cudaMalloc'd (legacy-IPC exportable) device memory, viewable as a torch tensor."""

    def __init__(self, nbytes, device):
        self.L = capi.load()
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        with torch.cuda.device(device):
            check(self.L.lg_device_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        self.owned = True

    @classmethod
    def from_ipc(cls, handle, nbytes, device):
        self = cls.__new__(cls)
        self.L = capi.load()
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        h = (C.c_char * 64).from_buffer_copy(handle)
        with torch.cuda.device(device):
            check(self.L.lg_ipc_open(h, C.byref(p)))
        self.ptr = p.value
        self.owned = False
        return self

    @classmethod
    def vmm(cls, nbytes, device):
        """cuMemCreate-backed memory whose POSIX descriptor other processes map at full page size (cache shards of
        the one-process-per-GPU deployment; see lg_vmm_alloc)"""
        self = cls.__new__(cls)
        self.L = capi.load()
        self.device = device
        self.nbytes = int(nbytes)
        p, fd = C.c_void_p(), C.c_int32(-1)
        with torch.cuda.device(device):
            check(self.L.lg_vmm_alloc(self.nbytes, C.byref(p), C.byref(fd)))
        self.ptr, self.fd = p.value, fd.value
        self.owned = "vmm"
        return self

    @classmethod
    def from_vmm_fd(cls, fd, nbytes, device):
        self = cls.__new__(cls)
        self.L = capi.load()
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        with torch.cuda.device(device):
            check(self.L.lg_vmm_import(int(fd), self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.owned = "vmm"
        return self

    def ipc_handle(self):
        h = (C.c_char * 64)()
        check(self.L.lg_ipc_export(C.c_void_p(self.ptr), h))
        return bytes(h.raw)

    def tensor(self, dtype, shape):
        np_dt = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
        holder = self

        class _CAI:
            __cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": np_dt,
                                        "data": (holder.ptr, False), "version": 2}
            _keep = holder

        return torch.as_tensor(_CAI(), device=f"cuda:{self.device}")

    def free(self):
        if self.ptr:
            if self.owned == "vmm":
                self.L.lg_vmm_free(C.c_void_p(self.ptr))
            elif self.owned:
                self.L.lg_device_free(C.c_void_p(self.ptr))
            else:
                self.L.lg_ipc_close(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):  # shards are tens of GB: a dropped cache must give its HBM back
        try:
            self.free()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class MappedHostBuffer:
    """cudaHostAllocMapped memory (storage/storage_management.cu:108-109): numpy view + UVA device pointer."""

    def __init__(self, nbytes):
        self.L = capi.load()
        hp, dp = C.c_void_p(), C.c_void_p()
        check(self.L.lg_host_alloc_mapped(C.byref(hp), C.byref(dp), int(nbytes)))
        self.host_ptr, self.dev_ptr, self.nbytes = hp.value, dp.value, int(nbytes)

    def numpy(self, dtype, shape):
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(self.host_ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def free(self):
        if self.host_ptr:
            self.L.lg_host_free(C.c_void_p(self.host_ptr))
            self.host_ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


class SharedHostBuffer:
    """One host copy of a large read-only array for all the ranks of a box: a /dev/shm file that every rank maps and
    page-locks (lg_host_register), readable by the kernels through UVA like the reference's cudaHostAllocMapped arrays
    (storage/storage_management.cu:108-109) — which live once in the reference's single server process.  The name is
    unlinked as soon as every rank holds a mapping.  world == 1: a plain MappedHostBuffer."""

    def __init__(self, nbytes, rank=0, world=1, dist=None, tag="buf"):
        import mmap
        import os
        import time
        self.L = capi.load()
        self.nbytes = int(nbytes)
        self._plain = None
        if world == 1:
            self._plain = MappedHostBuffer(self.nbytes)
            self.host_ptr, self.dev_ptr = self._plain.host_ptr, self._plain.dev_ptr
            return
        tok = [f"/dev/shm/legion_b200_{tag}_{os.getpid()}_{int(time.time() * 1e3)}" if rank == 0 else None]
        dist.broadcast_object_list(tok, src=0)
        path = tok[0]
        if rank == 0:
            fd0 = os.open(path, os.O_RDWR | os.O_CREAT, 0o600)
            os.posix_fallocate(fd0, 0, max(self.nbytes, 4096))  # the pages exist before anybody page-locks them
            os.close(fd0)
        dist.barrier()
        fd = os.open(path, os.O_RDWR)
        self._mm = mmap.mmap(fd, max(self.nbytes, 4096))
        os.close(fd)
        dist.barrier()
        if rank == 0:
            os.unlink(path)
        buf = (C.c_char * self.nbytes).from_buffer(self._mm)
        self.host_ptr = C.addressof(buf)
        self._keep = buf
        dp = C.c_void_p()
        check(self.L.lg_host_register(C.c_void_p(self.host_ptr), self.nbytes, C.byref(dp)))
        self.dev_ptr = dp.value

    def numpy(self, dtype, shape):
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(self.host_ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def copy_from_device(self, src_ptr, byte_off, nbytes, stream):
        """device -> this buffer at byte_off; split at 1 GB boundaries (lg_host_register page-locks in 1 GB chunks and
        one cudaMemcpy may not span two registrations)"""
        chunk = 1 << 30
        done = 0
        while done < nbytes:
            off = byte_off + done
            n = min(nbytes - done, chunk - (off % chunk))
            check(self.L.lg_memcpy_d2h(C.c_void_p(self.host_ptr + off), C.c_void_p(src_ptr + done), n, stream))
            done += n

    def free(self):
        if self._plain is not None:
            self._plain.free()
        elif self.host_ptr:
            self.L.lg_host_unregister(C.c_void_p(self.host_ptr))
        self.host_ptr = None


class BatchBuffers:
    """The seven IPC buffers of one (gpu, pipeline slot) — engine/ipc_service.cu:134-206."""

    def __init__(self, device, batch_size, num_ids, feature_rows, dim, exportable=False):
        self.device, self.num_ids, self.feature_rows, self.dim = device, int(num_ids), int(feature_rows), dim
        dev = f"cuda:{device}"
        self._raw = []

        def alloc(n, dtype):
            if exportable:
                item = 4
                raw = RawDeviceBuffer(max(int(n), 1) * item, device)
                self._raw.append(raw)
                return raw.tensor(dtype, (int(n),))
            return torch.empty(int(n), dtype=dtype, device=dev)

        self.ids = alloc(num_ids, I32)
        self.features = alloc(self.feature_rows * dim, torch.float32).view(self.feature_rows, dim)
        self.labels = alloc(batch_size, I32)
        self.agg_src = alloc(num_ids, I32)
        self.agg_dst = alloc(num_ids, I32)
        self.node_counter = alloc(16, I32)
        self.edge_counter = alloc(16, I32)
        self.node_counter.zero_()
        self.edge_counter.zero_()
        self.c = Batch(ids=self.ids.data_ptr(), features=self.features.data_ptr(), labels=self.labels.data_ptr(),
                       agg_src=self.agg_src.data_ptr(), agg_dst=self.agg_dst.data_ptr(),
                       node_counter=self.node_counter.data_ptr(), edge_counter=self.edge_counter.data_ptr(),
                       feature_rows=self.feature_rows, num_ids=self.num_ids, reserved=0)

    def to_host(self, hops):
        """the trainer's view (training_backend/ipc_cuda_kernel.cu:194-232) copied to numpy"""
        nc = self.node_counter.cpu().numpy()
        ec = self.edge_counter.cpu().numpy()
        n, e = int(nc[9 + hops]), int(ec[9 + hops])
        return dict(nc=nc, ec=ec, ids=self.ids[:n].cpu().numpy(), labels=self.labels[:int(nc[9])].cpu().numpy(),
                    agg_src=self.agg_src[:e].cpu().numpy(), agg_dst=self.agg_dst[:e].cpu().numpy(),
                    features=self.features[:n].cpu().numpy() if self.feature_rows >= n else None,
                    total_nodes=n, total_edges=e)


class DataPath:
    """One GPU's sampler + cache views + batch buffers."""

    def __init__(self, device, fanout, max_batch, num_nodes, dim, rank=0, world=1):
        self.L = capi.load()
        self.device = int(device)
        self.fanout = [int(f) for f in fanout]
        self.hops = len(self.fanout)
        self.max_batch = int(max_batch)
        self.N = int(num_nodes)
        self.dim = int(dim)
        self.rank, self.world = rank, world
        torch.cuda.set_device(self.device)
        fo = (C.c_int32 * self.hops)(*self.fanout)
        self.num_ids = int(self.L.lg_num_ids(self.max_batch, fo, self.hops))
        h = C.c_void_p()
        check(self.L.lg_sampler_create(self.device, self.max_batch, fo, self.hops, self.N, C.byref(h)))
        self.sampler = h
        self.topo = Topology()
        self.cache = FeatureCache()
        self._keep = []  # tensors referenced by the descriptors
        self._cache_keep = []  # buffers owned by the current feature cache (shard, peer mappings)
        self.tier_rows = torch.zeros(3, dtype=torch.int64, device=f"cuda:{self.device}")
        self.local_part = 0

    def close(self):
        if self.sampler:
            self.L.lg_sampler_destroy(self.sampler)
            self.sampler = None

    # ---- storage (GraphStorage::Build / FeatureStorage::Build) ----
    def set_full_graph(self, indptr_ptr, indices_ptr, keep=()):
        """slot P of the pointer tables = full CSR (storage/graph_storage.cu:60-62); HBM or host UVA pointer"""
        self._full = (int(indptr_ptr), int(indices_ptr))
        self._keep.extend(keep)
        self._set_topology([], [], None, 0)

    def _set_topology(self, shard_indptr_ptrs, shard_indices_ptrs, directory, shard_rows):
        t = Topology()
        t.n_parts = len(shard_indptr_ptrs)
        t.shard_rows = int(shard_rows)
        t.num_nodes = self.N
        for i, (a, b) in enumerate(zip(shard_indptr_ptrs, shard_indices_ptrs)):
            t.indptr[i], t.indices[i] = int(a), int(b)
        t.indptr[t.n_parts], t.indices[t.n_parts] = self._full
        t.directory = directory.data_ptr() if directory is not None else None
        self.topo = t
        self.topo_directory = directory

    def repoint_full_graph(self, indptr_ptr, indices_ptr, drop=()):
        """slot P now points at another copy of the full CSR (e.g. the pinned-host one once the HBM copy used for
        presampling and the cache fill is released); the cached shards and the directory stay"""
        self._full = (int(indptr_ptr), int(indices_ptr))
        t = self.topo
        t.indptr[t.n_parts], t.indices[t.n_parts] = self._full
        self._keep = [k for k in self._keep if not any(k is d for d in drop)]

    def set_backing_features(self, ptr, keep=()):
        self._backing = int(ptr)
        self._keep.extend(keep)
        self._set_cache([], None, 0)

    def _set_cache(self, shard_ptrs, directory, shard_rows, flags=0):
        c = FeatureCache()
        c.flags = int(flags)
        c.n_parts = len(shard_ptrs)
        c.shard_rows = int(shard_rows)
        c.dim = self.dim
        c.num_nodes = self.N
        for i, p in enumerate(shard_ptrs):
            c.shard[i] = int(p)
        c.backing = self._backing if self._backing else None
        c.directory = directory.data_ptr() if directory is not None else None
        self.cache = c
        self.feat_directory = directory

    # ---- cache construction ----
    def rank_hotness(self, hotness):
        """CandidateSelection (cache/cache.cu:399-440) for one already-aggregated u64[N] array (int64 tensor)."""
        n = hotness.numel()
        order = torch.empty(n, dtype=I32, device=hotness.device)
        sorted_h = torch.empty(n, dtype=torch.int64, device=hotness.device)
        nb = C.c_int64(0)
        st = self._stream()
        check(self.L.lg_hotness_rank(st, None, n, None, None, None, C.byref(nb)))
        tmp = torch.empty(nb.value, dtype=torch.uint8, device=hotness.device)
        check(self.L.lg_hotness_rank(st, _ptr(hotness), n, _ptr(order), _ptr(sorted_h), _ptr(tmp), C.byref(nb)))
        torch.cuda.current_stream().synchronize()
        return order, sorted_h

    def build_feature_cache(self, order, cap, kg=1, j=0, peers=None, dist=None, replicate=0):
        """FillUp feature part (cache/cache.cu:565-602).  kg > 1: this process fills shard j and maps the
        others through CUDA IPC (`dist` = torch.distributed module, already initialised).  replicate = rows at the
        head of every shard that hold the hottest ranks on EVERY part (hybrid placement, lg_place_features_hybrid)."""
        st = self._stream()
        dev = f"cuda:{self.device}"
        directory = torch.empty(self.N, dtype=I32, device=dev)
        check(self.L.lg_fill_i32(st, _ptr(directory), capi.CACHEMISS_FLAG, self.N))
        check(self.L.lg_place_features_hybrid(st, _ptr(order), cap, kg, replicate, j, self.N, _ptr(directory)))
        raw = self._shard_alloc(cap * self.dim * 4, kg)
        check(self.L.lg_fill_feature_shard_hybrid(st, _ptr(order), cap, kg, replicate, j, self.dim, self.N,
                                                  C.c_void_p(self._backing), C.c_void_p(raw.ptr)))
        torch.cuda.current_stream().synchronize()
        shard_ptrs = [0] * kg
        shard_ptrs[j] = raw.ptr
        self._cache_keep = [raw]
        if kg > 1:
            shard_ptrs = self._exchange(raw, cap * self.dim * 4, kg, j, dist, keep=self._cache_keep)
        self.local_part = j
        self._set_cache(shard_ptrs, directory, cap)
        self.feat_shard = raw
        return directory

    def build_feature_cache_synth(self, order, cap, seed, kg=1, j=0, dist=None, keep_backing=False, replicate=0):
        """as build_feature_cache, but the shard is generated in place from the synthetic feature function
        (include/legion_b200_synth.h): no [N x D] backing matrix exists anywhere (paper-scale shapes)"""
        st = self._stream()
        dev = f"cuda:{self.device}"
        directory = torch.empty(self.N, dtype=I32, device=dev)
        check(self.L.lg_fill_i32(st, _ptr(directory), capi.CACHEMISS_FLAG, self.N))
        check(self.L.lg_place_features_hybrid(st, _ptr(order), cap, kg, replicate, j, self.N, _ptr(directory)))
        raw = self._shard_alloc(cap * self.dim * 4, kg)
        check(self.L.lg_synth_feature_shard_hybrid(st, _ptr(order), cap, kg, replicate, j, self.dim, self.N, seed,
                                                   C.c_void_p(raw.ptr)))
        torch.cuda.current_stream().synchronize()
        shard_ptrs = [0] * kg
        shard_ptrs[j] = raw.ptr
        self._cache_keep = [raw]
        if kg > 1:
            shard_ptrs = self._exchange(raw, cap * self.dim * 4, kg, j, dist, keep=self._cache_keep)
        self.local_part = j
        if not keep_backing:
            self._backing = 0
        self._set_cache(shard_ptrs, directory, cap)
        self.feat_shard = raw
        return directory

    def build_feature_cache_identity(self, seed=None):
        """A cache that holds every vertex on this GPU: rows stored at row index = vertex id, no directory
        (LG_CACHE_IDENTITY).  seed given: the rows are generated in place from the synthetic feature function; else copied
        from the backing matrix."""
        st = self._stream()
        raw = self._shard_alloc(self.N * self.dim * 4, 1)
        if seed is not None:
            for r0 in range(0, self.N, 1 << 24):
                check(self.L.lg_synth_features(st, r0, min(1 << 24, self.N - r0), self.dim, seed,
                                               C.c_void_p(raw.ptr + r0 * self.dim * 4)))
        else:
            check(self.L.lg_memcpy_d2d(C.c_void_p(raw.ptr), C.c_void_p(self._backing), self.N * self.dim * 4, st))
        torch.cuda.current_stream().synchronize()
        self._cache_keep = [raw]
        self.local_part = 0
        self._set_cache([raw.ptr], None, self.N, flags=capi.CACHE_IDENTITY)
        self.feat_shard = raw
        return None

    def build_topology_cache(self, order, cap, kg=1, j=0, ki=0, dist=None):
        """GraphCache (storage/graph_storage.cu:76-111) + topology directory (cache/cache.cu:116-129)"""
        st = self._stream()
        dev = f"cuda:{self.device}"
        directory = torch.empty(self.N, dtype=I32, device=dev)
        check(self.L.lg_fill_i32(st, _ptr(directory), capi.CACHEMISS_FLAG, self.N))
        check(self.L.lg_place_topology(st, _ptr(order), cap, kg, ki, self.N, _ptr(directory)))
        ip = self._shard_alloc((cap + 1) * 8, kg)
        full_ip, full_ix = self._full
        check(self.L.lg_topo_shard_indptr(st, _ptr(order), cap, kg, j, self.N, C.c_void_p(full_ip), C.c_void_p(ip.ptr)))
        torch.cuda.current_stream().synchronize()
        total = int(ip.tensor(torch.int64, (cap + 1,))[cap].item())
        ix = self._shard_alloc(max(total, 1) * 4, kg)
        check(self.L.lg_topo_shard_fill(st, _ptr(order), cap, kg, j, self.N, C.c_void_p(full_ip), C.c_void_p(full_ix),
                                        C.c_void_p(ip.ptr), C.c_void_p(ix.ptr)))
        torch.cuda.current_stream().synchronize()
        self._keep += [ip, ix]
        ips, ixs = [0] * kg, [0] * kg
        ips[j], ixs[j] = ip.ptr, ix.ptr
        if kg > 1:
            ips = self._exchange(ip, (cap + 1) * 8, kg, j, dist)
            ixs = self._exchange(ix, None, kg, j, dist, nbytes_own=max(total, 1) * 4)
        self._set_topology(ips, ixs, directory, cap)
        self.topo_shard = (ip, ix, total)
        return directory

    def _shard_alloc(self, nbytes, kg):
        """cache shards that peers in OTHER processes will map are VMM allocations: a legacy cudaIpcOpenMemHandle
        mapping is read through small pages and random row reads out of a multi-GB peer shard collapse to ~90 GB/s
        (profiles/r01b_peer_mapping.md).  LG_SHARD_IPC=legacy keeps the reference's cudaIpc handles."""
        import os
        if kg > 1 and os.environ.get("LG_SHARD_IPC", "vmm") != "legacy":
            return RawDeviceBuffer.vmm(nbytes, self.device)
        return RawDeviceBuffer(nbytes, self.device)

    def drop_feature_cache(self):
        """release the current feature cache (shard, peer mappings, directory); every rank of the clique calls it
        before any rank frees — the caller puts a barrier in front"""
        self.feat_shard = None
        self.feat_directory = None
        for b in self._cache_keep:
            b.free()
        self._cache_keep = []
        self._set_cache([], None, 0)

    def _exchange(self, raw, nbytes, kg, j, dist, nbytes_own=None, keep=None):
        """one buffer per clique member -> the kg device pointers, own slot first-hand, peers mapped: VMM descriptors
        passed over AF_UNIX sockets, or (LG_SHARD_IPC=legacy) all-gathered CUDA IPC handles"""
        from .multigpu import exchange_fds, exchange_handles
        assert dist is not None and dist.is_initialized(), "multi-GPU cache needs torch.distributed"
        own_bytes = nbytes if nbytes is not None else nbytes_own
        self._xchg = getattr(self, "_xchg", 0) + 1
        keep = self._keep if keep is None else keep
        ptrs = []
        if raw.owned == "vmm":
            import os
            sizes = exchange_handles(dist, own_bytes, self.rank, kg)
            fds = exchange_fds(dist, raw.fd, self.rank, kg, tag=str(self._xchg))
            for p in range(kg):
                if p == j:
                    ptrs.append(raw.ptr)
                else:
                    peer = RawDeviceBuffer.from_vmm_fd(fds[p], sizes[p], self.device)
                    os.close(fds[p])
                    keep.append(peer)
                    ptrs.append(peer.ptr)
            dist.barrier()  # every import is done before anyone may close its exported descriptor
            os.close(raw.fd)
            raw.fd = -1
            return ptrs
        clique = exchange_handles(dist, (raw.ipc_handle(), own_bytes), self.rank, kg)
        for p in range(kg):
            if p == j:
                ptrs.append(raw.ptr)
            else:
                h, nb = clique[p]
                peer = RawDeviceBuffer.from_ipc(h, nb, self.device)
                keep.append(peer)
                ptrs.append(peer.ptr)
        return ptrs

    # ---- batches ----
    def alloc_batch(self, feature_rows=None, exportable=False):
        rows = self.num_ids if feature_rows is None else int(feature_rows)
        return BatchBuffers(self.device, self.max_batch, self.num_ids, rows, self.dim, exportable)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def params(self, all_ids, all_labels, batch_size, counter, mode=capi.TRAINMODE, rng_kind=capi.RNG_PHILOX,
               seed=0, batch_id=0, stream_id=None):
        return BatchParams(all_ids=all_ids.data_ptr(), all_labels=all_labels.data_ptr(), total_cap=all_ids.numel(),
                           batch_size=batch_size, counter=counter, mode=mode, rng_kind=rng_kind, batch_id=batch_id,
                           stream_id=self.rank if stream_id is None else stream_id, local_part=self.local_part,
                           rng_seed=seed)

    def run_once(self, p, buf, gather=True, tier=False):
        """GPURunner::RunOnce ops (engine/server.cu:311-317) on the current stream; no host sync."""
        check(self.L.lg_run_batch(self.sampler, self._stream(), C.byref(self.topo),
                                  C.byref(self.cache) if gather else None, C.byref(p), C.byref(buf.c),
                                  _ptr(self.tier_rows) if tier else None))

    def run_once_host(self, p, host_ids, host_labels, buf, host_nc, host_ec, gather=True):
        """end-to-end form: host seeds in, counters out (lg_run_batch_host)"""
        check(self.L.lg_run_batch_host(self.sampler, self._stream(), C.byref(self.topo),
                                       C.byref(self.cache) if gather else None, C.byref(p),
                                       C.c_void_p(host_ids.ctypes.data), C.c_void_p(host_labels.ctypes.data),
                                       C.byref(buf.c), C.c_void_p(host_nc.ctypes.data), C.c_void_p(host_ec.ctypes.data)))

    def run_once_host_async(self, p, host_ids, host_labels, buf, host_nc, host_ec, gather=True):
        """as run_once_host without the final synchronisation (host buffers must be pinned)"""
        check(self.L.lg_run_batch_host_async(self.sampler, self._stream(), C.byref(self.topo),
                                             C.byref(self.cache) if gather else None, C.byref(p),
                                             C.c_void_p(host_ids.ctypes.data), C.c_void_p(host_labels.ctypes.data),
                                             C.byref(buf.c), C.c_void_p(host_nc.ctypes.data),
                                             C.c_void_p(host_ec.ctypes.data)))

    def run_presc(self, p, buf, edge_hot, node_hot, max_ids):
        """GPURunner::RunPreSc (engine/server.cu:285-300): ops 0,3,6,..,last with is_presc=true"""
        st = self._stream()
        check(self.L.lg_batch_generate(self.sampler, st, C.c_void_p(p.all_ids), C.c_void_p(p.all_labels), p.total_cap,
                                       p.batch_size, p.counter, C.byref(buf.c)))
        for hop in range(1, self.hops + 1):
            check(self.L.lg_random_sample(self.sampler, st, C.byref(self.topo), hop, p.rng_kind, p.rng_seed,
                                          p.batch_id, p.stream_id, C.byref(buf.c), _ptr(edge_hot)))
        check(self.L.lg_io_complete(self.sampler, st, capi.TRAINMODE, C.byref(buf.c), _ptr(node_hot), _ptr(max_ids)))

    def status(self):
        s = C.c_int32(0)
        check(self.L.lg_sampler_status(self.sampler, self._stream(), C.byref(s)))
        return s.value

    def set_overlap(self, mode):
        """0 one stream, 1 gathers overlap the next hop (joined per batch), 2 pipelined across batches"""
        check(self.L.lg_sampler_set_overlap(self.sampler, int(mode)))

    def set_tail_mode(self, mode):
        """capi.TAIL_EXACT (default) or capi.TAIL_REFERENCE (the reference's stride of the clipped tail batch)"""
        check(self.L.lg_sampler_set_tail_mode(self.sampler, int(mode)))

    def set_gather_fusion(self, mode):
        """0 one gather per op, 1 seeds ride with hop 1, 2 single gather per batch"""
        check(self.L.lg_sampler_set_gather_fusion(self.sampler, int(mode)))

    def share_storage_from(self, other):
        """a second in-flight runner on the same GPU: same topology / cache descriptors, own scratch"""
        self._full, self._backing = other._full, other._backing
        self.topo, self.cache = other.topo, other.cache
        self.topo_directory, self.feat_directory = getattr(other, "topo_directory", None), getattr(other, "feat_directory", None)
        self.local_part = other.local_part
        self._keep.append(other)

    def batch_wait(self, buf):
        check(self.L.lg_batch_wait(self.sampler, self._stream(), C.byref(buf.c)))

    def set_gather_variant(self, v):
        check(self.L.lg_sampler_set_gather_variant(self.sampler, v))
