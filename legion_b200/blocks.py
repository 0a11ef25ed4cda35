"""Trainer-side block construction on top of the C ABI.

The reference trainer rebuilds each layer's DGL block from the COO views returned by
`ipc_service.get_next` (training_backend/legion_graphsage.py:66-79) and DGL converts it to CSC for the
SpMM on every step.  `BlockBuilder.csc` does that conversion once, on the trainer's CUDA stream,
directly from the CUDA-IPC buffers; with DGL the result is consumed as

    g = dgl.create_block(('csc', (indptr, indices, eids)), num_src_nodes=num_src, num_dst_nodes=num_dst)

(`indptr` over destinations, `indices` = sources, `eids` = positions in the COO, so edge data keeps the COO
order).  PyTorch is used for device memory only.
"""
import ctypes as C

import torch

from . import capi


class BlockBuilder:
    def __init__(self, max_edges, device=None):
        self.L = capi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_edges = int(max_edges)
        nb = C.c_int64(0)
        capi.check(self.L.lg_block_csc_workspace(self.max_edges, C.byref(nb)))
        self.workspace = torch.empty(nb.value, dtype=torch.uint8, device=self.device)

    def csc(self, src, dst, num_dst, with_eids=True):
        """src, dst: int32 CUDA tensors of one block's COO (batch-local indices); returns (indptr, indices, eids)"""
        e = int(src.numel())
        assert e <= self.max_edges and dst.numel() == e and src.dtype == torch.int32 and dst.dtype == torch.int32
        indptr = torch.empty(int(num_dst) + 1, dtype=torch.int32, device=self.device)
        indices = torch.empty(e, dtype=torch.int32, device=self.device)
        eids = torch.empty(e, dtype=torch.int32, device=self.device) if with_eids else None
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        capi.check(self.L.lg_block_csc(st, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), e, int(num_dst),
                                       C.c_void_p(indptr.data_ptr()), C.c_void_p(indices.data_ptr()),
                                       C.c_void_p(eids.data_ptr()) if with_eids else None,
                                       C.c_void_p(self.workspace.data_ptr()), self.workspace.numel()))
        return indptr, indices, eids

    def csc_batch(self, batch_c, hops, max_edges, max_dst):
        """every block of one batch in one set of launches, sizes read from the batch's counters on the device
        (lg_block_csc_batch — what the server runs next to the gather).  batch_c: capi.Batch of the batch buffers;
        max_edges[h-1] / max_dst[h-1] bound block h.  Returns [(indptr, indices, eids) for h = 1..hops], allocated at the
        bounds: the caller slices them with the counters."""
        me = (C.c_int64 * hops)(*[int(x) for x in max_edges])
        md = (C.c_int32 * hops)(*[int(x) for x in max_dst])
        nb = C.c_int64(0)
        capi.check(self.L.lg_block_csc_batch_workspace(hops, me, C.byref(nb)))
        if self.workspace.numel() < nb.value:
            self.workspace = torch.empty(nb.value, dtype=torch.uint8, device=self.device)
        out = [(torch.empty(int(max_dst[h]) + 1, dtype=torch.int32, device=self.device),
                torch.empty(int(max_edges[h]), dtype=torch.int32, device=self.device),
                torch.empty(int(max_edges[h]), dtype=torch.int32, device=self.device)) for h in range(hops)]
        arr = lambda k: (C.c_void_p * hops)(*[o[k].data_ptr() for o in out])  # noqa: E731
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        capi.check(self.L.lg_block_csc_batch(st, C.byref(batch_c), hops, me, md, arr(0), arr(1), arr(2),
                                             C.c_void_p(self.workspace.data_ptr()), self.workspace.numel()))
        return out
