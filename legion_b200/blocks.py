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
