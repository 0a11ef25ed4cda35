"""helpers shared by the -m gpu tests: build a DataPath on cuda:0 from numpy arrays"""
import numpy as np
import torch

from legion_b200.runner import DataPath, MappedHostBuffer


class Rig:
    def __init__(self, indptr, indices, feat, fanout, max_batch, host_topology=False, host_features=False, device=0):
        self.N, self.D = len(indptr) - 1, feat.shape[1]
        self.indptr, self.indices, self.feat = indptr, indices, feat
        dev = f"cuda:{device}"
        self.dev = dev
        self.dp = DataPath(device, fanout, max_batch, self.N, self.D)
        self.hold = []
        if host_topology:  # full CSR in pinned host memory read through UVA (storage_management.cu:108-109)
            a = MappedHostBuffer(indptr.nbytes)
            a.numpy(np.int64, indptr.shape)[:] = indptr
            b = MappedHostBuffer(max(indices.nbytes, 4))
            b.numpy(np.int32, indices.shape)[:] = indices
            self.hold += [a, b]
            self.dp.set_full_graph(a.dev_ptr, b.dev_ptr)
        else:
            self.d_ip, self.d_ix = torch.from_numpy(indptr).to(dev), torch.from_numpy(indices).to(dev)
            self.dp.set_full_graph(self.d_ip.data_ptr(), self.d_ix.data_ptr())
        if host_features:
            f = MappedHostBuffer(feat.nbytes)
            f.numpy(np.float32, feat.shape)[:] = feat
            self.hold.append(f)
            self.dp.set_backing_features(f.dev_ptr)
        else:
            self.d_feat = torch.from_numpy(feat).to(dev)
            self.dp.set_backing_features(self.d_feat.data_ptr())

    def sets(self, ids, labels):
        self.d_ids = torch.from_numpy(np.ascontiguousarray(ids, np.int32)).to(self.dev)
        self.d_labels = torch.from_numpy(np.ascontiguousarray(labels, np.int32)).to(self.dev)
        return self.d_ids, self.d_labels


def assert_batch_equal(got, want, hops, feat=None):
    assert np.array_equal(got["nc"], want["nc"]), (got["nc"], want["nc"])
    assert np.array_equal(got["ec"], want["ec"]), (got["ec"], want["ec"])
    n, e = want["total_nodes"], want["total_edges"]
    assert np.array_equal(got["ids"], want["ids"][:n])
    B = int(want["nc"][9])
    assert np.array_equal(got["labels"], want["labels"][:B])
    assert np.array_equal(got["agg_src"], want["agg_src"][:e])
    assert np.array_equal(got["agg_dst"], want["agg_dst"][:e])
    if feat is not None:
        ids = want["ids"][:n]
        ok = ids >= 0
        assert np.array_equal(got["features"][ok].view(np.uint32), feat[ids[ok]].view(np.uint32))
