"""Legion on-disk dataset format (reference dataset/README.md:3-10; loaders storage/storage_management.cu:100-232):

    <path>/edge_src       int64[N+1]   CSR offsets
    <path>/edge_dst       int32[E]     CSR neighbours
    <path>/features       fp32[N x D]
    <path>/labels         int32[N]
    <path>/trainingset    int32[ntrain]   (validationset, testingset likewise)
    <path>/partition      int32[N]        optional (xtrapulp output; absent => id % gpus)

plus the one-line `meta_config` the launcher writes into the server's cwd (legion_server.py:94-95):
    path batch N E D ntrain nvalid ntest cache_bytes epochs [fanout ...]
`path` must end with '/' because the server concatenates file names to it.
"""
import os

import numpy as np


def write_dataset(path, indptr, indices, features, labels, train, valid, test, partition=None):
    os.makedirs(path, exist_ok=True)
    np.ascontiguousarray(indptr, np.int64).tofile(os.path.join(path, "edge_src"))
    np.ascontiguousarray(indices, np.int32).tofile(os.path.join(path, "edge_dst"))
    np.ascontiguousarray(features, np.float32).tofile(os.path.join(path, "features"))
    np.ascontiguousarray(labels, np.int32).tofile(os.path.join(path, "labels"))
    np.ascontiguousarray(train, np.int32).tofile(os.path.join(path, "trainingset"))
    np.ascontiguousarray(valid, np.int32).tofile(os.path.join(path, "validationset"))
    np.ascontiguousarray(test, np.int32).tofile(os.path.join(path, "testingset"))
    if partition is not None:
        np.ascontiguousarray(partition, np.int32).tofile(os.path.join(path, "partition"))


def read_dataset(path, n, e, d, ntrain, nvalid, ntest):
    f = lambda name, dt, cnt: np.fromfile(os.path.join(path, name), dtype=dt, count=cnt)  # noqa: E731
    return dict(indptr=f("edge_src", np.int64, n + 1), indices=f("edge_dst", np.int32, e),
                features=f("features", np.float32, n * d).reshape(n, d), labels=f("labels", np.int32, n),
                train=f("trainingset", np.int32, ntrain), valid=f("validationset", np.int32, nvalid),
                test=f("testingset", np.int32, ntest))


def write_meta_config(cwd, path, batch, n, e, d, ntrain, nvalid, ntest, cache_bytes, epochs, fanout=None):
    if not path.endswith("/"):
        path += "/"
    fields = [path, batch, n, e, d, ntrain, nvalid, ntest, cache_bytes, epochs]
    if fanout:
        fields += list(fanout)
    line = " ".join(str(x) for x in fields)
    with open(os.path.join(cwd, "meta_config"), "w") as f:
        f.write(line)
    return line
