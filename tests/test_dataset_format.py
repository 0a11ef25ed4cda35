"""The Legion on-disk format written by legion_b200/dataset.py (reference dataset/README.md:3-10; loaders
storage/storage_management.cu:100-232; meta_config: legion_server.py:94-95, storage_management.cu:39-61)."""
import os

import numpy as np

from legion_b200 import dataset, synth


def test_dataset_files_round_trip(tmp_path):
    N, D = 500, 12
    indptr, indices = synth.graph(N, 4.0, 60, 3)
    feat = synth.features(0, N, D, 3)
    labels = synth.labels(N, 5)
    train, valid, test = synth.split_sets(N, 3, train_frac=0.3, valid=40, test=30)
    path = str(tmp_path / "ds") + "/"
    dataset.write_dataset(path, indptr, indices, feat, labels, train, valid, test, partition=np.arange(N) % 4)
    E = len(indices)
    # exact element types and sizes of the reference loaders: int64 offsets, int32 everything else, fp32 features
    sizes = {"edge_src": 8 * (N + 1), "edge_dst": 4 * E, "features": 4 * N * D, "labels": 4 * N,
             "trainingset": 4 * len(train), "validationset": 4 * len(valid), "testingset": 4 * len(test), "partition": 4 * N}
    for name, nbytes in sizes.items():
        assert os.path.getsize(os.path.join(path, name)) == nbytes, name
    got = dataset.read_dataset(path, N, E, D, len(train), len(valid), len(test))
    assert np.array_equal(got["indptr"], indptr) and got["indptr"].dtype == np.int64
    assert np.array_equal(got["indices"], indices) and got["indices"].dtype == np.int32
    assert np.array_equal(got["features"].view(np.uint32), feat.view(np.uint32))
    assert np.array_equal(got["labels"], labels)
    for k, want in (("train", train), ("valid", valid), ("test", test)):
        assert np.array_equal(got[k], want)
    # the three sets are disjoint (dataset/gen_sets.py:62-67 shuffles once and slices)
    assert len(set(train) | set(valid) | set(test)) == len(train) + len(valid) + len(test)


def test_meta_config_line(tmp_path):
    line = dataset.write_meta_config(str(tmp_path), "/data/products", 8000, 2449029, 123718280, 100, 196615, 39323, 2213091,
                                     40_000_000_000, 10)
    assert line == "/data/products/ 8000 2449029 123718280 100 196615 39323 2213091 40000000000 10"
    assert open(tmp_path / "meta_config").read() == line  # one line, ten fields, path with a trailing '/'
    line = dataset.write_meta_config(str(tmp_path), "/d/", 1, 2, 3, 4, 5, 6, 7, 8, 9, fanout=[15, 10, 5])
    assert line.split()[10:] == ["15", "10", "5"]  # extension: trailing integers = fan-out per hop
