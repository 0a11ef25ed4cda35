import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without a CUDA device skips the gpu tests instead of failing them"""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the -m gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


def small_graph(n=2000, avg=12.0, dmax=300, seed=7):
    from legion_b200 import synth
    return synth.graph(n, (avg + 0.5) / 2.0, dmax, seed)


@pytest.fixture(scope="session")
def graph_small():
    return small_graph()


def make_sets(n, frac=0.5, seed=3):
    rng = np.random.default_rng(seed)
    ids = rng.permutation(n).astype(np.int32)[: int(n * frac)]
    labels = (ids % 7).astype(np.int32)
    return ids, labels
