import numpy as np

from legion_b200 import synth


def test_generator_is_deterministic_and_well_formed():
    ip, ix = synth.graph(5000, 6.0, 400, 42)
    ip2, ix2 = synth.graph(5000, 6.0, 400, 42)
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2)
    assert ip[0] == 0 and (np.diff(ip) >= 6).all() and (np.diff(ip) <= 400).all()
    assert ix.min() >= 0 and ix.max() < 5000 and len(ix) == ip[-1]
    ip3, _ = synth.graph(5000, 6.0, 400, 43)
    assert not np.array_equal(ip, ip3)


def test_features_finite_and_row_addressable():
    f = synth.features(0, 64, 100, 9)
    assert np.isfinite(f).all()
    g = synth.features(10, 5, 100, 9)
    assert np.array_equal(f[10:15].view(np.uint32), g.view(np.uint32))


def test_partition_is_id_mod_parts():
    tr, va, te = synth.split_sets(10000, 1)
    assert len(tr) == 1000 and len(set(tr) & set(va)) == 0
    parts = synth.partition_ids(tr, 4)
    assert sum(len(p) for p in parts) == 1000
    for p, ids in enumerate(parts):
        assert (ids % 4 == p).all()
