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


def test_oracle_host_generator_is_the_numpy_twin(oracle):
    """oracle/synth_oracle.c (what the CPU arm of bench.py builds its dataset with) == legion_b200/synth.py, bit for bit"""
    oracle.lib().lgo_set_num_threads(4)
    for n, dmin, dmax, seed in ((5000, 6.0, 400, 42), (60000, 25.5, 20000, 0x1E910), (1000, 3.3, 50, 7)):
        ip, ix = synth.graph(n, dmin, dmax, seed)
        ip2, ix2 = oracle.synth_graph(n, dmin, dmax, seed)
        assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2)
    f = synth.features(10, 300, 100, 9)
    assert np.array_equal(f.view(np.uint32), oracle.synth_features(10, 300, 100, 9).view(np.uint32))
    ids = np.array([5, -1, 77, 123456], np.int32)
    r = oracle.synth_feature_rows(ids, 128, 3)
    assert (r[1] == 0).all()
    for k in (0, 2, 3):
        assert np.array_equal(r[k].view(np.uint32), synth.features(int(ids[k]), 1, 128, 3)[0].view(np.uint32))
