"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in sorted(f for f in os.listdir(os.path.join(ROOT, "include")) if f.endswith(".h")):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from legion_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 40
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header():
    from legion_b200 import capi
    assert set(_declared()) == set(capi.declared_symbols())
    capi.load()  # sets restype/argtypes for every prototype; raises if one is absent


def test_host_only_entry_points():
    from legion_b200 import capi
    L = capi.load()
    fo = (ctypes.c_int32 * 2)(25, 10)
    assert L.lg_num_ids(8000, fo, 2) == 2_208_000  # B + S1 + S2
    fo3 = (ctypes.c_int32 * 3)(15, 10, 5)
    assert L.lg_num_ids(8000, fo3, 3) == 7_328_000
    assert L.lg_version() >= 100
    # argument validation happens before any CUDA call
    h = ctypes.c_void_p()
    assert L.lg_sampler_create(0, 8000, fo, 9, 1000, ctypes.byref(h)) != 0
    assert b"n_hops" in L.lg_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from legion_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(capi.LegionError, match="no CPU fallback"):
        capi.load()


def test_cost_model_matches_oracle(oracle):
    """lg_cost_model is host arithmetic (cache/cache.cu:445-551): runs without a GPU"""
    import numpy as np
    from conftest import small_graph
    from legion_b200 import capi
    L = capi.load()
    indptr, indices = small_graph(3000, 10.0, 200)
    N = len(indptr) - 1
    rng = np.random.default_rng(2)
    nh = np.sort(rng.zipf(1.5, N).astype(np.uint64))[::-1].copy()
    eh = np.sort(rng.zipf(1.3, N).astype(np.uint64))[::-1].copy()
    order = rng.permutation(N).astype(np.int32)
    for cache_bytes, kg, tt, ft in [(200_000, 1, 0, 10 ** 6), (50_000, 4, 10 ** 5, 10 ** 6), (5_000_000, 2, 0, 10 ** 5),
                                    (123_457, 8, 7 * 10 ** 6, 10 ** 6)]:
        want = oracle.cost_model(nh, eh, order, indptr, 16, cache_bytes, kg, tt, ft)
        a, b, al = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        rc = L.lg_cost_model(nh.ctypes.data, eh.ctypes.data, order.ctypes.data, indptr.ctypes.data, N, 16, cache_bytes,
                             kg, tt, ft, ctypes.byref(a), ctypes.byref(b), ctypes.byref(al))
        assert rc == 0
        assert (a.value, b.value, al.value) == want
    # the saturating variant (DESIGN.md deviation 8): same sweep, but a cache larger than the dataset caches the dataset
    for cache_bytes, kg, tt, ft in [(200_000, 1, 0, 10 ** 6), (50_000, 4, 10 ** 5, 10 ** 6), (10 ** 9, 1, 10 ** 6, 10 ** 6),
                                    (10 ** 9, 4, 10 ** 6, 10 ** 6), (10 ** 11, 8, 0, 10 ** 6)]:
        want = oracle.cost_model(nh, eh, order, indptr, 16, cache_bytes, kg, tt, ft, saturating=True)
        a, b, al = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_double()
        rc = L.lg_cost_model_saturating(nh.ctypes.data, eh.ctypes.data, order.ctypes.data, indptr.ctypes.data, N, 16,
                                        cache_bytes, kg, tt, ft, ctypes.byref(a), ctypes.byref(b), ctypes.byref(al))
        assert rc == 0
        assert (a.value, b.value, al.value) == want
        if cache_bytes >= 10 ** 9:  # everything fits: every vertex is cached (+1 as in the reference, cache.cu:547-548)
            assert a.value == N // kg + 1 and b.value == N // kg + 1, (a.value, b.value)
    ref = oracle.cost_model(nh, eh, order, indptr, 16, 10 ** 9, 1, 10 ** 6, 10 ** 6)
    assert ref[0] == 1 and ref[1] == 1  # the reference's own rule ends with capacity 0 (+1) when everything fits
