"""The schedule oracle pinned on the reference's own host code (not gpu): oracle/_ref/libref_host.so is
/root/reference/sampling_server/src/engine/ipc_service.cu compiled in place (oracle/ref_host.cu), so
CUDAIPCEnv::Coordinate / GetMaxStep / GetCurrentMode / GetLocalBatchId / GetCurrentBatchsize
(engine/ipc_service.cu:60-132,213-253) run here and lgo_coordinate / lgo_mode_of must equal them."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref_host.so")

CASES = [
    # train per GPU, valid per GPU, test per GPU, batch, epochs
    ([24001, 25000], [1030, 1000], [600, 512], 8000, 3),
    ([244902], [24490], [24490], 8000, 10),                          # products shape, 1 GPU
    ([1670451] * 7 + [1670440], [166000] * 8, [167000] * 8, 8000, 10),  # UK-Union shape, 8 GPUs
    ([8001, 8001, 16000, 8002], [1, 513, 512, 2], [511, 1, 1, 1], 8000, 1),  # one train step; ragged eval sets
    ([9000], [512], [1024], 100, 2),
    ([300, 290, 310, 305], [40, 50, 45, 47], [33, 31, 38, 30], 64, 4),
]


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference at build time)")
    L = C.CDLL(REF)
    L.ref_env_coordinate.restype = C.c_void_p
    L.ref_env_schedule.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ref_env_free.argtypes = [C.c_void_p]
    return L


def test_shm_struct_size(ref):
    assert ref.ref_shm_struct_size() == 7180  # SURVEY 8b; static_assert'ed in sampling_server/src/ipc_service.cc too


@pytest.mark.parametrize("case", CASES)
def test_coordinate_and_schedule_equal_reference(ref, oracle, case):
    train, valid, test, batch, epoch = case
    P = len(train)
    a = lambda x: np.asarray(x, np.int32)  # noqa: E731
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    tr, va, te = a(train), a(valid), a(test)
    steps, tbs, vbs, sbs, shm = np.zeros(3, np.int32), np.zeros(P, np.int32), np.zeros(P, np.int32), np.zeros(P, np.int32), np.zeros(3, np.int32)
    ms = C.c_int32()
    env = ref.ref_env_coordinate(p(tr), p(va), p(te), P, batch, epoch, p(steps), p(tbs), p(vbs), p(sbs), C.byref(ms), p(shm))
    try:
        o_steps, o_vb, o_tb, o_max = oracle.coordinate(train, valid, test, batch, epoch)
        assert list(o_steps) == list(steps) == list(shm)
        assert list(o_vb) == list(vbs) and list(o_tb) == list(sbs) and list(tbs) == [batch] * P
        assert o_max == ms.value
        m, l = C.c_int32(), C.c_int32()
        for g in range(ms.value + 3):  # a few ids past the end: the reference keeps cycling the test set
            ref.ref_env_schedule(env, g, C.byref(m), C.byref(l))
            assert oracle.mode_of(g, steps, epoch) == (m.value, l.value), g
    finally:
        ref.ref_env_free(env)
