"""numpy restatement of the synthetic dataset functions of include/legion_b200_synth.h.

Used on CPU-only boxes (tests, oracle inputs) and to check the device generator bit-for-bit.
Shapes follow legion_server.py:41-88 of the reference; see SHAPES.
"""
import numpy as np

_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_S2 = 0xA5A5A5A55A5A5A5A
_S3 = 0xFEA7FEA7FEA7FEA7
PERM_A = 2654435761
PERM_B = 12345

# name: (num_nodes, num_edges, feature_dim, classes)  — reference dataset table
SHAPES = {
    "products": (2_449_029, 123_718_280, 100, 47),
    "paper100m": (111_059_956, 1_615_685_872, 128, 172),
    "ukunion": (133_633_040, 5_507_679_822, 128, 2),  # BASELINE.json says 128-d; the reference table has 256
    "clueweb": (955_207_488, 42_574_107_469, 128, 2),
}


def mix64(x):
    x = x ^ (x >> np.uint64(30))
    x = x * _C1
    x = x ^ (x >> np.uint64(27))
    x = x * _C2
    return x ^ (x >> np.uint64(31))


def hash2(seed, a):
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) ^ mix64(a.astype(np.uint64) + _GOLD))


def unit(h):
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def dmin_for(num_nodes, num_edges):
    """mean of floor(dmin/sqrt(U)) is about 2*dmin - 0.5"""
    return (num_edges / num_nodes + 0.5) / 2.0


def degrees(n, dmin, dmax, seed):
    v = np.arange(n, dtype=np.uint64)
    x = 1.0 - unit(hash2(seed, v))
    d = (dmin / np.sqrt(x)).astype(np.int64)
    return np.minimum(d, dmax)


def graph(n, dmin, dmax, seed):
    deg = degrees(n, dmin, dmax, seed)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    e = int(indptr[-1])
    v = np.repeat(np.arange(n, dtype=np.uint64), deg)
    k = np.arange(e, dtype=np.uint64) - indptr[:-1].astype(np.uint64)[v.astype(np.int64)]
    with np.errstate(over="ignore"):
        u = unit(hash2(seed ^ _S2, (v << np.uint64(21)) + k))
    t = (u * u) * u
    r = (t * float(n)).astype(np.int64)
    r = np.minimum(r, n - 1).astype(np.uint64)
    with np.errstate(over="ignore"):
        idx = (r * np.uint64(PERM_A) + np.uint64(PERM_B)) % np.uint64(n)
    return indptr, idx.astype(np.int32)


def features(row0, rows, dim, seed):
    i = np.arange(rows * dim, dtype=np.uint64) + np.uint64(row0 * dim)
    bits = (hash2(seed ^ _S3, i) & np.uint64(0xFFFFFFFF)).astype(np.uint32) & np.uint32(0xBFFFFFFF)
    return bits.view(np.float32).reshape(rows, dim)


def labels(n, classes):
    return (np.arange(n, dtype=np.int64) % classes).astype(np.int32)


def split_sets(n, seed, train_frac=0.10, valid=None, test=None):
    """dataset/gen_sets.py:62-76: shuffle all ids, take train / valid / test prefixes (seeded here)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n).astype(np.int32)
    n_train = int(n * train_frac)
    valid = valid if valid is not None else max(1, n // 100)
    test = test if test is not None else max(1, n // 100)
    return perm[:n_train], perm[n_train:n_train + valid], perm[n_train + valid:n_train + valid + test]


def partition_ids(ids, parts):
    """storage/storage_management.cu:171-203 without a partition file: part = id % parts"""
    ids = np.asarray(ids, np.int32)
    return [np.ascontiguousarray(ids[ids % parts == p]) for p in range(parts)]
