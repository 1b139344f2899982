"""ctypes front-end of the CPU oracle (oracle/ifb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libifb_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ifb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libifb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Forest(C.Structure):
    _fields_ = [
        ("num_trees", C.c_int32),
        ("node_off", C.c_void_p),
        ("left", C.c_void_p),
        ("right", C.c_void_p),
        ("num_instances", C.c_void_p),
        ("feature", C.c_void_p),
        ("threshold", C.c_void_p),
        ("offset", C.c_void_p),
        ("hp_off", C.c_void_p),
        ("hp_idx", C.c_void_p),
        ("hp_w", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ifbo_avg_path_length.restype = C.c_float
        L.ifbo_avg_path_length.argtypes = [C.c_int64]
        L.ifbo_path_length.restype = C.c_float
        L.ifbo_path_length.argtypes = [C.POINTER(_Forest), C.c_int, C.c_int, C.c_void_p]
        L.ifbo_score.restype = C.c_int
        L.ifbo_score.argtypes = [C.POINTER(_Forest), C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ifbo_height_limit.restype = C.c_int32
        L.ifbo_height_limit.argtypes = [C.c_int32]
        L.ifbo_fdlibm_log.restype = C.c_double
        L.ifbo_fdlibm_log.argtypes = [C.c_double]
        L.ifbo_jrandom_kat.restype = None
        L.ifbo_jrandom_kat.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ifbo_sample_tree.restype = None
        L.ifbo_sample_tree.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int,
                                       C.c_void_p, C.c_void_p]
        L.ifbo_fit_tree.restype = C.c_int32
        L.ifbo_fit_tree.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32,
                                    C.c_int32, C.c_int32] + [C.c_void_p] * 9 + [C.c_int32]
        L.ifbo_fit_forest.restype = C.c_int
        L.ifbo_fit_forest.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_int32] + [C.c_void_p] * 10 + [C.c_int32, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def avg_path_length(n: int) -> np.float32:
    return np.float32(lib().ifbo_avg_path_length(int(n)))


def height_limit(n: int) -> int:
    return int(lib().ifbo_height_limit(int(n)))


class Forest:
    """Pre-order node tables of one forest (the layout of the reference's Avro rows)."""

    def __init__(self, tables: dict):
        self.t = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in tables.items()}
        self.extended = bool(tables["extended"])
        self.num_trees = int(tables["num_trees"])
        self.num_samples = int(tables["num_samples"])
        t = self.t
        self._c = _Forest(self.num_trees, _p(t["node_off"]), _p(t["left"]), _p(t["right"]),
                          _p(t["num_instances"]), _p(t.get("feature")), _p(t.get("threshold")),
                          _p(t.get("offset")), _p(t.get("hp_off")), _p(t.get("hp_idx")), _p(t.get("hp_w")))

    def path_length(self, tree: int, x) -> np.float32:
        x = np.ascontiguousarray(x, np.float32)
        return np.float32(lib().ifbo_path_length(C.byref(self._c), int(self.extended), tree, _p(x)))

    def score(self, X: np.ndarray, threads: int = 1, want_parts: bool = False):
        """X: 2-D float32 array in any strided layout (rows x features)."""
        assert X.dtype == np.float32 and X.ndim == 2
        n = X.shape[0]
        rs, cs = X.strides[0] // 4, X.strides[1] // 4
        scores = np.empty(n, np.float64)
        dsum = np.empty(n, np.int32) if want_parts else None
        psum = np.empty(n, np.float32) if want_parts else None
        lib().ifbo_score(C.byref(self._c), int(self.extended), _p(X), n, rs, cs, self.num_samples, threads,
                         _p(scores), _p(dsum), _p(psum))
        return (scores, dsum, psum) if want_parts else scores


def fit_tree(data: np.ndarray, seed: int, feature_indices, ext_level: int = -1):
    """IsolationTree.fit / ExtendedIsolationTree.fit on an explicit (n x d) float32 sample."""
    data = np.ascontiguousarray(data, np.float32)
    n, d = data.shape
    feat = np.ascontiguousarray(feature_indices, np.int32)
    hl = height_limit(n)
    ext = ext_level >= 0
    cap = (2 ** (hl + 1) - 1) if ext else max(2 * n - 1, 1)
    k = min(ext_level + 1, len(feat)) if ext else 1
    left = np.empty(cap, np.int32); right = np.empty(cap, np.int32)
    ninst = np.empty(cap, np.int64)
    feature = np.empty(cap, np.int32); thr = np.empty(cap, np.float64)
    off = np.empty(cap, np.float64); hp_len = np.zeros(cap, np.int32)
    hp_idx = np.zeros(cap * k, np.int32); hp_w = np.zeros(cap * k, np.float32)
    nn = lib().ifbo_fit_tree(_p(data), n, d, int(seed), _p(feat), len(feat), int(ext_level), cap, _p(left),
                             _p(right), _p(feature), _p(thr), _p(ninst), _p(off), _p(hp_len), _p(hp_idx),
                             _p(hp_w), k)
    out = dict(left=left[:nn].copy(), right=right[:nn].copy(), num_instances=ninst[:nn].copy())
    if ext:
        out.update(offset=off[:nn].copy(), hp_len=hp_len[:nn].copy(), hp_idx=hp_idx[:nn * k].reshape(nn, k).copy(),
                   hp_w=hp_w[:nn * k].reshape(nn, k).copy(), k=k)
    else:
        out.update(feature=feature[:nn].copy(), threshold=thr[:nn].copy())
    return out


def sample_tree(tree_seed: int, N: int, n: int, d: int, num_features: int, bootstrap: bool = False):
    rows = np.empty(n, np.int64)
    feat = np.empty(num_features, np.int32)
    lib().ifbo_sample_tree(int(tree_seed), int(N), n, d, num_features, int(bootstrap), _p(rows), _p(feat))
    return rows, feat


def fit_forest(X: np.ndarray, num_trees: int, num_samples: int, num_features: int | None = None,
               bootstrap: bool = False, random_seed: int = 1, num_partitions: int = 1, ext_level: int = -1,
               return_samples: bool = False):
    """Whole-forest fit with the engine's sampling contract; returns forest tables (dict) for Forest()."""
    assert X.dtype == np.float32 and X.ndim == 2
    N, d = X.shape
    if num_features is None:
        num_features = d
    ext = ext_level >= 0
    hl = height_limit(num_samples)
    cap = (2 ** (hl + 1) - 1) if ext else 2 * num_samples - 1
    k = min(ext_level + 1, num_features) if ext else 1
    T = num_trees
    n_nodes = np.zeros(T, np.int32)
    left = np.empty(T * cap, np.int32); right = np.empty(T * cap, np.int32)
    ninst = np.empty(T * cap, np.int64)
    feature = thr = off = hp_len = hp_idx = hp_w = None
    if ext:
        off = np.empty(T * cap, np.float64); hp_len = np.zeros(T * cap, np.int32)
        hp_idx = np.zeros(T * cap * k, np.int32); hp_w = np.zeros(T * cap * k, np.float32)
    else:
        feature = np.empty(T * cap, np.int32); thr = np.empty(T * cap, np.float64)
    samples = np.empty((T, num_samples), np.int64) if return_samples else None
    rs, cs = X.strides[0] // 4, X.strides[1] // 4
    lib().ifbo_fit_forest(_p(X), N, d, rs, cs, T, num_samples, num_features, int(bootstrap), int(random_seed),
                          int(num_partitions), int(ext_level), cap, _p(n_nodes), _p(left), _p(right), _p(feature),
                          _p(thr), _p(ninst), _p(off), _p(hp_len), _p(hp_idx), _p(hp_w), k, _p(samples))
    node_off = np.zeros(T + 1, np.int32)
    node_off[1:] = np.cumsum(n_nodes)
    sel = np.concatenate([np.arange(t * cap, t * cap + n_nodes[t]) for t in range(T)])
    tables = dict(extended=ext, num_trees=T, num_samples=num_samples, num_features=num_features,
                  total_num_features=d, threshold_score=-1.0, node_off=node_off, left=left[sel], right=right[sel],
                  num_instances=ninst[sel])
    if ext:
        lens = hp_len[sel].astype(np.int64)
        hp_off = np.zeros(len(sel) + 1, np.int64)
        hp_off[1:] = np.cumsum(lens)
        idx2 = hp_idx.reshape(T * cap, k)[sel]
        w2 = hp_w.reshape(T * cap, k)[sel]
        mask = np.arange(k)[None, :] < lens[:, None]
        tables.update(offset=off[sel], hp_off=hp_off, hp_idx=idx2[mask].astype(np.int32),
                      hp_w=w2[mask].astype(np.float32), ext_level=ext_level)
    else:
        tables.update(feature=feature[sel], threshold=thr[sel])
    if return_samples:
        return tables, samples
    return tables


def jrandom_kat(seed: int):
    ints = np.zeros(3, np.int32); dbl = np.zeros(1, np.float64); g = np.zeros(2, np.float64)
    b = np.zeros(3, np.int32)
    lib().ifbo_jrandom_kat(int(seed), _p(ints), _p(dbl), _p(g), _p(b))
    return ints, float(dbl[0]), g, b


def exact_quantile_threshold(scores: np.ndarray, contamination: float) -> float:
    """Spark approxQuantile(col, [1-contamination], relativeError=0) -> exact order statistic.

    With relativeError 0 the Greenwald-Khanna summary is exact and `query` returns the element of rank
    ceil(q * n) (1-based) of the sorted sample (QuantileSummaries.query: targetRank = ceil(q*n)); used by
    IF/core/SharedTrainLogic.scala:191-198.
    """
    s = np.sort(np.asarray(scores, np.float64))
    q = 1.0 - contamination
    rank = int(math.ceil(q * len(s)))
    rank = min(max(rank, 1), len(s))
    return float(s[rank - 1])
