"""ctypes binding of the C ABI declared in include/ifb200.h (libifb200.so).

This is the only place Python touches the native library.  There is NO CPU fallback: if the shared
library is missing, or no sm_100 device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libifb200.so")

IFB_OK = 0
COL_MAJOR = 0
ROW_MAJOR = 1

_STATUS_EXC = {1: ValueError, 2: RuntimeError, 3: RuntimeError, 4: MemoryError, 5: RuntimeError, 6: RuntimeError}
SHARD_ALLREDUCE = 0
SHARD_REDUCE_SCATTER = 1


class NativeError(RuntimeError):
    pass


class ForestInfo(C.Structure):
    _fields_ = [
        ("extended", C.c_int32), ("device", C.c_int32), ("num_trees", C.c_int32), ("num_samples", C.c_int32),
        ("total_num_features", C.c_int32), ("max_feature_index", C.c_int32), ("max_depth", C.c_int32),
        ("max_nnz", C.c_int32), ("num_nodes", C.c_int64), ("num_hp_entries", C.c_int64),
        ("device_bytes", C.c_int64),
    ]


class FitParams(C.Structure):
    _fields_ = [
        ("num_estimators", C.c_int32), ("num_samples", C.c_int32), ("num_features", C.c_int32),
        ("bootstrap", C.c_int32), ("random_seed", C.c_int64), ("num_partitions", C.c_int32),
        ("extension_level", C.c_int32), ("tree_begin", C.c_int32), ("tree_end", C.c_int32),
    ]


# every symbol include/ifb200.h declares (tests/test_abi.py checks the exported set against the header)
SYMBOLS = {
    "ifb_abi_version": (C.c_int, []),
    "ifb_last_error": (C.c_char_p, []),
    "ifb_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "ifb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "ifb_host_free": (C.c_int, [C.c_void_p]),
    "ifb_device_alloc": (C.c_int, [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ifb_device_free": (C.c_int, [C.c_int32, C.c_void_p]),
    "ifb_copy_to_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ifb_copy_to_host": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "ifb_forest_create_standard": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 6 + [C.c_int32, C.c_int32,
                                                                                       C.POINTER(C.c_void_p)]),
    "ifb_forest_create_extended": (C.c_int, [C.c_int32, C.c_int32] + [C.c_void_p] * 8 + [C.c_int32, C.c_int32,
                                                                                       C.POINTER(C.c_void_p)]),
    "ifb_forest_destroy": (C.c_int, [C.c_void_p]),
    "ifb_forest_get_info": (C.c_int, [C.c_void_p, C.POINTER(ForestInfo)]),
    "ifb_forest_export": (C.c_int, [C.c_void_p] + [C.c_void_p] * 10),
    "ifb_score_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "ifb_score_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "ifb_score_partial_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "ifb_finalize_scores_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                             C.c_void_p]),
    "ifb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "ifb_comm_init": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ifb_comm_destroy": (C.c_int, [C.c_void_p]),
    "ifb_score_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                    C.c_void_p]),
    "ifb_ipc_export": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p]),
    "ifb_ipc_open": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ifb_ipc_close": (C.c_int, [C.c_int32, C.c_void_p]),
    "ifb_score_scatter_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ifb_finalize_gathered_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p]),
    "ifb_peer_signal_device": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_uint32, C.c_void_p]),
    "ifb_peer_wait_device": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_uint32, C.c_void_p]),
    "ifb_ext_tc_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ifb_std_rank_info": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "ifb_ext_tc_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "ifb_predict_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]),
    "ifb_fit_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                 C.POINTER(FitParams), C.POINTER(C.c_void_p), C.c_void_p]),
    "ifb_fit_host": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                               C.POINTER(FitParams), C.POINTER(C.c_void_p)]),
    "ifb_quantile_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_double, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.c_void_p]),
    "ifb_avg_path_length": (C.c_float, [C.c_int64]),
    "ifb_kernel_launch_count": (C.c_int64, [C.c_int32]),
}

_lib = None


def lib():
    """Load libifb200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != IFB_OK:
        msg = lib().ifb_last_error().decode("utf-8", "replace")
        raise _STATUS_EXC.get(rc, NativeError)(msg)


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def device_count() -> int:
    n = C.c_int32(0)
    rc = lib().ifb_device_count(C.byref(n))
    return int(n.value) if rc == IFB_OK else 0


def kernel_launch_count(reset: bool = False) -> int:
    return int(lib().ifb_kernel_launch_count(1 if reset else 0))


class NativeForest:
    """Owning wrapper of an ifb_forest handle."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @classmethod
    def from_tables(cls, t: dict, device: int = 0) -> "NativeForest":
        """`t` uses the persisted pre-order layout (keys: node_off, left, right, num_instances + feature/threshold or offset/hp_off/hp_idx/hp_w)."""
        out = C.c_void_p()
        c = lambda k, dt: np.ascontiguousarray(t[k], dt)  # noqa: E731
        node_off, left, right = c("node_off", np.int32), c("left", np.int32), c("right", np.int32)
        ninst = c("num_instances", np.int64)
        if t["extended"]:
            off, hp_off = c("offset", np.float64), c("hp_off", np.int64)
            hp_idx, hp_w = c("hp_idx", np.int32), c("hp_w", np.float32)
            check(lib().ifb_forest_create_extended(device, int(t["num_trees"]), _np_ptr(node_off), _np_ptr(left),
                                                   _np_ptr(right), _np_ptr(ninst), _np_ptr(off), _np_ptr(hp_off),
                                                   _np_ptr(hp_idx), _np_ptr(hp_w), int(t["num_samples"]),
                                                   int(t.get("total_num_features", -1)), C.byref(out)))
        else:
            feat, thr = c("feature", np.int32), c("threshold", np.float64)
            check(lib().ifb_forest_create_standard(device, int(t["num_trees"]), _np_ptr(node_off), _np_ptr(left),
                                                   _np_ptr(right), _np_ptr(feat), _np_ptr(thr), _np_ptr(ninst),
                                                   int(t["num_samples"]), int(t.get("total_num_features", -1)),
                                                   C.byref(out)))
        return cls(out.value)

    def close(self):
        if self._h is not None and self._h.value:
            lib().ifb_forest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("forest handle already destroyed")
        return self._h

    def info(self) -> ForestInfo:
        i = ForestInfo()
        check(lib().ifb_forest_get_info(self.handle, C.byref(i)))
        return i

    def export(self) -> dict:
        i = self.info()
        T, n = i.num_trees, i.num_nodes
        out = dict(extended=bool(i.extended), num_trees=T, num_samples=i.num_samples,
                   total_num_features=i.total_num_features, node_off=np.zeros(T + 1, np.int32),
                   left=np.zeros(n, np.int32), right=np.zeros(n, np.int32), num_instances=np.zeros(n, np.int64))
        if i.extended:
            out.update(offset=np.zeros(n, np.float64), hp_off=np.zeros(n + 1, np.int64),
                       hp_idx=np.zeros(i.num_hp_entries, np.int32), hp_w=np.zeros(i.num_hp_entries, np.float32))
        else:
            out.update(feature=np.zeros(n, np.int32), threshold=np.zeros(n, np.float64))
        g = lambda k: _np_ptr(out[k]) if k in out else None  # noqa: E731
        check(lib().ifb_forest_export(self.handle, g("node_off"), g("left"), g("right"), g("feature"), g("threshold"),
                                      g("num_instances"), g("offset"), g("hp_off"), g("hp_idx"), g("hp_w")))
        return out

    # ---- scoring -------------------------------------------------------------------------------
    @staticmethod
    def _layout_of(x_shape, strides_elems):
        """(n, d, ld, layout) of a 2-D (rows x features) view given element strides."""
        n, d = x_shape
        rs, cs = strides_elems
        if n == 0:
            return 0, d, d, ROW_MAJOR
        if cs == 1 and (rs >= d or n <= 1):
            return n, d, max(rs, d) if n > 1 else d, ROW_MAJOR
        if rs == 1 and (cs >= n or d <= 1):
            return n, d, max(cs, n) if d > 1 else n, COL_MAJOR
        raise ValueError("feature matrix must be a row-major or column-major (rows x features) view")

    def score_host(self, X: np.ndarray, want_parts: bool = False):
        """X: (rows x features) float32 numpy view, C order (row-major) or F order (column-major)."""
        assert X.dtype == np.float32 and X.ndim == 2
        n, d, ld, layout = self._layout_of(X.shape, (X.strides[0] // 4, X.strides[1] // 4))
        scores = np.empty(n, np.float64)
        dsum = np.empty(n, np.int32) if want_parts else None
        psum = np.empty(n, np.float32) if want_parts else None
        check(lib().ifb_score_host(self.handle, _np_ptr(X), n, d, ld, layout, _np_ptr(scores), _np_ptr(dsum),
                                   _np_ptr(psum)))
        return (scores, dsum, psum) if want_parts else scores

    def score_device(self, X, scores=None, want_parts: bool = False, stream=None):
        """X: (rows x features) float32 CUDA torch tensor (any of the two dense layouts)."""
        import torch

        assert X.is_cuda and X.dtype == torch.float32 and X.dim() == 2
        n, d, ld, layout = self._layout_of(tuple(X.shape), tuple(X.stride()))
        if scores is None:
            scores = torch.empty(n, dtype=torch.float64, device=X.device)
        dsum = torch.empty(n, dtype=torch.int32, device=X.device) if want_parts else None
        psum = torch.empty(n, dtype=torch.float32, device=X.device) if want_parts else None
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(X.device).cuda_stream)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        check(lib().ifb_score_device(self.handle, p(X), n, d, ld, layout, p(scores), p(dsum), p(psum), st))
        return (scores, dsum, psum) if want_parts else scores

    def std_rank_chunks(self, d: int) -> int:
        """Forest chunks of the rank-word layout for a matrix of d features; 0 = scored by the f32 kernel."""
        n = C.c_int32(0)
        check(lib().ifb_std_rank_info(self.handle, int(d), C.byref(n)))
        return int(n.value)

    def ext_tc_info(self):
        """(padded hyperplane width, accumulator columns) of the tensor-core layout; (0, 0) when the forest has none."""
        kp, nc = C.c_int32(0), C.c_int32(0)
        check(lib().ifb_ext_tc_info(self.handle, C.byref(kp), C.byref(nc)))
        return int(kp.value), int(nc.value)

    def ext_tc_probe(self, X):
        """Diagnostic: scores + raw tensor-core accumulators [rows<=128][columns] + weight slot of every column."""
        import torch

        n, d, ld, layout = self._layout_of(tuple(X.shape), tuple(X.stride()))
        _, nc = self.ext_tc_info()
        rows = min(n, 128)
        scores = torch.empty(rows, dtype=torch.float64, device=X.device)
        acc = torch.zeros((rows, nc), dtype=torch.float32, device=X.device)
        slots = np.empty(nc, np.int32)
        st = C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)
        check(lib().ifb_ext_tc_probe(self.handle, C.c_void_p(X.data_ptr()), n, d, ld, layout,
                                     C.c_void_p(scores.data_ptr()), C.c_void_p(acc.data_ptr()), _np_ptr(slots), st))
        return scores, acc, slots

    def score_partial_device(self, X, path_sum, depth_sum=None, stream=None):
        import torch

        n, d, ld, layout = self._layout_of(tuple(X.shape), tuple(X.stride()))
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(X.device).cuda_stream)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        check(lib().ifb_score_partial_device(self.handle, p(X), n, d, ld, layout, p(path_sum), p(depth_sum), st))


class NativeComm:
    """NCCL communicator owned by libifb200.so (ifb_comm_init): one per GPU process."""

    def __init__(self, device: int, world: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        out = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(lib().ifb_comm_init(device, world, rank, buf, C.byref(out)))
        self._h, self.world, self.rank = out, world, rank

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().ifb_comm_unique_id(buf))
        return bytes(buf.raw)

    def score_sharded(self, forest: "NativeForest", X, total_num_trees: int, mode: int = SHARD_ALLREDUCE, stream=None):
        """Returns (scores, begin, end): all rows (all-reduce) or this rank's slice (reduce-scatter)."""
        import torch

        n, d, ld, layout = NativeForest._layout_of(tuple(X.shape), tuple(X.stride()))
        per = (n + self.world - 1) // self.world
        rows = n if mode == SHARD_ALLREDUCE else max(0, min(n, (self.rank + 1) * per) - min(n, self.rank * per))
        scores = torch.empty(rows, dtype=torch.float64, device=X.device)
        b, e = C.c_int64(0), C.c_int64(0)
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(X.device).cuda_stream)
        check(lib().ifb_score_sharded(forest.handle, self._h, C.c_void_p(X.data_ptr()), n, d, ld, layout, total_num_trees,
                                      mode, C.c_void_p(scores.data_ptr()), C.byref(b), C.byref(e), st))
        return scores, int(b.value), int(e.value)

    def close(self):
        if self._h is not None:
            lib().ifb_comm_destroy(self._h)
            self._h = None


def finalize_scores_device(path_sum, total_num_trees: int, num_samples: int, scores=None, stream=None):
    import torch

    n = path_sum.numel()
    if scores is None:
        scores = torch.empty(n, dtype=torch.float64, device=path_sum.device)
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(path_sum.device).cuda_stream)
    check(lib().ifb_finalize_scores_device(path_sum.device.index or 0, C.c_void_p(path_sum.data_ptr()), n,
                                           total_num_trees, num_samples, C.c_void_p(scores.data_ptr()), st))
    return scores


def predict_device(scores, threshold: float, stream=None):
    import torch

    labels = torch.empty_like(scores)
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(scores.device).cuda_stream)
    check(lib().ifb_predict_device(scores.device.index or 0, C.c_void_p(scores.data_ptr()), scores.numel(),
                                   float(threshold), C.c_void_p(labels.data_ptr()), st))
    return labels


def quantile_device(scores, q: float, stream=None):
    import torch

    v, frac = C.c_double(), C.c_double()
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(scores.device).cuda_stream)
    check(lib().ifb_quantile_device(scores.device.index or 0, C.c_void_p(scores.data_ptr()), scores.numel(), float(q),
                                    C.byref(v), C.byref(frac), st))
    return v.value, frac.value


def fit_device(X, params: FitParams, stream=None) -> NativeForest:
    import torch

    n, d, ld, layout = NativeForest._layout_of(tuple(X.shape), tuple(X.stride()))
    out = C.c_void_p()
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(X.device).cuda_stream)
    check(lib().ifb_fit_device(X.device.index or 0, C.c_void_p(X.data_ptr()), n, d, ld, layout, C.byref(params),
                               C.byref(out), st))
    return NativeForest(out.value)


def fit_host(X: np.ndarray, params: FitParams, device: int = 0) -> NativeForest:
    assert X.dtype == np.float32 and X.ndim == 2
    n, d, ld, layout = NativeForest._layout_of(X.shape, (X.strides[0] // 4, X.strides[1] // 4))
    out = C.c_void_p()
    check(lib().ifb_fit_host(device, _np_ptr(X), n, d, ld, layout, C.byref(params), C.byref(out)))
    return NativeForest(out.value)


class PinnedBuffer:
    """Pinned host allocation from ifb_host_alloc exposed as a numpy array."""

    def __init__(self, shape, dtype=np.float32, order="C"):
        self.shape = tuple(shape)
        dt = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * dt.itemsize
        p = C.c_void_p()
        check(lib().ifb_host_alloc(nbytes, C.byref(p)))
        self._ptr = p
        buf = (C.c_char * nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dt).reshape(self.shape, order=order)

    def free(self):
        if self._ptr is not None:
            self.array = None
            lib().ifb_host_free(self._ptr)
            self._ptr = None
