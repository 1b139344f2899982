"""Python mirror of the reference's Estimator / Model surface, bound to libifb200_host.so (the C++ host
layer in host/, which in turn drives the CUDA kernels through the C ABI of libifb200.so).

Names, defaults, argument meaning and error behaviour follow the reference
(isolation-forest/src/main/scala/com/linkedin/relevance/isolationforest/):

    IsolationForest().setNumEstimators(100).setMaxSamples(256).setContamination(0.02).fit(X) -> IsolationForestModel
    model.transform(X) -> Scored(outlierScore, predictedLabel)
    model.write().overwrite().save(path);  IsolationForestModel.load(path)
    ExtendedIsolationForest().setExtensionLevel(5) ...

``X`` is the featuresCol content: a (rows x features) array of float64 (Spark Vector values; cast to float like
``.toFloat``) or float32, or a scipy CSR matrix for a column of SparseVectors (absent entries are 0.0).  `require` failures raise IllegalArgumentException (a ValueError).
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libifb200_host.so")


class IllegalArgumentException(ValueError):
    pass


class IllegalStateException(RuntimeError):
    pass


_hlib = None


def hlib():
    global _hlib
    if _hlib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        C.CDLL(os.path.join(_HERE, "libifb200.so"), mode=C.RTLD_GLOBAL)
        L = C.CDLL(HOST_LIB_PATH)
        vp, cp = C.c_void_p, C.c_char_p
        L.ifbh_last_error.restype = cp
        L.ifbh_last_error_kind.restype = C.c_int
        L.ifbh_estimator_create.argtypes = [C.c_int, cp, C.POINTER(vp)]
        L.ifbh_estimator_destroy.argtypes = [vp]
        L.ifbh_estimator_set.argtypes = [vp, cp, cp]
        L.ifbh_estimator_save.argtypes = [vp, cp, C.c_int]
        L.ifbh_estimator_load.argtypes = [C.c_int, cp, C.POINTER(vp)]
        L.ifbh_estimator_describe.argtypes = [vp, vp, C.c_int64]
        L.ifbh_estimator_describe.restype = C.c_int64
        L.ifbh_estimator_fit.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, C.POINTER(vp)]
        L.ifbh_model_create.argtypes = [C.c_int, cp, C.c_int32] + [vp] * 10 + [C.c_int32] * 3 + [C.POINTER(vp)]
        L.ifbh_model_destroy.argtypes = [vp]
        L.ifbh_model_set.argtypes = [vp, cp, cp]
        L.ifbh_model_transform.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
        L.ifbh_estimator_fit_csr.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int32, C.POINTER(vp)]
        L.ifbh_model_transform_csr.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
        L.ifbh_model_save.argtypes = [vp, cp, C.c_int]
        L.ifbh_model_load.argtypes = [C.c_int, cp, C.POINTER(vp)]
        L.ifbh_model_describe.argtypes = [vp, vp, C.c_int64]
        L.ifbh_model_describe.restype = C.c_int64
        L.ifbh_model_tables.argtypes = [vp] * 11
        L.ifbh_model_tree_string.argtypes = [vp, C.c_int32, vp, C.c_int64]
        L.ifbh_model_tree_string.restype = C.c_int64
        _hlib = L
    return _hlib


def _check(rc):
    if rc == 0:
        return
    msg = hlib().ifbh_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise IllegalArgumentException(msg)
    if rc == 2:
        raise IllegalStateException(msg)
    raise RuntimeError(msg)


def _matrix(X):
    X = np.asarray(X)
    if X.ndim != 2:
        raise IllegalArgumentException("features must be a 2-D (rows x features) array")
    if X.dtype == np.float32:
        X = np.ascontiguousarray(X)
        return X, None, C.c_void_p(X.ctypes.data)
    X = np.ascontiguousarray(X, np.float64)
    return X, C.c_void_p(X.ctypes.data), None


def _csr(X):
    """A column of SparseVectors: any object with CSR attributes (scipy.sparse.csr_matrix / csr_array)."""
    if not (hasattr(X, "indptr") and hasattr(X, "indices") and hasattr(X, "data") and hasattr(X, "shape")):
        return None
    indptr = np.ascontiguousarray(X.indptr, np.int64)
    indices = np.ascontiguousarray(X.indices, np.int32)
    values = np.ascontiguousarray(X.data, np.float64)
    n, d = X.shape
    if len(indptr) != n + 1:
        raise IllegalArgumentException("sparse features must be in CSR form (one indptr entry per row + 1)")
    return indptr, indices, values, int(n), int(d)


class Scored:
    """The two columns transform appends: $(scoreCol) and $(predictionCol)."""

    def __init__(self, score, label):
        self.outlierScore = score
        self.predictedLabel = label


class _ParamsMixin:
    _PARAMS = ("numEstimators", "maxSamples", "contamination", "contaminationError", "maxFeatures", "bootstrap",
               "randomSeed", "featuresCol", "predictionCol", "scoreCol")

    def _set(self, name, value):
        raise NotImplementedError

    def setNumEstimators(self, v): return self._set("numEstimators", int(v))
    def setMaxSamples(self, v): return self._set("maxSamples", float(v))
    def setContamination(self, v): return self._set("contamination", float(v))
    def setContaminationError(self, v): return self._set("contaminationError", float(v))
    def setMaxFeatures(self, v): return self._set("maxFeatures", float(v))
    def setBootstrap(self, v): return self._set("bootstrap", bool(v))
    def setRandomSeed(self, v): return self._set("randomSeed", int(v))
    def setFeaturesCol(self, v): return self._set("featuresCol", str(v))
    def setPredictionCol(self, v): return self._set("predictionCol", str(v))
    def setScoreCol(self, v): return self._set("scoreCol", str(v))
    # getters of IsolationForestParamsBase (values come from the native param map: set or default)
    def getNumEstimators(self): return self.extractParamMap()["numEstimators"]
    def getMaxSamples(self): return self.extractParamMap()["maxSamples"]
    def getContamination(self): return self.extractParamMap()["contamination"]
    def getContaminationError(self): return self.extractParamMap()["contaminationError"]
    def getMaxFeatures(self): return self.extractParamMap()["maxFeatures"]
    def getBootstrap(self): return self.extractParamMap()["bootstrap"]
    def getRandomSeed(self): return self.extractParamMap()["randomSeed"]
    def getFeaturesCol(self): return self.extractParamMap()["featuresCol"]
    def getPredictionCol(self): return self.extractParamMap()["predictionCol"]
    def getScoreCol(self): return self.extractParamMap()["scoreCol"]

    def getExtensionLevel(self):
        pm = self.extractParamMap()
        if "extensionLevel" not in pm:   # Params.getOrDefault on a param without default: NoSuchElementException
            raise IllegalStateException("Failed to find a default value for extensionLevel")
        return pm["extensionLevel"]

    # engine parameters without a reference counterpart
    def setDevice(self, v): return self._set("device", int(v))
    def setNumPartitions(self, v): return self._set("numPartitions", int(v))


class _Model(_ParamsMixin):
    _EXTENDED = False

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self._h is not None and self._h.value:
                hlib().ifbh_model_destroy(self._h)
        except Exception:
            pass

    def _set(self, name, value):
        _check(hlib().ifbh_model_set(self._h, name.encode(), json.dumps(value).encode()))
        return self

    def _describe(self):
        n = hlib().ifbh_model_describe(self._h, None, 0)
        buf = C.create_string_buffer(n)
        hlib().ifbh_model_describe(self._h, buf, n)
        return json.loads(buf.value.decode())

    # getters of the reference
    @property
    def uid(self): return self._describe()["uid"]
    def getNumSamples(self): return self._describe()["numSamples"]
    def getNumFeatures(self): return self._describe()["numFeatures"]
    def getTotalNumFeatures(self): return self._describe()["totalNumFeatures"]
    def getOutlierScoreThreshold(self): return self._describe()["outlierScoreThreshold"]
    def setOutlierScoreThreshold(self, v): return self._set("outlierScoreThreshold", float(v))
    def extractParamMap(self): return self._describe()["paramMap"]
    @property
    def numTrees(self): return self._describe()["numTrees"]

    def transform(self, X) -> Scored:
        sp = _csr(X)
        if sp is not None:
            indptr, indices, values, n, d = sp
            scores = np.empty(n, np.float64)
            labels = np.empty(n, np.float64)
            _check(hlib().ifbh_model_transform_csr(self._h, C.c_void_p(indptr.ctypes.data), C.c_void_p(indices.ctypes.data),
                                                   C.c_void_p(values.ctypes.data), n, d, C.c_void_p(scores.ctypes.data),
                                                   C.c_void_p(labels.ctypes.data)))
            return Scored(scores, labels)
        X, p64, p32 = _matrix(X)
        n, d = X.shape
        scores = np.empty(n, np.float64)
        labels = np.empty(n, np.float64)
        _check(hlib().ifbh_model_transform(self._h, p64, p32, n, d, C.c_void_p(scores.ctypes.data),
                                           C.c_void_p(labels.ctypes.data)))
        return Scored(scores, labels)

    def tables(self) -> dict:
        d = self._describe()
        T, n, h = d["numTrees"], d["numNodes"], d["numHpEntries"]
        t = dict(extended=self._EXTENDED, num_trees=T, num_samples=d["numSamples"],
                 total_num_features=d["totalNumFeatures"], node_off=np.zeros(T + 1, np.int32),
                 left=np.zeros(n, np.int32), right=np.zeros(n, np.int32), num_instances=np.zeros(n, np.int64))
        if self._EXTENDED:
            t.update(offset=np.zeros(n, np.float64), hp_off=np.zeros(n + 1, np.int64), hp_idx=np.zeros(h, np.int32),
                     hp_w=np.zeros(h, np.float32))
        else:
            t.update(feature=np.zeros(n, np.int32), threshold=np.zeros(n, np.float64))
        g = lambda k: C.c_void_p(t[k].ctypes.data) if k in t else None  # noqa: E731
        _check(hlib().ifbh_model_tables(self._h, g("node_off"), g("left"), g("right"), g("feature"), g("threshold"),
                                        g("num_instances"), g("offset"), g("hp_off"), g("hp_idx"), g("hp_w")))
        return t

    def treeToString(self, tree: int) -> str:
        n = hlib().ifbh_model_tree_string(self._h, tree, None, 0)
        if n < 0:
            _check(hlib().ifbh_last_error_kind())
        buf = C.create_string_buffer(n)
        hlib().ifbh_model_tree_string(self._h, tree, buf, n)
        return buf.value.decode()

    # MLWritable
    class _Writer:
        def __init__(self, model):
            self._m, self._ow = model, False

        def overwrite(self):
            self._ow = True
            return self

        def save(self, path):
            _check(hlib().ifbh_model_save(self._m._h, os.fspath(path).encode(), int(self._ow)))

    def write(self):
        return _Model._Writer(self)

    def save(self, path):
        self.write().save(path)

    @classmethod
    def load(cls, path):
        out = C.c_void_p()
        _check(hlib().ifbh_model_load(int(cls._EXTENDED), os.fspath(path).encode(), C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_tables(cls, uid, t, num_samples, num_features, total_num_features=-1):
        """new IsolationForestModel(uid, trees, numSamples, numFeatures[, totalNumFeatures])."""
        out = C.c_void_p()
        c = lambda k, dt: np.ascontiguousarray(t[k], dt) if k in t else None  # noqa: E731
        arrs = [c("node_off", np.int32), c("left", np.int32), c("right", np.int32), c("feature", np.int32),
                c("threshold", np.float64), c("num_instances", np.int64), c("offset", np.float64),
                c("hp_off", np.int64), c("hp_idx", np.int32), c("hp_w", np.float32)]
        ptrs = [None if a is None else C.c_void_p(a.ctypes.data) for a in arrs]
        _check(hlib().ifbh_model_create(int(cls._EXTENDED), uid.encode(), int(t["num_trees"]), *ptrs, int(num_samples),
                                        int(num_features), int(total_num_features), C.byref(out)))
        return cls(out.value)


class IsolationForestModel(_Model):
    _EXTENDED = False
    UnknownTotalNumFeatures = -1


class ExtendedIsolationForestModel(_Model):
    _EXTENDED = True

    def setExtensionLevel(self, v): return self._set("extensionLevel", int(v))


class _Estimator(_ParamsMixin):
    _EXTENDED = False
    _MODEL = IsolationForestModel

    def __init__(self, uid: str | None = None, _handle=None):
        if _handle is not None:
            self._h = C.c_void_p(_handle)
            return
        h = C.c_void_p()
        _check(hlib().ifbh_estimator_create(int(self._EXTENDED), uid.encode() if uid else None, C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            hlib().ifbh_estimator_destroy(self._h)
        except Exception:
            pass

    def _set(self, name, value):
        _check(hlib().ifbh_estimator_set(self._h, name.encode(), json.dumps(value).encode()))
        return self

    def _describe(self):
        n = hlib().ifbh_estimator_describe(self._h, None, 0)
        if n < 0:
            _check(hlib().ifbh_last_error_kind())
        buf = C.create_string_buffer(n)
        hlib().ifbh_estimator_describe(self._h, buf, n)
        return json.loads(buf.value.decode())

    @property
    def uid(self): return self._describe()["uid"]

    def isSet(self, name):
        return name in self._describe()["set"]

    def extractParamMap(self):
        """Every param that has a value (explicitly set or default), like Params.extractParamMap."""
        return self._describe()["paramMap"]

    # DefaultParamsWritable / DefaultParamsReadable
    class _Writer:
        def __init__(self, est):
            self._e, self._ow = est, False

        def overwrite(self):
            self._ow = True
            return self

        def save(self, path):
            _check(hlib().ifbh_estimator_save(self._e._h, os.fspath(path).encode(), int(self._ow)))

    def write(self):
        return _Estimator._Writer(self)

    def save(self, path):
        self.write().save(path)

    @classmethod
    def load(cls, path):
        out = C.c_void_p()
        _check(hlib().ifbh_estimator_load(int(cls._EXTENDED), os.fspath(path).encode(), C.byref(out)))
        return cls(_handle=out.value)

    def fit(self, X):
        out = C.c_void_p()
        sp = _csr(X)
        if sp is not None:
            indptr, indices, values, n, d = sp
            _check(hlib().ifbh_estimator_fit_csr(self._h, C.c_void_p(indptr.ctypes.data), C.c_void_p(indices.ctypes.data),
                                                 C.c_void_p(values.ctypes.data), n, d, C.byref(out)))
            return self._MODEL(out.value)
        X, p64, p32 = _matrix(X)
        n, d = X.shape
        _check(hlib().ifbh_estimator_fit(self._h, p64, p32, n, d, C.byref(out)))
        return self._MODEL(out.value)


class IsolationForest(_Estimator):
    _EXTENDED = False
    _MODEL = IsolationForestModel


class ExtendedIsolationForest(_Estimator):
    _EXTENDED = True
    _MODEL = ExtendedIsolationForestModel

    def setExtensionLevel(self, v): return self._set("extensionLevel", int(v))
