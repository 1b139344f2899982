"""CPU tests of the host mirror (libifb200_host.so): params, persistence format, tree rendering.
Mirrors IFT/IsolationForestModelWriteReadTest.scala and IFT/extended/ExtendedIsolationForestModelWriteReadTest.scala
where no GPU is needed (IFT = isolation-forest/src/test/scala/com/linkedin/relevance/isolationforest)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, ROOT

IFR = os.path.join(REFERENCE, "isolation-forest/src/test/resources")
ONNX = os.path.join(REFERENCE, "isolation-forest-onnx/test/resources")
has_ref = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")


def tables_equal(a, b, extended):
    keys = ["node_off", "left", "right", "num_instances"] + (
        ["offset", "hp_off", "hp_idx", "hp_w"] if extended else ["feature", "threshold"])
    for k in keys:
        assert np.array_equal(a[k], b[k]), k


def test_param_defaults_and_validators(pkg):
    est = pkg.IsolationForest("isolation-forest_test")
    # defaults of IF/core/IsolationForestParamsBase.scala:98-109 survive an untouched fit-less round trip
    m = pkg.IsolationForestModel.from_tables("uid_x", _leaf_tables(), 256, 3, 3)
    assert m.extractParamMap() == {"randomSeed": 1, "scoreCol": "outlierScore", "contamination": 0.0, "maxFeatures": 1.0,
                                   "contaminationError": 0.0, "featuresCol": "features", "bootstrap": False,
                                   "predictionCol": "predictedLabel", "numEstimators": 100, "maxSamples": 256.0}
    for setter, bad in (("setNumEstimators", 0), ("setMaxSamples", 0.0), ("setContamination", 0.5),
                        ("setContamination", -0.1), ("setContaminationError", 1.5), ("setMaxFeatures", 0.0),
                        ("setRandomSeed", 0)):
        with pytest.raises(pkg.IllegalArgumentException, match="given invalid value"):
            getattr(est, setter)(bad)
    with pytest.raises(pkg.IllegalArgumentException, match="extensionLevel given invalid value -1"):
        pkg.ExtendedIsolationForest().setExtensionLevel(-1)
    with pytest.raises(pkg.IllegalArgumentException, match="outlierScoreThreshold must be equal to -1"):
        m.setOutlierScoreThreshold(1.5)


def _leaf_tables():
    return dict(extended=False, num_trees=1, node_off=np.array([0, 1], np.int32), left=np.array([-1], np.int32),
                right=np.array([-1], np.int32), feature=np.array([-1], np.int32), threshold=np.zeros(1),
                num_instances=np.array([256], np.int64))


def test_model_constructor_requires(pkg):
    t = _leaf_tables()
    with pytest.raises(pkg.IllegalArgumentException, match="parameter numSamples must be >0"):
        pkg.IsolationForestModel.from_tables("u", t, 0, 3, 3)
    with pytest.raises(pkg.IllegalArgumentException, match="parameter numFeatures must be >0"):
        pkg.IsolationForestModel.from_tables("u", t, 256, 0, 3)
    with pytest.raises(pkg.IllegalArgumentException, match="numFeatures must be <= totalNumFeatures"):
        pkg.IsolationForestModel.from_tables("u", t, 256, 4, 3)
    # legacy 4-argument constructor: totalNumFeatures unknown (IFT/IsolationForestModelWriteReadTest.scala:378-389)
    m = pkg.IsolationForestModel.from_tables("u", t, 256, 3)
    assert m.getTotalNumFeatures() == pkg.IsolationForestModel.UnknownTotalNumFeatures


@has_ref
@pytest.mark.parametrize("path,cls,golden_name", [
    (os.path.join(IFR, "savedIsolationForestModel"), "IsolationForestModel", "std_mammography_spark23"),
    (os.path.join(IFR, "savedExtendedIsolationForestModel"), "ExtendedIsolationForestModel", "ext_mammography"),
    (os.path.join(ONNX, "savedIsolationForestModel/mammographyModel"), "IsolationForestModel", "std_mammography_onnx"),
    (os.path.join(ONNX, "savedIsolationForestModel/shuttleModel"), "IsolationForestModel", "std_shuttle_onnx"),
])
def test_load_reference_saved_models(pkg, golden, path, cls, golden_name):
    """The native reader (snappy + deflate codecs) against the independent Python reader's tables."""
    m = getattr(pkg, cls).load(path)
    g = golden.model(golden_name)
    tables_equal(m.tables(), g, g["extended"])
    meta = g["metadata"]
    assert m.uid == meta["uid"] and m.getNumSamples() == meta["numSamples"] and m.getNumFeatures() == meta["numFeatures"]
    assert m.getOutlierScoreThreshold() == meta["outlierScoreThreshold"]
    assert m.getTotalNumFeatures() == meta.get("totalNumFeatures", -1)       # legacy metadata loads as -1
    pm = m.extractParamMap()
    for k, v in meta["paramMap"].items():
        assert pm[k] == v, k


@has_ref
def test_tree_text_matches_reference_goldens(pkg):
    """IFT/IsolationForestModelWriteReadTest.scala:391-408 and the extended twin (:513-530): tree 0 of the saved
    model prints EXACTLY as expectedTreeStructure.txt (Java number formatting reproduced)."""
    m = pkg.IsolationForestModel.load(os.path.join(IFR, "savedIsolationForestModel"))
    assert m.treeToString(0) == open(os.path.join(IFR, "expectedTreeStructure.txt")).read().strip()
    e = pkg.ExtendedIsolationForestModel.load(os.path.join(IFR, "savedExtendedIsolationForestModel"))
    assert e.treeToString(0) == open(os.path.join(IFR, "expectedExtendedTreeStructure.txt")).read().strip()


def test_tree_text_hashes_without_reference(pkg, golden):
    want = json.load(open(os.path.join(GOLDEN, "tree_text.json")))
    g = golden.model("std_mammography_spark23")
    m = pkg.IsolationForestModel.from_tables("u", g, 256, 6, 6)
    s = m.treeToString(0)
    assert hashlib.sha256(s.encode()).hexdigest() == want["expectedTreeStructure.txt"]["sha256"]
    g = golden.model("ext_mammography")
    e = pkg.ExtendedIsolationForestModel.from_tables("u", g, 256, 6, 6)
    assert hashlib.sha256(e.treeToString(0).encode()).hexdigest() == want["expectedExtendedTreeStructure.txt"]["sha256"]


@pytest.mark.parametrize("name,cls", [("std_shuttle_onnx", "IsolationForestModel"),
                                      ("ext_mammography", "ExtendedIsolationForestModel")])
def test_save_load_roundtrip_and_independent_reader(pkg, golden, tmp_path, name, cls):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import avro_min

    g = golden.model(name)
    ext = bool(g["extended"])
    m = getattr(pkg, cls).from_tables("round-trip_uid", g, 256, g["num_features"], max(g["total_num_features"], g["num_features"]))
    m.setContamination(0.07).setOutlierScoreThreshold(0.5650686965484887).setRandomSeed(7)
    if ext:
        m.setExtensionLevel(5)
    p = tmp_path / "model"
    m.save(p)
    assert sorted(os.listdir(p)) == ["data", "metadata"]
    with pytest.raises(RuntimeError, match="already exists"):
        m.save(p)
    m.write().overwrite().save(p)
    # the independent Python reader sees the reference's layout
    meta, recs, codec = avro_min.read_model_dir(str(p))
    assert codec == "deflate"
    assert meta["class"].endswith(cls) and meta["uid"] == "round-trip_uid" and meta["numSamples"] == 256
    assert set(meta) >= {"class", "timestamp", "sparkVersion", "uid", "paramMap", "outlierScoreThreshold", "numSamples",
                         "numFeatures", "totalNumFeatures"}
    tables_equal(avro_min.forest_arrays(meta, recs), g, ext)
    # and the native reader round-trips params, threshold and trees (IFT/...WriteReadTest.scala:41-110)
    m2 = getattr(pkg, cls).load(p)
    tables_equal(m2.tables(), g, ext)
    assert m2.extractParamMap() == m.extractParamMap()
    assert m2.getOutlierScoreThreshold() == m.getOutlierScoreThreshold()
    assert m2.uid == m.uid and m2.getNumSamples() == 256
    assert all(m2.treeToString(t) == m.treeToString(t) for t in (0, 1, 99))
    # wrong class is rejected like parseMetadata does
    other = "ExtendedIsolationForestModel" if not ext else "IsolationForestModel"
    with pytest.raises(pkg.IllegalArgumentException, match="Expected class .* but found"):
        getattr(pkg, other).load(p)


def test_empty_forest_roundtrip(pkg, tmp_path):
    # IFT/IsolationForestModelWriteReadTest.scala:251-293
    empty = dict(extended=False, num_trees=0, node_off=np.zeros(1, np.int32), left=np.zeros(0, np.int32),
                 right=np.zeros(0, np.int32), feature=np.zeros(0, np.int32), threshold=np.zeros(0),
                 num_instances=np.zeros(0, np.int64))
    m = pkg.IsolationForestModel.from_tables("empty_uid", empty, 256, 2, 2)
    m.save(tmp_path / "e")
    m2 = pkg.IsolationForestModel.load(tmp_path / "e")
    assert m2.numTrees == 0 and m2.getNumSamples() == 256


def test_resolved_params_messages(pkg):
    """validateAndResolveParams messages surface through fit without touching the GPU."""
    X = np.zeros((100, 4))
    with pytest.raises(pkg.IllegalArgumentException, match=r"maxSamples given invalid value 1.5 specifying the use of 1 samples, but >=2"):
        pkg.IsolationForest().setMaxSamples(1.5).fit(X)        # IFT/IsolationForestTest.scala:241-266
    with pytest.raises(pkg.IllegalArgumentException, match=r"maxSamples given invalid value 101.0 specifying the use of 101 samples, but only 100"):
        pkg.IsolationForest().setMaxSamples(101).fit(X)
    with pytest.raises(pkg.IllegalArgumentException, match=r"maxFeatures given invalid value 5.0 specifying the use of 5 features, but only 4"):
        pkg.IsolationForest().setMaxSamples(10).setMaxFeatures(5).fit(X)
    with pytest.raises(pkg.IllegalArgumentException, match=r"extensionLevel given invalid value 4, but must be in \[0, 3\] for a subspace of 4 features"):
        pkg.ExtendedIsolationForest().setMaxSamples(10).setExtensionLevel(4).fit(X)   # extended test :184-211


def test_sparse_column_validation_and_no_cpu_fallback(pkg):
    """SparseVector (CSR) ingest: the SparseVector invariants are checked on the host before any device work, and a
    valid matrix on a box without a GPU fails loudly instead of being scored by some CPU path."""
    sp = pytest.importorskip("scipy.sparse")
    E = pkg.estimators
    t = dict(num_trees=1, node_off=np.array([0, 3], np.int32), left=np.array([1, -1, -1], np.int32),
             right=np.array([2, -1, -1], np.int32), feature=np.array([0, -1, -1], np.int32),
             threshold=np.array([0.5, 0.0, 0.0]), num_instances=np.array([-1, 1, 1], np.int64))
    m = E.IsolationForestModel.from_tables("uid", t, 2, 3, 3)
    X = sp.csr_matrix(np.eye(3))
    X.indices[0] = 7
    with pytest.raises(E.IllegalArgumentException, match=r"sparse index 7 outside \[0, 3\)"):
        m.transform(X)
    X = sp.csr_matrix(np.array([[1.0, 2.0, 0.0], [0.0, 0.0, 3.0]]))
    X.indices[:2] = [1, 0]
    with pytest.raises(E.IllegalArgumentException, match="strictly increasing"):
        m.transform(X)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device available"):
            m.transform(sp.csr_matrix(np.eye(3)))


def test_estimator_write_read(pkg, tmp_path):
    """isolationForestEstimatorWriteReadTest (IFT/IsolationForestTest.scala:16-45) and its extended twin: the estimator's
    params survive write.overwrite().save / load; set and default params stay told apart (DefaultParamsWriter layout)."""
    E = pkg.estimators
    contamination = 0.02
    est1 = (E.IsolationForest().setNumEstimators(200).setBootstrap(True).setMaxSamples(10000).setMaxFeatures(0.7)
            .setFeaturesCol("featuresTestColumn").setPredictionCol("predictedLabelTestColumn")
            .setScoreCol("outlierScoreTestColumn").setContamination(contamination)
            .setContaminationError(contamination * 0.01).setRandomSeed(1))
    path = tmp_path / "isolationForestEstimatorWriteReadTest"
    est1.write().overwrite().save(path)
    with pytest.raises(RuntimeError, match="already exists"):
        est1.save(path)
    est2 = E.IsolationForest.load(path)
    assert est1.extractParamMap() == est2.extractParamMap()
    assert est2.uid == est1.uid and est2.isSet("maxFeatures") and est2.isSet("randomSeed")
    meta = json.loads((path / "metadata" / "part-00000").read_text())
    assert meta["class"] == "com.linkedin.relevance.isolationforest.IsolationForest"
    assert meta["paramMap"]["maxSamples"] == 10000.0 and meta["paramMap"]["bootstrap"] is True
    assert meta["defaultParamMap"]["numEstimators"] == 100 and "extensionLevel" not in meta["defaultParamMap"]
    assert (path / "metadata" / "_SUCCESS").exists()

    # only what was set lands in paramMap; defaults are restored as defaults
    d = tmp_path / "defaults"
    E.IsolationForest().setNumEstimators(7).save(d)
    meta = json.loads((d / "metadata" / "part-00000").read_text())
    assert meta["paramMap"] == {"numEstimators": 7}
    back = E.IsolationForest.load(d)
    assert back.isSet("numEstimators") and not back.isSet("maxSamples")
    assert back.extractParamMap()["maxSamples"] == 256.0

    assert (est2.getNumEstimators(), est2.getBootstrap(), est2.getMaxSamples(), est2.getMaxFeatures()) == (200, True, 10000.0, 0.7)
    assert (est2.getFeaturesCol(), est2.getPredictionCol(), est2.getScoreCol()) == (
        "featuresTestColumn", "predictedLabelTestColumn", "outlierScoreTestColumn")
    assert est2.getContamination() == contamination and est2.getRandomSeed() == 1
    with pytest.raises(E.IllegalStateException, match="extensionLevel"):
        E.ExtendedIsolationForest().getExtensionLevel()

    # extended estimator: extensionLevel travels only when set; class names are checked on load
    x = tmp_path / "ext"
    ex1 = E.ExtendedIsolationForest().setExtensionLevel(3).setNumEstimators(50)
    ex1.save(x)
    ex2 = E.ExtendedIsolationForest.load(x)
    assert ex2.extractParamMap() == ex1.extractParamMap() and ex2.extractParamMap()["extensionLevel"] == 3
    y = tmp_path / "ext_default"
    E.ExtendedIsolationForest().save(y)
    assert "extensionLevel" not in E.ExtendedIsolationForest.load(y).extractParamMap()
    with pytest.raises(E.IllegalArgumentException, match="Expected class"):
        E.IsolationForest.load(x)
    with pytest.raises(E.IllegalArgumentException, match="Expected class"):
        E.ExtendedIsolationForestModel.load(x)


def test_cpp_host_params_program(pkg, tmp_path):
    """The C++ classes used directly, CPU-only part: tests/cpp/host_params.cpp (params, estimator persistence,
    resolve messages, empty model, a reference-written model when /root/reference is mounted)."""
    import subprocess

    lib = os.path.join(ROOT, "isolation-forest_b200")
    exe = tmp_path / "host_params"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_params.cpp"), "-o", str(exe), "-L" + lib,
                           "-lifb200_host", "-lifb200", "-Wl,-rpath," + lib])
    args = [str(exe), str(tmp_path / "out")]
    ref_model = os.path.join(IFR, "savedIsolationForestModel")
    if os.path.isdir(ref_model):
        args.append(ref_model)
    out = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_params ok" in out.stdout


def test_malformed_metadata_is_an_error_not_a_crash(pkg, tmp_path):
    E = pkg.estimators
    for body in ('{"class": "com.linkedin.relevance.isolationforest.IsolationForest"}', '{"uid": "x"}', "not json", ""):
        p = tmp_path / f"bad{abs(hash(body))}"
        (p / "metadata").mkdir(parents=True)
        (p / "metadata" / "part-00000").write_text(body + "\n")
        with pytest.raises((E.IllegalArgumentException, RuntimeError)):
            E.IsolationForest.load(p)
        with pytest.raises((E.IllegalArgumentException, RuntimeError)):
            E.IsolationForestModel.load(p)
    with pytest.raises(RuntimeError, match="does not exist"):
        E.IsolationForest.load(tmp_path / "nowhere")
