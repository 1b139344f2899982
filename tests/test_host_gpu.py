"""GPU integration tests of the host mirror, written after the reference's own suites:
IFT/IsolationForestTest.scala, IFT/extended/ExtendedIsolationForestTest.scala,
IFT/IsolationForestModelWriteReadTest.scala, IFT/extended/ExtendedIsolationForestModelWriteReadTest.scala
(IFT = isolation-forest/src/test/scala/com/linkedin/relevance/isolationforest)."""
import numpy as np
import pytest

from test_oracle_golden import _auroc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mammography(golden):
    return golden.mammography["X"].astype(np.float64), golden.mammography["label"]


@pytest.fixture(scope="module")
def shuttle(golden):
    return golden.shuttle["X"].astype(np.float64), golden.shuttle["label"]


def test_isolation_forest_mammography(pkg, mammography):
    # IsolationForestTest.scala:47-88: 100 trees, 256 samples, contamination 0.02, contaminationError 0.02*0.01
    X, y = mammography
    est = (pkg.IsolationForest().setNumEstimators(100).setBootstrap(False).setMaxSamples(256).setMaxFeatures(1.0)
           .setFeaturesCol("features").setPredictionCol("predictedLabel").setScoreCol("outlierScore")
           .setContamination(0.02).setContaminationError(0.02 * 0.01).setRandomSeed(1))
    model = est.fit(X)
    out = model.transform(X)
    assert abs(_auroc(out.outlierScore, y) - 0.86) < 0.02
    # threshold is the exact order statistic; observed contamination within 1 % of the request
    assert abs(out.predictedLabel.mean() - 0.02) < 0.02 * 0.01 + 1.0 / len(X)
    thr = model.getOutlierScoreThreshold()
    assert (out.outlierScore == thr).sum() >= 1
    assert np.array_equal(out.predictedLabel, (out.outlierScore >= thr).astype(float))
    assert model.getNumSamples() == 256 and model.getNumFeatures() == 6 and model.getTotalNumFeatures() == 6


def test_zero_contamination_gives_all_zero_labels(pkg, mammography):
    # IsolationForestTest.scala:132-168
    X, _ = mammography
    model = pkg.IsolationForest().setContamination(0.0).setRandomSeed(1).fit(X)
    out = model.transform(X)
    assert model.getOutlierScoreThreshold() == -1.0 and (out.predictedLabel == 0.0).all()


def test_isolation_forest_shuttle(pkg, shuttle):
    # IsolationForestTest.scala:170-239
    X, y = shuttle
    model = pkg.IsolationForest().setNumEstimators(100).setMaxSamples(256).setContamination(0.07).setRandomSeed(1).fit(X)
    s = model.transform(X).outlierScore
    assert _auroc(s, y) > 0.99
    assert abs(s[y == 1].mean() - 0.61) < 0.02 and abs(s[y == 0].mean() - 0.41) < 0.02


@pytest.mark.parametrize("ext,band", [(5, 0.86), (0, 0.86)])
def test_extended_mammography(pkg, mammography, ext, band):
    # ExtendedIsolationForestTest.scala:15-100
    X, y = mammography
    model = (pkg.ExtendedIsolationForest().setNumEstimators(100).setMaxSamples(256).setContamination(0.02)
             .setContaminationError(0.02 * 0.01).setExtensionLevel(ext).setRandomSeed(1).fit(X))
    out = model.transform(X)
    assert abs(_auroc(out.outlierScore, y) - band) < 0.025
    assert model.getExtensionLevel() == ext


def test_extension_level_resolution(pkg, mammography, tmp_path):
    # ExtendedIsolationForestTest.scala:213-331: levels 1..4 work and persist; the default is resolved per fit
    # (fully extended for the data at hand) and never written back to the estimator
    X, y = mammography
    for lvl in (1, 4):
        m = pkg.ExtendedIsolationForest().setExtensionLevel(lvl).setRandomSeed(1).fit(X)
        assert _auroc(m.transform(X).outlierScore, y) > 0.7 and m.getExtensionLevel() == lvl
        m.write().overwrite().save(tmp_path / "lvl")
        assert pkg.ExtendedIsolationForestModel.load(tmp_path / "lvl").getExtensionLevel() == lvl
    est = pkg.ExtendedIsolationForest().setRandomSeed(1)
    m6 = est.fit(X)
    assert m6.getExtensionLevel() == 5 and not est.isSet("extensionLevel")
    m3 = est.fit(X[:, :3])
    assert m3.getExtensionLevel() == 2 and not est.isSet("extensionLevel")
    t = m6.tables()
    internal = t["left"] != -1
    assert (np.diff(t["hp_off"])[internal] == 6).all()


def test_write_read_preserves_everything(pkg, mammography, tmp_path):
    # IsolationForestModelWriteReadTest.scala:41-110 and the extended twin :76-145
    X, y = mammography
    for est, cls in ((pkg.IsolationForest(), pkg.IsolationForestModel),
                     (pkg.ExtendedIsolationForest().setExtensionLevel(3), pkg.ExtendedIsolationForestModel)):
        m = est.setNumEstimators(50).setContamination(0.02).setRandomSeed(3).fit(X)
        p = tmp_path / cls.__name__
        m.write().overwrite().save(p)
        m2 = cls.load(p)
        assert m2.extractParamMap() == m.extractParamMap()
        assert m2.getOutlierScoreThreshold() == m.getOutlierScoreThreshold()
        a, b = m.transform(X), m2.transform(X)
        assert np.array_equal(a.outlierScore, b.outlierScore) and np.array_equal(a.predictedLabel, b.predictedLabel)
        assert all(m.treeToString(t) == m2.treeToString(t) for t in range(50))


def test_identical_features_make_leaf_roots(pkg, tmp_path):
    # IsolationForestModelWriteReadTest.scala:186-249
    X = np.ones((400, 3))
    m = pkg.IsolationForest().setNumEstimators(10).setMaxSamples(100).setRandomSeed(1).fit(X)
    assert all(m.treeToString(t) == "ExternalNode(numInstances = 100)" for t in range(10))
    m.save(tmp_path / "same")
    m2 = pkg.IsolationForestModel.load(tmp_path / "same")
    assert np.array_equal(m2.transform(X).outlierScore, m.transform(X).outlierScore)


def test_transform_guards(pkg, mammography):
    # IsolationForestModelWriteReadTest.scala:295-376
    X, _ = mammography
    empty = dict(extended=False, num_trees=0, node_off=np.zeros(1, np.int32), left=np.zeros(0, np.int32),
                 right=np.zeros(0, np.int32), feature=np.zeros(0, np.int32), threshold=np.zeros(0),
                 num_instances=np.zeros(0, np.int64))
    with pytest.raises(pkg.IllegalArgumentException, match="Cannot score with an empty IsolationForestModel"):
        pkg.IsolationForestModel.from_tables("u", empty, 256, 2, 2).transform(X[:10, :2])
    leaf = dict(empty, num_trees=1, node_off=np.array([0, 1], np.int32), left=np.array([-1], np.int32),
                right=np.array([-1], np.int32), feature=np.array([-1], np.int32), threshold=np.zeros(1),
                num_instances=np.array([1], np.int64))
    with pytest.raises(pkg.IllegalArgumentException, match="Cannot score with numSamples=1; expected numSamples >= 2"):
        pkg.IsolationForestModel.from_tables("u", leaf, 1, 2, 2).transform(X[:10, :2])
    m = pkg.IsolationForest().setNumEstimators(5).setMaxSamples(64).fit(X[:, :2])
    with pytest.raises(pkg.IllegalArgumentException,
                       match="Input feature vector size 6 did not match the model's training dimension 2"):
        m.transform(X)


def test_float32_and_float64_inputs_agree(pkg, mammography):
    X, _ = mammography
    m = pkg.IsolationForest().setNumEstimators(20).setRandomSeed(5).fit(X)
    assert np.array_equal(m.transform(X).outlierScore, m.transform(X.astype(np.float32)).outlierScore)


def test_fit_is_deterministic_and_matches_oracle(pkg, oracle, mammography):
    """Same seed => same forest; and the forest is the oracle's restatement of the reference builder."""
    X, _ = mammography
    a = pkg.IsolationForest().setNumEstimators(30).setRandomSeed(11).fit(X).tables()
    b = pkg.IsolationForest().setNumEstimators(30).setRandomSeed(11).fit(X).tables()
    ref = oracle.fit_forest(X.astype(np.float32), 30, 256, random_seed=11)
    for k in ("node_off", "left", "right", "feature", "threshold", "num_instances"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], ref[k]), k


def test_cpp_host_program(pkg, tmp_path):
    """The C++ API used directly (no Python in between): tests/cpp/host_roundtrip.cpp."""
    import os
    import subprocess

    from conftest import ROOT

    lib = os.path.join(ROOT, "isolation-forest_b200")
    exe = tmp_path / "host_roundtrip"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_roundtrip.cpp"), "-o", str(exe), "-L" + lib,
                           "-lifb200_host", "-lifb200", "-Wl,-rpath," + lib])
    out = subprocess.run([str(exe), str(tmp_path / "model")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_roundtrip ok" in out.stdout


def test_bootstrap_maxfeatures_and_partitions_reach_the_builder(pkg, oracle, mammography):
    """bootstrap / maxFeatures / the P of the tree-seed formula flow from the Estimator params into the GPU builder:
    forests equal the oracle's restatement with the same resolved values (IF/core/SharedTrainLogic.scala:33-60,
    IF/IsolationForest.scala:76-78)."""
    X, _ = mammography
    X32 = X.astype(np.float32)
    cases = [
        (dict(setBootstrap=True), dict(bootstrap=True)),
        (dict(setMaxFeatures=0.5), dict(num_features=3)),          # floor(0.5 * 6)
        (dict(setMaxFeatures=4.0), dict(num_features=4)),          # > 1.0 is a count
        (dict(setNumPartitions=4), dict(num_partitions=4)),
        (dict(setMaxSamples=0.01), dict()),                         # fraction of N: floor(0.01 * 11183) = 111 samples
    ]
    for setters, okw in cases:
        est = pkg.IsolationForest().setNumEstimators(12).setRandomSeed(9)
        for k, v in setters.items():
            getattr(est, k)(v)
        m = est.fit(X)
        ns = 111 if "setMaxSamples" in setters else 256
        ref = oracle.fit_forest(X32, 12, ns, random_seed=9, **okw)
        got = m.tables()
        for key in ("node_off", "left", "right", "feature", "threshold", "num_instances"):
            assert np.array_equal(got[key], ref[key]), (setters, key)
        assert m.getNumSamples() == ns
        assert m.getNumFeatures() == okw.get("num_features", 6)


def test_sparse_vector_column_scores_like_its_dense_form(pkg, oracle):
    """SparseVector rows (CSR ingest): fit and transform give exactly what the densified column gives."""
    sp = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(5)
    dense = np.where(rng.random((4000, 12)) < 0.3, rng.standard_normal((4000, 12)), 0.0)
    dense[::50] = 0.0                                   # all-zero rows
    X = sp.csr_matrix(dense)
    E = pkg.estimators
    est = E.IsolationForest().setNumEstimators(40).setMaxSamples(128).setContamination(0.05).setRandomSeed(3)
    m_sparse, m_dense = est.fit(X), est.fit(dense)
    ts, td = m_sparse.tables(), m_dense.tables()
    for k in ("node_off", "left", "right", "feature", "threshold", "num_instances"):
        assert np.array_equal(ts[k], td[k]), k
    assert m_sparse.getOutlierScoreThreshold() == m_dense.getOutlierScoreThreshold()
    a, b = m_sparse.transform(X), m_dense.transform(dense)
    assert np.array_equal(a.outlierScore, b.outlierScore) and np.array_equal(a.predictedLabel, b.predictedLabel)
    ref = oracle.Forest(ts).score(dense.astype(np.float32))
    np.testing.assert_allclose(a.outlierScore, ref, rtol=1e-12, atol=0)
    bad = sp.csr_matrix(dense[:10])
    bad.indices[0] = 99
    with pytest.raises(E.IllegalArgumentException, match="sparse index 99"):
        m_sparse.transform(bad)
