"""CPU tests: the oracle (oracle/ifb_oracle.c) against every golden vector the reference's tests hold.

Citations: IFT = isolation-forest/src/test/scala/com/linkedin/relevance/isolationforest,
IFR = isolation-forest/src/test/resources, ONNX = isolation-forest-onnx/test (all under the reference).
"""
import os
import re

import numpy as np
import pytest

from conftest import REFERENCE


def f32(x):
    return np.float32(x)


def test_avg_path_length_kats(oracle):
    # IFT/core/UtilsTest.scala:12-16 -- exact f32 equality
    assert oracle.avg_path_length(0) == f32(0.0)
    assert oracle.avg_path_length(1) == f32(0.0)
    assert oracle.avg_path_length(2) == f32(0.15443134)
    assert oracle.avg_path_length(10) == f32(3.7488806)
    assert oracle.avg_path_length(2**63 - 1) == f32(86.49098)


def _three_node_standard():
    # IFT/IsolationTreeTest.scala:27-42: root splits feature 0 at 1.5; leaves of 10 and 20 instances
    return dict(extended=False, num_trees=1, num_samples=256, node_off=np.array([0, 3], np.int32),
                left=np.array([1, -1, -1], np.int32), right=np.array([2, -1, -1], np.int32),
                feature=np.array([0, -1, -1], np.int32), threshold=np.array([1.5, 0, 0], np.float64),
                num_instances=np.array([-1, 10, 20], np.int64))


def test_three_node_tree_path_length(oracle):
    F = oracle.Forest(_three_node_standard())
    assert F.path_length(0, [1.0]) == f32(4.7488804)
    assert F.path_length(0, [2.0]) == f32(6.143309)


def _three_node_extended(left_n=10, right_n=20):
    # IFT/extended/ExtendedIsolationTreeTest.scala:32-49: normal (0.70710678, 0.70710678), offset 2.5
    w = np.float32(0.7071067812)
    return dict(extended=True, num_trees=1, num_samples=256, node_off=np.array([0, 3], np.int32),
                left=np.array([1, -1, -1], np.int32), right=np.array([2, -1, -1], np.int32),
                offset=np.array([2.5, 0, 0], np.float64), num_instances=np.array([-1, left_n, right_n], np.int64),
                hp_off=np.array([0, 2, 2, 2], np.int64), hp_idx=np.array([0, 1], np.int32),
                hp_w=np.array([w, w], np.float32))


def test_three_node_extended_path_length(oracle):
    F = oracle.Forest(_three_node_extended())
    # dot([1,2]) = 2.12 < 2.5 -> left (10 instances); dot([2,3]) = 3.54 -> right (20)
    assert F.path_length(0, [1.0, 2.0]) == f32(4.7488804)
    assert F.path_length(0, [2.0, 3.0]) == f32(6.143309)


def test_zero_size_leaf_contributes_nothing(oracle):
    # IFT/extended/ExtendedIsolationTreeTest.scala:51-82
    t = _three_node_extended(left_n=0, right_n=5)
    t.update(offset=np.array([0.5, 0, 0], np.float64), hp_off=np.array([0, 1, 1, 1], np.int64),
             hp_idx=np.array([0], np.int32), hp_w=np.array([1.0], np.float32))
    F = oracle.Forest(t)
    assert F.path_length(0, [0.0, 1.0]) == f32(1.0)
    assert F.path_length(0, [1.0, 1.0]) > f32(1.0)


def test_java_random_known_answers(oracle):
    ints, dbl, gauss, bounded = oracle.jrandom_kat(42)
    assert ints.tolist() == [-1170105035, 234785527, -1360544799]   # new Random(42).nextInt() x3
    assert dbl == 0.7275636800328681                                # new Random(42).nextDouble()
    assert gauss[0] == 1.1419053154730547                           # new Random(42).nextGaussian()
    assert bounded.tolist() == [0, 3, 8]                            # new Random(42).nextInt(10) x3


def test_fdlibm_log_close_to_libm(oracle):
    import math
    rng = np.random.default_rng(3)
    for x in np.concatenate([rng.random(2000), rng.random(200) * 1e-300, 1 + rng.random(200) * 1e-9, [1.0, 0.5, 2.0]]):
        a, b = oracle.lib().ifbo_fdlibm_log(float(x)), math.log(float(x))
        assert a == b or abs(a - b) <= 2 * np.spacing(abs(b)), (x, a, b)


def test_golden_scores_mammography(oracle, golden):
    """ONNX/resources/savedIsolationForestModel/mammographyModel/mammographyOutlierScores.csv: 11,183
    reference-computed scores.  Java's Math.pow and glibc's pow differ by 1 ulp on a handful of rows."""
    F = oracle.Forest(golden.model("std_mammography_onnx"))
    s = F.score(golden.scores["X"], threads=4)
    g = golden.scores["score"]
    assert (s == g).sum() >= 11170
    assert np.max(np.abs(s - g) / g) <= 2.3e-16
    thr = golden.model("std_mammography_onnx")["threshold_score"]
    assert np.array_equal((s >= thr).astype(np.uint8), golden.scores["predicted"])
    assert (s == thr).sum() >= 1
    assert oracle.exact_quantile_threshold(s, 0.0232) == thr      # contaminationError 0 => exact rank


def test_saved_model_thresholds(oracle, golden):
    X = golden.mammography["X"]
    # IFR/savedExtendedIsolationForestModel: contamination 0.0232, contaminationError 0.0 (exact)
    m = golden.model("ext_mammography")
    s = oracle.Forest(m).score(X, threads=4)
    assert (s == m["threshold_score"]).sum() == 1
    assert oracle.exact_quantile_threshold(s, 0.0232) == m["threshold_score"]
    assert abs((s >= m["threshold_score"]).mean() - 0.0232) < 0.0232 * 0.01
    # IFR/savedIsolationForestModel (Spark 2.3 era, approximate quantile): threshold is one of the scores
    m = golden.model("std_mammography_spark23")
    s = oracle.Forest(m).score(X, threads=4)
    assert (s == m["threshold_score"]).sum() >= 1
    assert abs((s >= m["threshold_score"]).mean() - 0.02) < 0.002


def _auroc(scores, labels):
    order = np.argsort(scores, kind="mergesort")
    ranks = np.empty(len(scores))
    ranks[order] = np.arange(1, len(scores) + 1)
    # average ranks of ties
    s_sorted = scores[order]
    i = 0
    while i < len(s_sorted):
        j = i
        while j + 1 < len(s_sorted) and s_sorted[j + 1] == s_sorted[i]:
            j += 1
        if j > i:
            ranks[order[i:j + 1]] = (i + j + 2) / 2.0
        i = j + 1
    pos = labels == 1
    n1, n0 = pos.sum(), (~pos).sum()
    return (ranks[pos].sum() - n1 * (n1 + 1) / 2.0) / (n1 * n0)


def test_saved_models_auroc(oracle, golden):
    # ONNX/test/test_isolation_forest_converter.py:140-159: 0.8596 / 0.9976 (+-2% rel)
    s = oracle.Forest(golden.model("std_mammography_onnx")).score(golden.mammography["X"], threads=4)
    assert abs(_auroc(s, golden.mammography["label"]) - 0.8596) < 0.02 * 0.8596
    s = oracle.Forest(golden.model("std_shuttle_onnx")).score(golden.shuttle["X"], threads=4)
    assert abs(_auroc(s, golden.shuttle["label"]) - 0.9976) < 0.02 * 0.9976


def test_oracle_fit_statistical_bands(oracle, golden):
    """Fit has no stream-level pin in the reference; its tests are bands:
    IFT/IsolationForestTest.scala:78-85 (mammography AUROC 0.86+-0.02), :211-236 (shuttle AUROC > 0.99, mean
    score outliers 0.61+-0.02 / inliers 0.41+-0.02); IFT/extended/ExtendedIsolationForestTest.scala:46-53."""
    Xm, ym = golden.mammography["X"], golden.mammography["label"]
    Xs, ys = golden.shuttle["X"], golden.shuttle["label"]
    s = oracle.Forest(oracle.fit_forest(Xm, 100, 256, random_seed=1)).score(Xm, threads=4)
    assert abs(_auroc(s, ym) - 0.86) < 0.02
    s = oracle.Forest(oracle.fit_forest(Xs, 100, 256, random_seed=1)).score(Xs, threads=4)
    assert _auroc(s, ys) > 0.99
    assert abs(s[ys == 1].mean() - 0.61) < 0.02 and abs(s[ys == 0].mean() - 0.41) < 0.02
    for ext in (0, 5):
        s = oracle.Forest(oracle.fit_forest(Xm, 100, 256, random_seed=1, ext_level=ext)).score(Xm, threads=4)
        assert abs(_auroc(s, ym) - 0.86) < 0.025, ext
    s = oracle.Forest(oracle.fit_forest(Xs, 100, 256, random_seed=1, ext_level=8)).score(Xs, threads=4)
    assert _auroc(s, ys) > 0.99


def test_oracle_fit_structure(oracle, golden):
    # IFT/IsolationTreeTest.scala:11-25: all 49,097 shuttle rows, seed 1 -> depth == heightLimit == 16? (ceil(log2 n))
    Xs = golden.shuttle["X"]
    t = oracle.fit_tree(Xs, 1, np.arange(9))
    depth = np.zeros(len(t["left"]), np.int32)
    for i in range(len(depth)):
        if t["left"][i] != -1:
            depth[t["left"][i]] = depth[i] + 1
            depth[t["right"][i]] = depth[i] + 1
    assert oracle.height_limit(len(Xs)) == 16
    assert depth.max() == 16
    # IFT/extended/ExtendedIsolationTreeTest.scala:147-293: unit normals, min(ext+1, dim) ascending non-zeros
    rng = np.random.default_rng(0)
    data = rng.standard_normal((256, 7)).astype(np.float32)
    feats = np.array([0, 2, 3, 5, 6], np.int32)
    for ext in (0, 2, 4, 9):
        t = oracle.fit_tree(data, 5, feats, ext_level=ext)
        k = min(ext + 1, len(feats))
        internal = t["left"] != -1
        assert (t["hp_len"][internal] == k).all() and (t["hp_len"][~internal] == 0).all()
        w = t["hp_w"][internal].astype(np.float64)
        assert np.allclose(np.sqrt((w * w).sum(1)), 1.0, atol=1e-6)
        idx = t["hp_idx"][internal]
        assert (np.diff(idx, axis=1) > 0).all() and np.isin(idx, feats).all()
    # identical rows => every root is a leaf (IFT/IsolationForestModelWriteReadTest.scala:186-237)
    same = np.ones((64, 3), np.float32)
    tb = oracle.fit_forest(same, 5, 16, random_seed=3)
    assert (np.diff(tb["node_off"]) == 1).all() and (tb["num_instances"] == 16).all()


def test_sampling_contract(oracle):
    rows, feat = oracle.sample_tree(12345, 1000, 256, 10, 4)
    assert len(set(rows.tolist())) == 256 and rows.min() >= 0 and rows.max() < 1000
    assert (np.diff(feat) > 0).all() and feat.min() >= 0 and feat.max() < 10
    rows_b, _ = oracle.sample_tree(12345, 300, 256, 10, 10, bootstrap=True)
    assert len(set(rows_b.tolist())) < 256
    # uniformity: every row is drawn about n/N of the time
    cnt = np.zeros(50)
    for s in range(2000):
        r, _ = oracle.sample_tree(s, 50, 10, 3, 3)
        cnt[r] += 1
    assert abs(cnt / 2000 - 0.2).max() < 0.04


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_avro_reader_matches_reference_tree_text(golden):
    """IFT/IsolationForestModelWriteReadTest.scala:391-408 and
    IFT/extended/ExtendedIsolationForestModelWriteReadTest.scala:513-530: tree 0 of the saved models prints as
    the expected*TreeStructure.txt files.  Compared numerically token by token (Java's number formatting is not
    reproduced)."""
    num = re.compile(r"-?\d+\.\d+(?:E-?\d+)?|-?\d+")

    def render(t, ext):
        base = t["node_off"][0]

        def rec(i):
            g = base + i
            if t["left"][g] == -1:
                name = "ExtendedExternalNode" if ext else "ExternalNode"
                return f"{name}(numInstances = {t['num_instances'][g]})"
            l, r = rec(t["left"][g]), rec(t["right"][g])
            if ext:
                b, e = t["hp_off"][g], t["hp_off"][g + 1]
                idx = ", ".join(str(v) for v in t["hp_idx"][b:e])
                w = ", ".join(repr(float(np.float32(v))) for v in t["hp_w"][b:e])
                return (f"ExtendedInternalNode(splitHyperplane = SplitHyperplane(indices = ({idx}), weights = ({w}), "
                        f"offset = {float(t['offset'][g])!r}), leftChild = ({l}), rightChild = ({r}))")
            return (f"InternalNode(splitAttribute = {t['feature'][g]}, splitValue = {float(t['threshold'][g])!r},"
                    f" leftChild = ({l}), rightChild = ({r}))")

        return rec(0)

    ifr = os.path.join(REFERENCE, "isolation-forest/src/test/resources")
    for name, fn, ext in (("std_mammography_spark23", "expectedTreeStructure.txt", False),
                          ("ext_mammography", "expectedExtendedTreeStructure.txt", True)):
        want = open(os.path.join(ifr, fn)).read().strip()
        got = render(golden.model(name), ext)
        assert num.sub("#", got) == num.sub("#", want)                      # identical structure
        gn, wn = num.findall(got), num.findall(want)
        assert len(gn) == len(wn)
        for a, b in zip(gn, wn):
            if ext and "." in b and "E" not in b and len(b) < 14:           # f32 weights print as floats
                assert np.float32(float(a)) == np.float32(float(b)), (a, b)
            else:
                assert float(a) == float(b), (a, b)
