"""The rank-transform design (tools/rank_transform_prototype.py) is bit-exact: walking on integer
ranks reaches the same leaves as the oracle's f32-widened-to-f64 compares, including NaN, +-inf, -0.0 and values
sitting exactly on / one ulp around a threshold."""
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
from rank_transform_prototype import RankedForest, ceil32  # noqa: E402


def test_ceil32_is_the_exact_f32_image_of_a_f64_threshold():
    rng = np.random.default_rng(1)
    t = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-30, 30, 2000), [0.0, -0.0, 1.0, 1e-46, -1e-46, 3.5e38]])
    c = ceil32(t)
    x = np.concatenate([c, np.nextafter(c, np.float32(-np.inf)), np.nextafter(c, np.float32(np.inf))]).astype(np.float32)
    tt = np.concatenate([t, t, t])
    cc = np.concatenate([c, c, c])
    assert np.array_equal(x.astype(np.float64) < tt, x < cc)


def test_rank_walk_reaches_the_oracles_leaves(oracle):
    rng = np.random.default_rng(2)
    d, T = 7, 25
    train = rng.standard_normal((4000, d)).astype(np.float32)
    train[::3, 2] = np.float32(0.5)                      # repeated values => repeated thresholds across trees
    tables = oracle.fit_forest(train, T, 128, random_seed=9)
    rf = RankedForest(tables, d)
    X = rng.standard_normal((3000, d)).astype(np.float32) * 2
    # adversarial rows: exactly on a cut, one ulp either side, specials
    cuts = np.concatenate([c for c in rf.cuts if len(c)])
    k = min(len(cuts), 500)
    X[:k, 0] = cuts[:k]
    X[k:2 * k, 1] = np.nextafter(cuts[:k], np.float32(-np.inf))
    X[2 * k:3 * k, 2] = np.nextafter(cuts[:k], np.float32(np.inf))
    X[-5] = np.nan
    X[-4] = np.inf
    X[-3] = -np.inf
    X[-2] = -0.0
    X[-1, ::2] = np.nan
    depth, leaves = rf.depth_and_leaf(X)
    _, ref_depth, ref_sum = oracle.Forest(tables).score(X, want_parts=True)
    assert np.array_equal(depth, ref_depth)
    # same leaves => the same f32 path lengths: depth (exact in f32) + c(numInstances of the reached leaf), one f32 add
    F = oracle.Forest(tables)
    leaf_n = tables["num_instances"][leaves]
    for i in range(0, len(X), 61):
        for t in range(T):
            want = F.path_length(t, X[i])
            got = np.float32(rf.tree_depth[i, t]) + oracle.avg_path_length(int(leaf_n[i, t]))
            assert got == want, (i, t)
    b = rf.budget()
    assert b["rank_bits"] <= 16 and 0.0 < b["saving"] < 1.0 / 3.0
