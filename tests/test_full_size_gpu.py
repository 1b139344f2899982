"""BASELINE configs 3, 4 (its one-GPU multi-chunk shape) and 5 at the stated size on one B200.

Same style as test_score_gpu.py::test_full_size_properties (config 2): rows come from the counter-based generator
(synthdata.py), so a strided sample of the full matrix is regenerated on the CPU and scored by the oracle against the
same forest (integer depth sums and sequential f32 path sums bit-exact, scores <= 1e-12 relative), plus size-independent
properties: scores in (0, 1), a slice starting in the middle of a tile reproduces the full pass, both layouts agree.
"""
import numpy as np
import pytest
import torch

import synthdata
from test_score_gpu import assert_parity, dev  # noqa: F401

pytestmark = pytest.mark.gpu


def _check_sample(oracle, F, tables, X, seed, d, out, sample_rows):
    n = X.shape[0]
    s, dsum, psum = out
    assert float(s.min()) > 0.0 and float(s.max()) < 1.0
    idx = np.arange(sample_rows, dtype=np.int64) * (n // sample_rows)
    sub = synthdata.rows_numpy(idx.astype(np.uint64), d, seed)
    ti = torch.from_numpy(idx).cuda()
    assert np.array_equal(sub[:256], X[ti[:256]].cpu().numpy()), "CPU and GPU generators disagree"
    ref = oracle.Forest(tables).score(sub, threads=8, want_parts=True)
    assert_parity((s[ti], dsum[ti], psum[ti]), ref)


def test_config3_full_size(nat, oracle, dev):
    """ExtendedIsolationForest (extensionLevel = d - 1) transform 10M x 64, 200 trees: the tcgen05 path."""
    n, d, T, seed = 10_000_000, 64, 200, 1003
    train = synthdata.matrix_torch(torch, 1 << 20, d, 4242, "cuda")
    F = nat.fit_device(train, nat.FitParams(T, 256, d, 0, 1, 1, d - 1, 0, 0))
    del train
    assert F.ext_tc_info()[1] > 0
    tables = F.export()
    X = synthdata.matrix_torch(torch, n, d, seed, "cuda")
    out = F.score_device(X, want_parts=True)
    torch.cuda.synchronize()
    _check_sample(oracle, F, tables, X, seed, d, out, 1 << 14)
    lo, hi = 3_333_331, 3_333_331 + 400_000
    assert torch.equal(F.score_device(X[lo:hi]), out[0][lo:hi])
    assert torch.equal(F.score_device(X[:1_000_000].contiguous()), out[0][:1_000_000])   # row-major input


def test_config5_full_size(nat, oracle, dev):
    """ExtendedIsolationForest 1M x 1024 high-dimensional hyperplanes, 256 trees."""
    n, d, T, seed = 1_000_000, 1024, 256, 1005
    train = synthdata.matrix_torch(torch, 1 << 17, d, 4242, "cuda")
    F = nat.fit_device(train, nat.FitParams(T, 256, d, 0, 1, 1, d - 1, 0, 0))
    del train
    assert F.ext_tc_info()[1] > 0
    tables = F.export()
    X = synthdata.matrix_torch(torch, n, d, seed, "cuda")
    out = F.score_device(X, want_parts=True)
    torch.cuda.synchronize()
    _check_sample(oracle, F, tables, X, seed, d, out, 1 << 12)
    lo, hi = 500_003, 500_003 + 70_000
    assert torch.equal(F.score_device(X[lo:hi]), out[0][lo:hi])
    assert torch.equal(F.score_device(X[:100_000].contiguous()), out[0][:100_000])


def test_config4_one_gpu_shape(nat, oracle, dev):
    """fit + transform, d = 128, 512 trees on ONE GPU: the forest does not fit next to a row tile, so the transform runs
    several forest chunks and carries the f32 path sums between the launches (20M of the 100M rows: same chunking)."""
    n, d, T, seed = 20_000_000, 128, 512, 1004
    X = synthdata.matrix_torch(torch, n, d, seed, "cuda")
    F = nat.fit_device(X, nat.FitParams(T, 256, d, 0, 1, 1, -1, 0, 0))
    tables = F.export()
    out = F.score_device(X, want_parts=True)
    torch.cuda.synchronize()
    _check_sample(oracle, F, tables, X, seed, d, out, 1 << 14)
    assert torch.equal(F.score_device(X), out[0])           # no-depth instantiation, scratch path sums
    lo, hi = 7_000_001, 7_000_001 + 300_000
    assert torch.equal(F.score_device(X[lo:hi]), out[0][lo:hi])
    # the oracle's builder gives the same trees on the rows the GPU builder sampled (fit parity at size)
    t2 = oracle.fit_forest(np.ascontiguousarray(X[: 1 << 16].cpu().numpy()), 8, 256, random_seed=1)
    F2 = nat.fit_device(X[: 1 << 16], nat.FitParams(8, 256, d, 0, 1, 1, -1, 0, 0)).export()
    for k in ("node_off", "left", "right", "feature", "threshold", "num_instances"):
        assert np.array_equal(t2[k], F2[k]), k
