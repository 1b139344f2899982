"""The counter-based generator bench.py and the full-size tests share: numpy and torch produce the same bits."""
import numpy as np
import torch

import synthdata


def test_numpy_and_torch_agree_bit_for_bit():
    for n, d, seed, row0 in ((1000, 10, 1001, 0), (3000, 64, 1003, 12345), (200, 1024, 1005, 99), (5, 1, 3, 2**31)):
        a = synthdata.matrix_numpy(n, d, seed, row0)
        b = synthdata.matrix_torch(torch, n, d, seed, "cpu", row0).contiguous().numpy()
        assert a.dtype == np.float32 and np.array_equal(a, b)
        rows = np.array([row0, row0 + n // 2, row0 + n - 1], np.uint64)
        assert np.array_equal(synthdata.rows_numpy(rows, d, seed), a[[0, n // 2, n - 1]])


def test_mixture_shape():
    x = synthdata.matrix_numpy(200_000, 8, 7)
    core = np.abs(x).max(axis=1) < 6
    assert 0.97 < core.mean() <= 1.0
    assert abs(np.median(x[:, 0]) - 0.5 * 3 / np.sqrt(8)) < 0.3
    assert 0.9 < x[core][:, 0].std() < 1.3
    # streams of different seeds / columns are uncorrelated
    y = synthdata.matrix_numpy(200_000, 8, 8)
    assert abs(np.corrcoef(x[:, 0], y[:, 0])[0, 1]) < 0.02
    assert abs(np.corrcoef(x[:, 0] - x[:, 1], x[:, 2] - x[:, 3])[0, 1]) < 0.02
