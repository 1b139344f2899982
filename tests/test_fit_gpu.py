"""GPU parity tests of the fit hot path: the CUDA tree builder vs the oracle's restatement of
IsolationTree.fit / ExtendedIsolationTree.fit (bit-identical node tables for identical seeds), plus the
reference's own statistical acceptance bands."""
import numpy as np
import pytest

from conftest import synth_mixture
from test_oracle_golden import _auroc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def fit_gpu(nat, X, T, n, num_features=None, bootstrap=False, seed=1, parts=1, ext=-1, tree_range=(0, 0), colmajor=True):
    d = X.shape[1]
    prm = nat.FitParams(T, n, num_features or d, int(bootstrap), seed, parts, ext, tree_range[0], tree_range[1])
    Xd = torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t() if colmajor else torch.from_numpy(X).cuda()
    return nat.fit_device(Xd, prm)


def assert_tables_equal(got, ref):
    keys = ["node_off", "left", "right", "num_instances"]
    keys += ["offset", "hp_off", "hp_idx", "hp_w"] if ref["extended"] else ["feature", "threshold"]
    for k in keys:
        assert np.array_equal(got[k], ref[k]), f"{k} differs"


@pytest.mark.parametrize("n_rows,d,T,n,nf,boot,seed,parts", [
    (5000, 10, 100, 256, None, False, 1, 1),
    (20000, 32, 64, 256, None, False, 7, 4),
    (3000, 9, 20, 100, 4, False, 3, 1),        # maxFeatures < 1, non power-of-two samples
    (600, 5, 30, 256, None, True, 5, 2),       # bootstrap
    (40000, 128, 16, 256, 64, False, 11, 1),
    (2000, 3, 8, 1000, None, False, 2, 1),     # more samples than threads
    (300, 4, 5, 2, None, False, 9, 1),         # smallest legal sample
])
def test_standard_fit_bit_identical(nat, oracle, n_rows, d, T, n, nf, boot, seed, parts):
    X = synth_mixture(n_rows, d, 100 + d)
    X[::5, 0] = 1.0                             # repeated values: exercises the constant-feature retry
    ref = oracle.fit_forest(X, T, n, num_features=nf, bootstrap=boot, random_seed=seed, num_partitions=parts)
    got = fit_gpu(nat, X, T, n, nf, boot, seed, parts).export()
    assert_tables_equal(got, ref)
    got_rm = fit_gpu(nat, X, T, n, nf, boot, seed, parts, colmajor=False).export()
    assert_tables_equal(got_rm, ref)


@pytest.mark.parametrize("n_rows,d,T,n,nf,ext,seed", [
    (5000, 6, 50, 256, None, 5, 1),
    (5000, 6, 50, 256, None, 0, 1),
    (8000, 64, 24, 256, None, 63, 3),
    (3000, 12, 10, 128, 7, 3, 4),
    (1000, 200, 4, 64, None, 199, 5),
])
def test_extended_fit_bit_identical(nat, oracle, n_rows, d, T, n, nf, ext, seed):
    X = synth_mixture(n_rows, d, 300 + d)
    ref = oracle.fit_forest(X, T, n, num_features=nf, random_seed=seed, ext_level=ext)
    got = fit_gpu(nat, X, T, n, nf, seed=seed, ext=ext).export()
    assert_tables_equal(got, ref)


@pytest.mark.parametrize("mode", ["smem", "scratch", "unstaged"])
def test_sample_staging_modes_build_the_same_trees(nat, oracle, mode, monkeypatch):
    """The builder stages each tree's sample in shared memory, in an L2-resident scratch when it does not fit,
    or (IFB_FIT_NO_STAGE test hook / very large samples) reads the training matrix directly: same trees."""
    d = 300 if mode == "scratch" else 48           # 300 x 256 x 4 B = 307 KB does not fit a CTA's shared memory
    if mode == "unstaged":
        monkeypatch.setenv("IFB_FIT_NO_STAGE", "1")
    X = synth_mixture(6000, d, 77)
    X[::7, 1] = -2.0
    for ext in (-1, d - 1, 2):
        ref = oracle.fit_forest(X, 12, 256, random_seed=21, ext_level=ext)
        for colmajor in (True, False):
            got = fit_gpu(nat, X, 12, 256, seed=21, ext=ext, colmajor=colmajor).export()
            assert_tables_equal(got, ref)


def test_identical_rows_and_constant_features(nat, oracle):
    same = np.ones((64, 3), np.float32)
    got = fit_gpu(nat, same, 5, 16, seed=3).export()
    assert (np.diff(got["node_off"]) == 1).all() and (got["num_instances"] == 16).all()
    assert_tables_equal(got, oracle.fit_forest(same, 5, 16, random_seed=3))
    # extended IF does NOT retry: degenerate splits create size-0 leaves (README remark)
    X = synth_mixture(2000, 4, 8)
    X[:, 1] = 2.5
    ref = oracle.fit_forest(X, 20, 256, random_seed=6, ext_level=0)
    got = fit_gpu(nat, X, 20, 256, seed=6, ext=0).export()
    assert_tables_equal(got, ref)
    assert (got["num_instances"] == 0).any()


def test_tree_shards_are_independent_of_the_split(nat, oracle):
    X = synth_mixture(10000, 16, 77)
    whole = fit_gpu(nat, X, 24, 256, seed=13).export()
    a = fit_gpu(nat, X, 24, 256, seed=13, tree_range=(0, 10)).export()
    b = fit_gpu(nat, X, 24, 256, seed=13, tree_range=(10, 24)).export()
    na = a["node_off"][-1]
    for k in ("left", "right", "feature", "threshold", "num_instances"):
        assert np.array_equal(np.concatenate([a[k], b[k]]), whole[k])
    assert np.array_equal(np.concatenate([a["node_off"], b["node_off"][1:] + na]), whole["node_off"])


def test_reference_statistical_bands(nat, oracle, golden):
    """IFT/IsolationForestTest.scala:78-85,211-236; IFT/extended/ExtendedIsolationForestTest.scala:46-53,364-370:
    the acceptance bands the reference's own fit must meet, applied to GPU fit + GPU scoring."""
    Xm, ym = golden.mammography["X"], golden.mammography["label"]
    Xs, ys = golden.shuttle["X"], golden.shuttle["label"]

    def run(X, ext):
        F = fit_gpu(nat, X, 100, 256, seed=1, ext=ext)
        return F.score_device(torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()).cpu().numpy()

    assert abs(_auroc(run(Xm, -1), ym) - 0.86) < 0.02
    s = run(Xs, -1)
    assert _auroc(s, ys) > 0.99
    assert abs(s[ys == 1].mean() - 0.61) < 0.02 and abs(s[ys == 0].mean() - 0.41) < 0.02
    assert abs(_auroc(run(Xm, 5), ym) - 0.86) < 0.025
    assert abs(_auroc(run(Xm, 0), ym) - 0.86) < 0.025
    assert _auroc(run(Xs, 8), ys) > 0.99


def test_fit_argument_errors(nat):
    X = torch.zeros(100, 4, device="cuda")
    with pytest.raises(ValueError, match="but >=2 samples are required"):
        nat.fit_device(X, nat.FitParams(10, 1, 4, 0, 1, 1, -1, 0, 0))
    with pytest.raises(ValueError, match="but only 100 samples are in the input dataset"):
        nat.fit_device(X, nat.FitParams(10, 101, 4, 0, 1, 1, -1, 0, 0))
    with pytest.raises(ValueError, match="but only 4 features are available"):
        nat.fit_device(X, nat.FitParams(10, 50, 5, 0, 1, 1, -1, 0, 0))
    with pytest.raises(ValueError, match=r"extensionLevel given invalid value 4, but must be in \[0, 3\]"):
        nat.fit_device(X, nat.FitParams(10, 50, 4, 0, 1, 1, 4, 0, 0))


def test_degenerate_splits_are_an_error_not_a_memory_fault(nat, oracle):
    """A column holding both -inf and +inf gives NaN / infinite split values: every row goes right and a 0-row left leaf
    appears level after level.  The reference fails the ExternalNode requirement (numInstances > 0, IF/Nodes.scala:27-31);
    the builder must report an error and must not write past the tree's slice of the node tables."""
    X = synth_mixture(4096, 4, 3)
    X[::2, 1] = -np.inf
    X[1::2, 1] = np.inf
    with pytest.raises((ValueError, RuntimeError)):
        fit_gpu(nat, X, 16, 256)
    # the library is still healthy afterwards
    ok = synth_mixture(4096, 4, 3)
    assert_tables_equal(fit_gpu(nat, ok, 8, 256).export(), oracle.fit_forest(ok, 8, 256, random_seed=1))


def test_fit_host_reads_only_the_addressed_extent(nat, oracle):
    """ifb_fit_host on strided views (an F-order slice base[k:k+n], a C-order column slice): only (d-1)*ld + n_rows resp.
    (n_rows-1)*ld + d elements belong to the caller."""
    n, d = 3000, 7
    X = synth_mixture(n, d, 21)
    ref = oracle.fit_forest(X, 10, 256, random_seed=1)
    prm = nat.FitParams(10, 256, d, 0, 1, 1, -1, 0, 0)
    base = np.asfortranarray(np.concatenate([np.full((5, d), np.nan, np.float32), X]))
    view = base[5:5 + n]                       # column-major, ld = n + 5, ends exactly at the allocation's end
    assert_tables_equal(nat.fit_host(view, prm).export(), ref)
    wide = np.full((n, d + 3), np.nan, np.float32)
    wide[:, :d] = X
    assert_tables_equal(nat.fit_host(wide[:, :d], prm).export(), ref)   # row-major, ld = d + 3


@pytest.mark.parametrize("d", [160, 64, 12])
def test_fully_extended_fit_keeps_the_hyperplanes_on_the_device(nat, oracle, monkeypatch, d):
    """k == d: the weights are gathered into the scoring tables on the device and only come back to the host on
    export.  Tables (incl. every weight bit) == oracle == the host-staged build; scoring parity on the same forest, also
    through the CUDA-core kernels (whose per-tree blobs, for d <= 64, are built on first use from the device tables)."""
    n_rows, T = 6000, 7
    X = synth_mixture(n_rows, d, 300 + d)
    ref = oracle.fit_forest(X, T, 256, random_seed=2, ext_level=d - 1)
    F = fit_gpu(nat, X, T, 256, seed=2, ext=d - 1)
    got = F.export()
    assert_tables_equal(got, ref)
    assert F.info().num_hp_entries == len(ref["hp_w"])
    monkeypatch.setenv("IFB_FIT_HOST_HP", "1")
    assert_tables_equal(fit_gpu(nat, X, T, 256, seed=2, ext=d - 1).export(), ref)
    monkeypatch.delenv("IFB_FIT_HOST_HP")
    Xd = torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()
    s, dd, pp = F.score_device(Xd, want_parts=True)
    rs, rd, rp = oracle.Forest(ref).score(X, threads=8, want_parts=True)
    assert np.array_equal(dd.cpu().numpy(), rd) and np.array_equal(pp.cpu().numpy(), rp)
    assert np.max(np.abs(s.cpu().numpy() - rs) / rs) <= 1e-12
    for env in ("IFB_EXT_NO_TC", "IFB_EXT_GENERIC"):        # the CUDA-core kernels read the gathered tables too
        monkeypatch.setenv(env, "1")
        s2, d2, p2 = F.score_device(Xd, want_parts=True)
        assert np.array_equal(d2.cpu().numpy(), rd) and np.array_equal(p2.cpu().numpy(), rp)
        monkeypatch.delenv(env)
