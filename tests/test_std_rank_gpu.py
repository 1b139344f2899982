"""GPU parity tests of the rank-word standard kernel (csrc/score_std_rank.cu, opt-in with IFB_STD_RANK=1): matrices
of <= 32 features are scored on per-feature ranks; every decision, hence every f32 path sum, must equal the oracle's
and the f32 kernel's."""
import numpy as np
import pytest

from conftest import synth_mixture

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def colmajor_cuda(X):
    return torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()


@pytest.fixture(scope="module")
def dev(nat):
    if nat.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible")
    return 0


@pytest.fixture(autouse=True)
def rank_path_on(monkeypatch):
    monkeypatch.setenv("IFB_STD_RANK", "1")


def rank_sums_and_scores(F, Xd):
    """Path sums (accumulate call without depth output) and scores of the product path for this shape."""
    psum = torch.zeros(Xd.shape[0], dtype=torch.float32, device="cuda")
    F.score_partial_device(Xd, psum)
    return psum.cpu().numpy(), F.score_device(Xd).cpu().numpy()


def check_against_oracle_and_f32_kernel(nat, oracle, monkeypatch, tables, X, expect_chunks=None):
    F = nat.NativeForest.from_tables(tables)
    d = X.shape[1]
    chunks = F.std_rank_chunks(d)
    assert chunks >= 1, "shape was expected to take the rank-word kernel"
    if expect_chunks is not None:
        assert chunks == expect_chunks
    Xd = colmajor_cuda(X)
    ref_s, _, ref_p = oracle.Forest(tables).score(X, threads=8, want_parts=True)
    psum, scores = rank_sums_and_scores(F, Xd)
    assert np.array_equal(psum, ref_p), "sequential f32 path sums must be bit-exact"
    assert np.max(np.abs(scores - ref_s) / ref_s) <= 1e-12
    monkeypatch.setenv("IFB_STD_RANK", "0")               # the f32 kernel of score_std.cu on the same call
    assert F.std_rank_chunks(d) == 0
    psum2, scores2 = rank_sums_and_scores(F, Xd)
    monkeypatch.setenv("IFB_STD_RANK", "1")
    assert np.array_equal(psum, psum2) and np.array_equal(scores, scores2)


@pytest.mark.parametrize("n,d,T,ns", [(50_000, 32, 100, 256), (1_000, 10, 100, 256), (777, 1, 5, 64), (5, 3, 3, 4),
                                      (20_011, 17, 33, 256), (4_096, 32, 7, 128), (100_000, 6, 64, 256),
                                      (513, 2, 9, 32), (30_000, 24, 50, 200)])
def test_rank_kernel_matches_oracle(nat, oracle, dev, monkeypatch, n, d, T, ns):
    X = synth_mixture(n, d, 3000 + d)
    tables = oracle.fit_forest(X, T, min(ns, n), random_seed=1)
    check_against_oracle_and_f32_kernel(nat, oracle, monkeypatch, tables, X)


def test_rank_kernel_special_values(nat, oracle, dev, monkeypatch):
    """NaN goes right everywhere, +-inf and -0.0 compare as in the JVM; rows made of the cut values themselves and
    their float neighbours sit exactly on the rank boundaries."""
    n, d = 6_000, 6
    X = synth_mixture(n, d, 11)
    tables = oracle.fit_forest(X, 40, 256, random_seed=2)
    Xs = X.copy()
    Xs[::7, 1] = np.nan
    Xs[::11, 3] = np.inf
    Xs[::13, 0] = -np.inf
    Xs[::17, 2] = -0.0
    Xs[::19, 4] = np.finfo(np.float32).max
    Xs[::23, 5] = -np.finfo(np.float32).max
    Xs[5] = np.nan
    Xs[6] = np.inf
    Xs[7] = -np.inf
    # rows on / one ulp around the split values of the forest
    internal = tables["left"] != -1
    feat, thr = tables["feature"][internal], tables["threshold"][internal]
    k = 0
    for f_, t_ in list(zip(feat, thr))[:1500]:
        c = np.float32(t_)
        for v in (c, np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))):
            Xs[100 + k % (n - 100), f_] = v
            k += 1
    check_against_oracle_and_f32_kernel(nat, oracle, monkeypatch, tables, Xs)


def test_rank_kernel_unaligned_layouts_and_partial_tiles(nat, oracle, dev):
    """ld not a multiple of 4 / misaligned base => plain loads instead of bulk copies; every tail length of a tile."""
    d = 12
    X = synth_mixture(10_001, d, 5)
    tables = oracle.fit_forest(X, 20, 256, random_seed=9)
    F = nat.NativeForest.from_tables(tables)
    assert F.std_rank_chunks(d) == 1
    ref = oracle.Forest(tables).score(X, threads=4, want_parts=True)
    for n in (10_001, 512, 511, 513, 1024, 1, 1537):
        for ld, off in ((n, 0), (n + 3, 0), (n + 7, 1), (((n + 1023) // 1024) * 1024, 0)):
            buf = torch.zeros(d * ld + 8, dtype=torch.float32, device="cuda")
            view = buf[off:off + d * ld].view(d, ld)[:, :n]
            view.copy_(torch.from_numpy(np.ascontiguousarray(X[:n].T)))
            psum, scores = rank_sums_and_scores(F, view.t())
            assert np.array_equal(psum, ref[2][:n]), (n, ld, off)
            assert np.max(np.abs(scores - ref[0][:n]) / ref[0][:n]) <= 1e-12


def test_rank_kernel_forest_in_several_chunks(nat, oracle, dev, monkeypatch):
    """More trees than one shared-memory chunk holds: the f32 sums are carried between launches in tree order."""
    n, d, T = 20_000, 8, 600
    X = synth_mixture(n, d, 77)
    tables = oracle.fit_forest(X, T, 256, random_seed=3)
    F = nat.NativeForest.from_tables(tables)
    assert F.std_rank_chunks(d) >= 3
    Xd = colmajor_cuda(X)
    ref_s, _, ref_p = oracle.Forest(tables).score(X, threads=8, want_parts=True)
    psum = torch.zeros(n, dtype=torch.float32, device="cuda")
    F.score_partial_device(Xd, psum)
    assert np.array_equal(psum.cpu().numpy(), ref_p)
    # the plain scoring call has no sum buffer to carry: it must still be right (whichever kernel takes it)
    s = F.score_device(Xd).cpu().numpy()
    assert np.max(np.abs(s - ref_s) / ref_s) <= 1e-12


def test_rank_kernel_root_leaves_and_shallow_trees(nat, oracle, dev, monkeypatch):
    """Constant data => every tree is a root leaf; tiny samples => trees of depth 1-3 (fewer levels than the two that
    ride in the kernel parameters)."""
    X = np.ones((3_000, 4), np.float32)
    tables = oracle.fit_forest(X, 12, 256, random_seed=5)
    assert np.all(tables["left"] == -1)
    check_against_oracle_and_f32_kernel(nat, oracle, monkeypatch, tables, X)
    for ns in (2, 3, 5):
        Y = synth_mixture(2_000, 5, 40 + ns)
        check_against_oracle_and_f32_kernel(nat, oracle, monkeypatch, oracle.fit_forest(Y, 25, ns, random_seed=ns), Y)


def test_shapes_outside_the_rank_kernel_fall_back(nat, oracle, dev):
    """Wide matrices, infinite thresholds and trees beyond 511 nodes are scored by the f32 kernel."""
    X = synth_mixture(4_000, 40, 8)
    t40 = oracle.fit_forest(X, 10, 256, random_seed=1)
    assert nat.NativeForest.from_tables(t40).std_rank_chunks(40) == 0
    Y = synth_mixture(20_000, 4, 9)
    big = oracle.fit_forest(Y, 6, 8192, random_seed=1)            # height limit 13: thousands of nodes per tree
    assert np.max(np.diff(big["node_off"])) > 511
    Fb = nat.NativeForest.from_tables(big)
    assert Fb.std_rank_chunks(4) == 0
    ref = oracle.Forest(big).score(Y, threads=4)
    assert np.max(np.abs(Fb.score_device(colmajor_cuda(Y)).cpu().numpy() - ref) / ref) <= 1e-12
    inf_thr = dict(extended=False, num_trees=1, num_samples=256, total_num_features=1,
                   node_off=np.array([0, 3], np.int32), left=np.array([1, -1, -1], np.int32),
                   right=np.array([2, -1, -1], np.int32), feature=np.array([0, -1, -1], np.int32),
                   threshold=np.array([np.inf, 0, 0]), num_instances=np.array([-1, 3, 200], np.int64))
    Fi = nat.NativeForest.from_tables(inf_thr)
    assert Fi.std_rank_chunks(1) == 0
    Z = np.array([[0.0], [np.inf], [np.nan], [-np.inf]], np.float32)
    ref = oracle.Forest(inf_thr).score(Z)
    assert np.max(np.abs(Fi.score_device(colmajor_cuda(Z)).cpu().numpy() - ref) / ref) <= 1e-12


def test_rank_kernel_many_cuts_on_one_feature(nat, oracle, dev, monkeypatch):
    """d = 1: every threshold of the chunk lands on the same feature (thousands of cuts, long grid cells), with
    clustered data so that many cuts share a cell."""
    rng = np.random.default_rng(4)
    X = np.concatenate([rng.normal(0, 1e-3, 3_000), rng.normal(5, 1, 3_000), rng.normal(-1e4, 10, 500)]).astype(np.float32)
    X = X.reshape(-1, 1)
    rng.shuffle(X)
    tables = oracle.fit_forest(X, 120, 256, random_seed=6)
    check_against_oracle_and_f32_kernel(nat, oracle, monkeypatch, tables, X)
