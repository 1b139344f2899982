"""GPU parity tests of the scoring hot path: CUDA kernels (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest

from conftest import synth_mixture

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def colmajor_cuda(X):
    """(rows x features) CUDA view whose memory is column-major."""
    return torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()


def assert_parity(got, ref, exact_sum=True, score_rel=1e-12):
    gs, gd, gp = (g.cpu().numpy() if hasattr(g, "cpu") else g for g in got)
    rs, rd, rp = ref
    assert np.array_equal(gd, rd), "integer depth sums must be bit-exact"
    if exact_sum:
        assert np.array_equal(gp, rp), "sequential f32 path sums must be bit-exact"
    # tolerance stated by BASELINE.json: <= 1e-5 relative; exp2 vs pow differ by a few ulp at most
    assert np.max(np.abs(gs - rs) / rs) <= score_rel


@pytest.fixture(scope="module")
def dev(nat):
    if nat.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible")
    return 0


def test_golden_forest_scores_match_reference_csv(nat, oracle, golden, dev):
    m = golden.model("std_mammography_onnx")
    F = nat.NativeForest.from_tables(m)
    X = golden.scores["X"]
    s, d, p = F.score_device(colmajor_cuda(X), want_parts=True)
    ref = oracle.Forest(m).score(X, threads=4, want_parts=True)
    assert_parity((s, d, p), ref)
    g = golden.scores["score"]
    s = s.cpu().numpy()
    assert np.max(np.abs(s - g) / g) <= 1e-15           # vs the reference's own stored scores
    assert np.array_equal((s >= m["threshold_score"]).astype(np.uint8), golden.scores["predicted"])


@pytest.mark.parametrize("name,data", [("std_mammography_spark23", "mammography"), ("std_shuttle_onnx", "shuttle"),
                                       ("ext_mammography", "mammography")])
def test_saved_models(nat, oracle, golden, dev, name, data):
    m = golden.model(name)
    X = getattr(golden, data)["X"]
    F = nat.NativeForest.from_tables(m)
    ref = oracle.Forest(m).score(X, threads=4, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    # row-major device input and both host layouts go through the same kernels
    assert_parity(F.score_device(torch.from_numpy(X).cuda(), want_parts=True), ref)
    assert_parity(F.score_host(np.ascontiguousarray(X), want_parts=True), ref)
    assert_parity(F.score_host(np.asfortranarray(X), want_parts=True), ref)
    info = F.info()
    assert info.num_trees == 100 and info.extended == int(m["extended"]) and info.max_depth == 8


@pytest.mark.parametrize("n,d,T", [(1000, 10, 100), (50_000, 32, 100), (20_000, 64, 37), (8_192, 128, 512),
                                   (3_000, 300, 64), (777, 1, 5), (5, 3, 3), (2_000, 700, 40), (1_500, 1200, 30),
                                   (600, 4000, 12)])
def test_standard_synthetic(nat, oracle, dev, n, d, T):
    X = synth_mixture(n, d, 1000 + d)
    ns = min(256, n)
    tables = oracle.fit_forest(X, T, ns, random_seed=1)
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(X, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    s = F.score_device(colmajor_cuda(X))                 # the instantiation the bench times (no depth output)
    assert np.max(np.abs(s.cpu().numpy() - ref[0]) / ref[0]) <= 1e-12


@pytest.mark.parametrize("n,d,T,ext", [(5_000, 8, 50, 0), (5_000, 8, 50, 3), (20_000, 64, 40, 63), (2_000, 200, 8, 199),
                                        (3_000, 40, 16, 9), (9_000, 16, 33, 15), (7_777, 32, 21, 31), (4_000, 24, 10, 23),
                                        (300_000, 5, 12, 4), (1_000, 2, 7, 1), (3_000, 1024, 6, 1023), (5_000, 132, 10, 131),
                                        (4_097, 68, 9, 67)])
def test_extended_synthetic(nat, oracle, dev, n, d, T, ext):
    X = synth_mixture(n, d, 2000 + d)
    tables = oracle.fit_forest(X, T, 256, random_seed=1, ext_level=ext)
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(X, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    assert_parity(F.score_device(torch.from_numpy(X).cuda(), want_parts=True), ref)


def test_extended_generic_and_dense_kernels_agree(nat, oracle, dev, monkeypatch):
    """The generic kernel (any hyperplane shape) and the dense register kernel implement one contract."""
    n, d = 50_000, 48
    X = synth_mixture(n, d, 31)
    tables = oracle.fit_forest(X, 25, 256, random_seed=3, ext_level=d - 1)
    F = nat.NativeForest.from_tables(tables)
    Xd = colmajor_cuda(X)
    dense = [t.cpu().numpy() for t in F.score_device(Xd, want_parts=True)]
    monkeypatch.setenv("IFB_EXT_GENERIC", "1")
    generic = [t.cpu().numpy() for t in F.score_device(Xd, want_parts=True)]
    for a, b in zip(dense, generic):
        assert np.array_equal(a, b)
    assert_parity(dense, oracle.Forest(tables).score(X, threads=8, want_parts=True))


def test_unaligned_and_padded_layouts(nat, oracle, dev):
    """ld not a multiple of 4 / misaligned base => the non-TMA tile loader; padded ld => TMA with pitch."""
    n, d = 10_001, 12
    X = synth_mixture(n, d, 5)
    tables = oracle.fit_forest(X, 20, 256, random_seed=9)
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(X, threads=4, want_parts=True)
    for ld, off in ((n, 0), (n + 3, 0), (n + 7, 1), (10_240, 0)):
        buf = torch.zeros(d * ld + 8, dtype=torch.float32, device="cuda")
        view = buf[off:off + d * ld].view(d, ld)[:, :n]
        view.copy_(torch.from_numpy(np.ascontiguousarray(X.T)))
        assert_parity(F.score_device(view.t(), want_parts=True), ref)


def test_special_values(nat, oracle, dev):
    """NaN features go right at every split (every comparison is false); +-inf and -0.0 behave as in the JVM."""
    n, d = 4096, 6
    X = synth_mixture(n, d, 11)
    tables = oracle.fit_forest(X, 30, 256, random_seed=2)
    Xs = X.copy()
    Xs[::7, 1] = np.nan
    Xs[::11, 3] = np.inf
    Xs[::13, 0] = -np.inf
    Xs[::17, 2] = -0.0
    Xs[5] = np.nan
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(Xs, threads=4, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)
    te = oracle.fit_forest(X, 10, 256, random_seed=2, ext_level=5)
    Fe = nat.NativeForest.from_tables(te)
    refe = oracle.Forest(te).score(Xs, threads=4, want_parts=True)
    assert_parity(Fe.score_device(colmajor_cuda(Xs), want_parts=True), refe)


def test_extended_dense_special_values_take_the_exact_path(nat, oracle, dev):
    """Fully-extended forests decide most visits from a fast filter (tcgen05 accumulators, or the f32 FMA tier of the
    CUDA-core kernels) that is only valid for 'ordinary' rows (finite, 2^-60 <= |x| <= 2^60); rows with zeros,
    denormals, huge values, infs or NaNs must take the exact path and still match the oracle bit for bit -- on the
    tensor-core path and, with IFB_EXT_NO_TC=1, on the dense register kernel."""
    n, d = 8192, 64
    X = synth_mixture(n, d, 77)
    te = oracle.fit_forest(X, 12, 256, random_seed=8, ext_level=d - 1)
    Xs = X.copy()
    Xs[3::64, 5] = 0.0                      # zero products (+0 / -0)
    Xs[7::64, 9] = -0.0
    Xs[11::64, 1] = np.float32(1e-42)       # denormal feature
    Xs[13::64, 2] = np.float32(3e-38)       # tiny normal: product underflows to a denormal
    Xs[17::64, 3] = np.float32(2e38)        # product overflows to inf
    Xs[19::64, 4] = np.inf
    Xs[23::64, 6] = np.nan
    Xs[29::64, :] = 0.0                     # all-zero rows
    F = nat.NativeForest.from_tables(te)
    with np.errstate(all="ignore"):
        ref = oracle.Forest(te).score(Xs, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)
    # and the all-ordinary matrix (fast path everywhere) as well
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), oracle.Forest(te).score(X, threads=8, want_parts=True))
    import os
    os.environ["IFB_EXT_NO_TC"] = "1"          # the CUDA-core dense kernel on the same inputs
    try:
        assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)
    finally:
        del os.environ["IFB_EXT_NO_TC"]


def test_extended_identity_hyperplanes_narrower_than_the_matrix(nat, oracle, dev):
    """Hyperplanes over the first k of d > k columns (a hand-built forest): columns the model never reads may hold
    Inf / NaN without poisoning the dot product (the reference never touches them)."""
    rng = np.random.default_rng(12)
    n, d, k, T = 4096, 10, 6, 9
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[::3, 7] = np.inf
    X[1::3, 8] = np.nan
    X[2::3, 9] = -np.inf
    w = rng.standard_normal((T, k)).astype(np.float32)
    w /= np.linalg.norm(w, axis=1, keepdims=True).astype(np.float32)
    tables = dict(extended=True, num_trees=T, num_samples=256, total_num_features=d,
                  node_off=np.arange(0, 3 * T + 1, 3, dtype=np.int32), left=np.tile([1, -1, -1], T).astype(np.int32),
                  right=np.tile([2, -1, -1], T).astype(np.int32), num_instances=np.tile([-1, 3, 200], T).astype(np.int64),
                  offset=np.concatenate([[o, 0.0, 0.0] for o in rng.standard_normal(T)]),
                  hp_off=np.concatenate([[0], np.cumsum(np.tile([k, 0, 0], T))]).astype(np.int64),
                  hp_idx=np.tile(np.arange(k, dtype=np.int32), T), hp_w=w.reshape(-1))
    F = nat.NativeForest.from_tables(tables)
    with np.errstate(all="ignore"):
        ref = oracle.Forest(tables).score(X, threads=4, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    assert_parity(F.score_device(torch.from_numpy(X).cuda(), want_parts=True), ref)


def test_extended_wide_kernel_special_values(nat, oracle, dev):
    """k = d > 64 runs the wide kernel (lanes split the terms; re-association guarded by a rounding bound with a
    sequential fallback).  NaN / inf / overflow / exact-zero ties must come out exactly as the sequential oracle."""
    n, d = 2048, 256
    X = synth_mixture(n, d, 91)
    te = oracle.fit_forest(X, 8, 256, random_seed=4, ext_level=d - 1)
    Xs = X.copy()
    Xs[5::32, 7] = np.nan
    Xs[9::32, 100] = np.inf
    Xs[13::32, 200] = -np.inf
    Xs[17::32, :] = 0.0                      # sum == 0 exactly: zero tolerance => sequential path
    Xs[21::32, 3] = np.float32(3e38)         # products overflow
    Xs[25::32, :] *= np.float32(1e-30)       # tiny magnitudes
    F = nat.NativeForest.from_tables(te)
    with np.errstate(all="ignore"):
        ref = oracle.Forest(te).score(Xs, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)
    assert_parity(F.score_device(torch.from_numpy(Xs).cuda(), want_parts=True), ref)


def test_threshold_boundary_values(nat, oracle, dev):
    """Features sitting exactly on / one ulp around a split value (the f32-ceil threshold transform)."""
    thr = np.array([0.1, 1.0 + 2**-30, -3.0000000001, 1e-50, 3.5e38, -3.5e38], np.float64)
    T = len(thr)
    tables = dict(extended=False, num_trees=T, num_samples=256, total_num_features=1,
                  node_off=np.arange(0, 3 * T + 1, 3, dtype=np.int32), left=np.tile([1, -1, -1], T).astype(np.int32),
                  right=np.tile([2, -1, -1], T).astype(np.int32), feature=np.tile([0, -1, -1], T).astype(np.int32),
                  threshold=np.repeat(thr, 3) * np.tile([1, 0, 0], T), num_instances=np.tile([-1, 3, 200], T).astype(np.int64))
    vals = []
    for t in thr:
        f = np.float32(t)
        vals += [f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))]
    vals += [0.0, -0.0, np.inf, -np.inf, np.nan, np.float32(1e-45), np.finfo(np.float32).max]
    X = np.array(vals, np.float32).reshape(-1, 1)
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(X, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)


def test_tree_sharded_partial_sums(nat, oracle, dev):
    """Tree-sharded scoring: two forest shards accumulate into path/depth sums, then one finalize."""
    n, d, T = 30_000, 16, 64
    X = synth_mixture(n, d, 21)
    tables = oracle.fit_forest(X, T, 256, random_seed=4)
    ref = oracle.Forest(tables).score(X, threads=8, want_parts=True)

    def shard(t0, t1):
        nb, ne = tables["node_off"][t0], tables["node_off"][t1]
        s = dict(tables)
        s.update(num_trees=t1 - t0, node_off=(tables["node_off"][t0:t1 + 1] - nb).astype(np.int32))
        for k in ("left", "right", "feature", "threshold", "num_instances"):
            s[k] = tables[k][nb:ne]
        return s

    Xd = colmajor_cuda(X)
    psum = torch.zeros(n, dtype=torch.float32, device="cuda")
    dsum = torch.zeros(n, dtype=torch.int32, device="cuda")
    for t0, t1 in ((0, 40), (40, 64)):
        nat.NativeForest.from_tables(shard(t0, t1)).score_partial_device(Xd, psum, dsum)
    scores = nat.finalize_scores_device(psum, T, 256)
    # one rank after the other on one stream == the sequential order, so even the f32 sums are exact here
    assert_parity((scores, dsum, psum), ref)


def test_predict_and_quantile(nat, oracle, golden, dev):
    m = golden.model("ext_mammography")
    X = golden.mammography["X"]
    F = nat.NativeForest.from_tables(m)
    s = F.score_device(colmajor_cuda(X))
    thr, frac = nat.quantile_device(s, 1 - 0.0232)
    assert thr == m["threshold_score"]                      # the reference's stored exact quantile
    assert thr == oracle.exact_quantile_threshold(s.cpu().numpy(), 0.0232)
    lab = nat.predict_device(s, thr).cpu().numpy()
    assert lab.sum() == round(frac * len(X)) and abs(frac - 0.0232) < 0.0232 * 0.01
    assert nat.predict_device(s, -1.0).sum().item() == 0   # no threshold => all 0.0
    for q in (0.0, 0.5, 1.0, 0.999999):
        v, _ = nat.quantile_device(s, q)
        ss = np.sort(s.cpu().numpy())
        r = min(max(int(np.ceil(q * len(ss))), 1), len(ss))
        assert v == ss[r - 1]


def test_error_messages_match_reference(nat, oracle, golden, dev):
    m = golden.model("ext_mammography")            # totalNumFeatures = 6
    F = nat.NativeForest.from_tables(m)
    X = torch.zeros(10, 2, device="cuda")
    with pytest.raises(ValueError, match="Input feature vector size 2 did not match the model's training dimension 6"):
        F.score_device(X)
    empty = dict(extended=False, num_trees=0, num_samples=256, node_off=np.zeros(1, np.int32),
                 left=np.zeros(0, np.int32), right=np.zeros(0, np.int32), feature=np.zeros(0, np.int32),
                 threshold=np.zeros(0), num_instances=np.zeros(0, np.int64))
    with pytest.raises(ValueError, match="Cannot score with an empty IsolationForestModel"):
        nat.NativeForest.from_tables(empty).score_device(torch.zeros(4, 2, device="cuda"))
    one = dict(_leaf_forest(), num_samples=1)
    with pytest.raises(ValueError, match="Cannot score with numSamples=1; expected numSamples >= 2"):
        nat.NativeForest.from_tables(one).score_device(torch.zeros(4, 2, device="cuda"))
    bad = _leaf_forest()
    bad["num_instances"] = np.array([0], np.int64)      # ExternalNode requires numInstances > 0
    with pytest.raises(ValueError):
        nat.NativeForest.from_tables(bad)
    hp = dict(extended=True, num_trees=1, num_samples=8, node_off=np.array([0, 3], np.int32),
              left=np.array([1, -1, -1], np.int32), right=np.array([2, -1, -1], np.int32),
              offset=np.zeros(3), num_instances=np.array([-1, 0, 4], np.int64), hp_off=np.array([0, 2, 2, 2], np.int64),
              hp_idx=np.array([1, 0], np.int32), hp_w=np.array([0.6, 0.8], np.float32))
    with pytest.raises(ValueError, match="indices must be sorted in ascending order"):
        nat.NativeForest.from_tables(hp)


def _leaf_forest():
    return dict(extended=False, num_trees=1, num_samples=256, node_off=np.array([0, 1], np.int32),
                left=np.array([-1], np.int32), right=np.array([-1], np.int32), feature=np.array([-1], np.int32),
                threshold=np.zeros(1), num_instances=np.array([256], np.int64))


def test_root_leaf_forest_and_export_roundtrip(nat, oracle, golden, dev):
    F = nat.NativeForest.from_tables(_leaf_forest())
    s = F.score_device(torch.randn(100, 3, device="cuda"))
    ref = oracle.Forest(_leaf_forest()).score(np.zeros((100, 3), np.float32))
    assert np.array_equal(s.cpu().numpy(), ref)
    for name in ("std_shuttle_onnx", "ext_mammography"):
        m = golden.model(name)
        e = nat.NativeForest.from_tables(m).export()
        for k in ("node_off", "left", "right", "num_instances") + (
                ("offset", "hp_off", "hp_idx", "hp_w") if m["extended"] else ("feature", "threshold")):
            assert np.array_equal(e[k], m[k]), k


def test_full_size_properties(nat, oracle, dev):
    """BASELINE config 2 at full size (10M x 32, 100 trees): oracle on a strided sample + size-independent
    properties (scores in (0,1), permutation equivariance, chunk/tile independence)."""
    n, d, T = 10_000_000, 32, 100
    g = torch.Generator(device="cuda").manual_seed(1002)
    Xt = torch.randn(d, n, device="cuda", generator=g)           # [d][n] memory == column-major rows x features
    Xt[:, : n // 2] += 3.0 / np.sqrt(d)
    Xt[:, -n // 50:] *= 4.0
    X = Xt.t()
    fit_rows = X[:: n // 4096][:4096].cpu().numpy()
    tables = oracle.fit_forest(np.ascontiguousarray(fit_rows), T, 256, random_seed=1)
    F = nat.NativeForest.from_tables(tables)
    s, dsum, psum = F.score_device(X, want_parts=True)
    torch.cuda.synchronize()
    assert float(s.min()) > 0.0 and float(s.max()) < 1.0
    idx = torch.arange(0, n, 997, device="cuda")
    sub = X[idx].cpu().numpy()
    ref = oracle.Forest(tables).score(np.ascontiguousarray(sub), threads=8, want_parts=True)
    assert_parity((s[idx], dsum[idx], psum[idx]), ref)
    # scoring a slice that starts in the middle of a tile gives the same rows the full pass gave
    lo, hi = 1_234_564, 1_234_564 + 500_000
    s2 = F.score_device(X[lo:hi])
    assert torch.equal(s2, s[lo:hi])
    # no-depth instantiation == depth instantiation
    assert torch.equal(F.score_device(X), s)


def test_concurrent_scoring_on_a_shared_forest(nat, oracle, dev):
    """A forest handle is immutable and shared by every executor task thread in the reference
    (IF/IsolationForestModel.scala:129-142): ifb_score_host must be re-entrant on one handle."""
    import threading

    n, d = 200_000, 16
    X = synth_mixture(n, d, 55)
    tables = oracle.fit_forest(X, 40, 256, random_seed=6)
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(X, threads=8)
    parts = [np.ascontiguousarray(X[i::8]) for i in range(8)]
    out = [None] * 8

    def work(i):
        for _ in range(3):
            out[i] = F.score_host(parts[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(8):
        assert np.max(np.abs(out[i] - ref[i::8]) / ref[i::8]) <= 1e-12


def test_empty_and_tiny_inputs(nat, oracle, dev):
    X = synth_mixture(1000, 7, 2)
    tables = oracle.fit_forest(X, 10, 64, random_seed=1)
    F = nat.NativeForest.from_tables(tables)
    assert F.score_host(np.zeros((0, 7), np.float32)).shape == (0,)
    assert F.score_device(torch.zeros(0, 7, device="cuda")).shape == (0,)
    one = F.score_host(X[:1])
    assert one[0] == oracle.Forest(tables).score(X[:1])[0] or abs(one[0] / oracle.Forest(tables).score(X[:1])[0] - 1) < 1e-12


@pytest.mark.parametrize("world,T,d", [(1, 30, 8), (2, 64, 16), (3, 50, 12), (4, 300, 96)])
def test_fused_scatter_emulated_ranks(nat, oracle, dev, world, T, d):
    """ifb_score_scatter_device + ifb_finalize_gathered_device with `world` ranks emulated on one GPU: every rank's
    kernel scatters its partial sums into each owner's [world][rows_o] buffer; owners add them in rank order."""
    import ctypes as C

    n = 50_001
    X = synth_mixture(n, d, 40 + world)
    tables = oracle.fit_forest(X, T, 256, random_seed=12)
    ref, ref_d, ref_p = oracle.Forest(tables).score(X, threads=8, want_parts=True)
    Xd = colmajor_cuda(X)
    cuts = [r * n // world for r in range(world)] + [n]
    bufs = [torch.full((world * (cuts[o + 1] - cuts[o]),), float("nan"), dtype=torch.float32, device="cuda")
            for o in range(world)]
    peer = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    cut_arr = (C.c_int64 * (world + 1))(*cuts)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def shard(t0, t1):
        nb, ne = tables["node_off"][t0], tables["node_off"][t1]
        s = dict(tables)
        s.update(num_trees=t1 - t0, node_off=(tables["node_off"][t0:t1 + 1] - nb).astype(np.int32))
        for k in ("left", "right", "feature", "threshold", "num_instances"):
            s[k] = tables[k][nb:ne]
        return s

    for r in range(world):
        F = nat.NativeForest.from_tables(shard(r * T // world, (r + 1) * T // world))
        nat.check(nat.lib().ifb_score_scatter_device(F.handle, C.c_void_p(Xd.data_ptr()), n, d, n, nat.COL_MAJOR, world, r,
                                                     cut_arr, peer, st))
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    for o in range(world):
        rows = cuts[o + 1] - cuts[o]
        nat.check(nat.lib().ifb_finalize_gathered_device(0, C.c_void_p(bufs[o].data_ptr()), world, rows, T, 256,
                                                         C.c_void_p(out[cuts[o]:].data_ptr()), st))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert not np.isnan(torch.cat(bufs).cpu().numpy()).any()          # every slot was written exactly once
    if world == 1:
        assert np.max(np.abs(got - ref) / ref) <= 1e-12
    else:
        assert np.max(np.abs(got - ref) / ref) < 1e-6                  # f32 sum re-associated across shards
    # the rank-ordered sum of per-shard sequential sums, recomputed on the host, is reproduced bit for bit
    parts = [oracle.Forest(shard(r * T // world, (r + 1) * T // world)).score(X, threads=8, want_parts=True)[2]
             for r in range(world)]
    acc = np.zeros(n, np.float32)
    for part in parts:
        acc = acc + part
    e = acc / np.float32(T)
    z = (-e / oracle.avg_path_length(256)).astype(np.float64)
    assert np.max(np.abs(got - np.power(2.0, z)) / got) <= 1e-12


@pytest.mark.parametrize("d", [16, 64, 6, 128, 260])
def test_extended_exact_ties_are_decided_like_the_reference(nat, oracle, dev, d, monkeypatch):
    """The dense extended kernel decides most visits from an f32 FMA dot product guarded by a rounding bound and
    re-evaluates near ties with the reference's arithmetic.  Hand-built trees put the offset exactly ON the
    reference's sequential value for a row, and one f64 ulp on either side of it."""
    rng = np.random.default_rng(5 + d)
    n = 96
    X = rng.standard_normal((n, d)).astype(np.float32)
    trees = []
    for r in range(n):
        w = rng.standard_normal(d)
        w = (w / np.linalg.norm(w)).astype(np.float32)
        prod = (w * X[r]).astype(np.float32).astype(np.float64)      # f32 products (one rounding each)
        sref = np.cumsum(prod)[-1]                                   # sequential f64 sum, ascending index
        for off in (sref, np.nextafter(sref, np.inf), np.nextafter(sref, -np.inf)):
            trees.append((w, off))
    T = len(trees)
    tables = dict(extended=True, num_trees=T, num_samples=256, total_num_features=d,
                  node_off=np.arange(0, 3 * T + 1, 3, dtype=np.int32), left=np.tile([1, -1, -1], T).astype(np.int32),
                  right=np.tile([2, -1, -1], T).astype(np.int32),
                  num_instances=np.tile([-1, 3, 200], T).astype(np.int64),      # the two leaves are distinguishable
                  offset=np.concatenate([[o, 0.0, 0.0] for _, o in trees]),
                  hp_off=np.concatenate([[0], np.cumsum(np.tile([d, 0, 0], T))]).astype(np.int64),
                  hp_idx=np.tile(np.arange(d, dtype=np.int32), T), hp_w=np.concatenate([w for w, _ in trees]))
    F = nat.NativeForest.from_tables(tables)
    ref = oracle.Forest(tables).score(X, threads=4, want_parts=True)
    # sanity of the construction: row r goes right / left / right in its own three trees
    one = oracle.Forest(tables)
    for r in (0, 17, 95):
        got = [one.path_length(3 * r + j, X[r]) for j in range(3)]
        assert got[0] == got[2] != got[1]
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    # the same through the always-exact path (bound scaled up so that every visit falls back)
    monkeypatch.setenv("IFB_EXT_FAST_SCALE", "1e30")
    F2 = nat.NativeForest.from_tables(tables)
    assert_parity(F2.score_device(colmajor_cuda(X), want_parts=True), ref)


def test_peer_flag_barrier_emulated(nat, dev):
    """ifb_peer_signal_device / ifb_peer_wait_device with three ranks emulated on one GPU (flags are plain device
    memory here; across processes they live in CUDA-IPC mapped peer memory)."""
    import ctypes as C

    world = 3
    flags = [torch.zeros(64, dtype=torch.int32, device="cuda") for _ in range(world)]
    arr = (C.c_void_p * world)(*[f.data_ptr() for f in flags])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for epoch in (1, 2, 7):
        for r in range(world):
            nat.check(nat.lib().ifb_peer_signal_device(0, world, r, arr, epoch, st))
        for r in range(world):
            nat.check(nat.lib().ifb_peer_wait_device(0, world, C.c_void_p(flags[r].data_ptr()), epoch, st))
        torch.cuda.synchronize()
        for r in range(world):
            assert flags[r][:world].tolist() == [epoch] * world


def test_device_buffers_can_be_exported_to_peer_processes(nat, dev):
    """The fused tree-sharded scatter hands ifb_device_alloc buffers to peer ranks through CUDA IPC: the allocation
    must stay a plain device allocation (memory from the stream-ordered pool cannot be exported)."""
    import ctypes as C
    p = C.c_void_p()
    nat.check(nat.lib().ifb_device_alloc(dev, 1 << 20, C.byref(p)))
    try:
        handle = C.create_string_buffer(64)
        nat.check(nat.lib().ifb_ipc_export(dev, p, handle))
        assert any(handle.raw)
    finally:
        nat.check(nat.lib().ifb_device_free(dev, p))
