"""GPU tests of the tensor-core path of fully-extended forests (csrc/score_ext_tc.cu).

The tensor cores only FILTER visits (a visit whose accumulator is provably far from the offset is accepted, everything
else is decided with the reference's arithmetic), so the parity bar is the same as everywhere: integer depth sums and
f32 path sums bit-identical to the oracle.  `test_accumulators_*` additionally measures, on the part, the one hardware
assumption the filter's bound constant makes: the accumulation error of tcgen05.mma (kind::f16, f32 accumulate).
"""
import numpy as np
import pytest
import torch

from conftest import synth_mixture
from test_score_gpu import assert_parity, colmajor_cuda, dev  # noqa: F401

pytestmark = pytest.mark.gpu


def bfs_slot_rows(t):
    """forest.cu numbers slots along the BFS order with the two children of a node adjacent (a queue, not levels)."""
    rows = []
    for tr in range(int(t["num_trees"])):
        b = int(t["node_off"][tr])
        order = [0]
        h = 0
        while h < len(order):
            g = b + order[h]
            if t["left"][g] != -1:
                order += [int(t["left"][g]), int(t["right"][g])]
            h += 1
        rows += [b + i for i in order if t["left"][b + i] != -1]
    return rows


def split16(v, by_norm=False, extra=None):
    """Power-of-two scaling (weights: max |w| into [0.5, 1); rows: ||x||_2 into [0.5, 1)) and the two-term fp16 split
    the kernels use.  `extra` shifts the exponent per row (the kernel's f32 norm may land on the other side of a power
    of two than the f64 norm computed here)."""
    v = np.asarray(v, np.float32)
    mx = np.abs(v).max(axis=-1, keepdims=True)
    _, e = np.frexp(mx)
    e = np.where(mx > 0, e, 0)
    if by_norm:
        s1 = v.astype(np.float64) * np.exp2(-e.astype(np.float64))
        nrm = np.sqrt((s1 * s1).sum(axis=-1, keepdims=True)) * (1.0 + 2.0 ** -10)
        _, e2 = np.frexp(nrm)
        e = e + np.where(nrm > 0, e2, 0)
    if extra is not None:
        e = e + extra
    s = (v * np.exp2(-e).astype(np.float32)).astype(np.float32)
    hi = s.astype(np.float16)
    lo = (s - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), e


@pytest.mark.parametrize("d,T", [(64, 12), (1024, 6), (6, 20), (200, 8)])
def test_accumulators_are_within_the_assumed_error(nat, oracle, dev, d, T):
    X = synth_mixture(3000, d, 4000 + d)
    t = oracle.fit_forest(X, T, 256, random_seed=2, ext_level=d - 1)
    F = nat.NativeForest.from_tables(t)
    kp, ncols = F.ext_tc_info()
    assert kp >= d and kp % 32 == 0 and ncols > 0 and ncols % 256 == 0
    Xp = X[:128]
    scores, acc, slots = F.ext_tc_probe(colmajor_cuda(Xp))
    acc = acc.cpu().numpy().astype(np.float64)
    ref = oracle.Forest(t).score(Xp, threads=4)
    assert np.max(np.abs(scores.cpu().numpy() - ref) / ref) <= 1e-12
    rows = bfs_slot_rows(t)
    assert slots.max() == len(rows) - 1
    W = np.stack([t["hp_w"][int(t["hp_off"][g]):int(t["hp_off"][g + 1])] for g in rows])
    wh, wl, _ = split16(W)
    cols = np.where(slots >= 0)[0]
    wh, wl = wh[slots[cols]], wl[slots[cols]]
    xh, xl, _ = split16(Xp, by_norm=True)
    exact = xh @ wh.T + xl @ wh.T + xh @ wl.T
    # the kernel's row exponent comes from an f32 norm: allow it to differ by one from the f64 estimate
    big = np.abs(exact) > 1e-3
    shift = np.array([np.round(np.median(np.log2(np.abs(acc[i, cols][big[i]] / exact[i][big[i]])))) if big[i].any() else 0.0
                      for i in range(len(Xp))])
    assert np.all(np.abs(shift) <= 1)
    if np.any(shift != 0):
        xh, xl, _ = split16(Xp, by_norm=True, extra=-shift.astype(np.int64)[:, None])
        exact = xh @ wh.T + xl @ wh.T + xh @ wl.T
    mag = np.abs(xh) @ np.abs(wh).T + np.abs(xl) @ np.abs(wh).T + np.abs(xh) @ np.abs(wl).T
    err = np.abs(acc[:, cols] - exact) / mag
    nsteps = 3 * (kp // 16)
    worst = float(err.max())
    print(f"\n[tc accuracy] d={d} kp={kp} columns={len(cols)} max |acc - exact| / sum|a b| = 2^{np.log2(max(worst, 1e-300)):.2f} "
          f"(assumed <= {nsteps} steps x 2^-22 = 2^{np.log2(nsteps * 2.0**-22):.2f}; per step 2^{np.log2(max(worst, 1e-300) / nsteps):.2f})")
    assert worst <= nsteps * 2.0 ** -22
    # padding columns accumulate exact zeros
    pad = np.where(slots < 0)[0]
    if len(pad):
        assert np.all(acc[:, pad] == 0.0)


@pytest.mark.parametrize("n,d,T", [(40_000, 64, 60), (130, 64, 3), (6_000, 1024, 12), (10_000, 96, 30), (4_097, 33, 17)])
def test_tc_path_matches_the_oracle(nat, oracle, dev, n, d, T, monkeypatch):
    X = synth_mixture(n, d, 5000 + d)
    t = oracle.fit_forest(X[:min(n, 20_000)], T, min(256, n), random_seed=4, ext_level=d - 1)
    F = nat.NativeForest.from_tables(t)
    assert F.ext_tc_info()[1] > 0, "forest did not get a tensor-core layout"
    ref = oracle.Forest(t).score(X, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    assert_parity(F.score_device(torch.from_numpy(X).cuda(), want_parts=True), ref)
    # the CUDA-core kernels and the tensor-core path implement one contract
    monkeypatch.setenv("IFB_EXT_NO_TC", "1")
    old = [a.cpu().numpy() for a in F.score_device(colmajor_cuda(X), want_parts=True)]
    monkeypatch.delenv("IFB_EXT_NO_TC")
    new = [a.cpu().numpy() for a in F.score_device(colmajor_cuda(X), want_parts=True)]
    for a, b in zip(old, new):
        assert np.array_equal(a, b)
    # every visit through the exact path (bound scaled up), and none (bound 0: only exact zeros stay ambiguous)
    monkeypatch.setenv("IFB_TC_SCALE", "1e30")
    assert_parity(F.score_device(colmajor_cuda(X[:2048]), want_parts=True), [a[:2048] for a in ref])


@pytest.mark.parametrize("n,d,T,ext,nf", [(30_000, 64, 40, 9, None), (5_000, 8, 50, 3, None), (9_000, 200, 12, 20, None),
                                           (20_000, 32, 30, 0, None), (12_000, 48, 25, 47, 16), (4_000, 1024, 6, 63, None)])
def test_sparse_hyperplanes_take_the_tensor_core_path(nat, oracle, dev, n, d, T, ext, nf, monkeypatch):
    """extensionLevel < d - 1 (or a feature subspace, maxFeatures < 1): the stored terms become zero-padded accumulator
    columns; ambiguous visits are decided on the stored terms in their stored order."""
    X = synth_mixture(n, d, 7000 + d + ext)
    kw = {} if nf is None else {"num_features": nf}
    t = oracle.fit_forest(X[:min(n, 20_000)], T, 256, random_seed=6, ext_level=min(ext, (nf or d) - 1), **kw)
    F = nat.NativeForest.from_tables(t)
    assert F.ext_tc_info() == ((d + 31) // 32 * 32, F.ext_tc_info()[1]) and F.ext_tc_info()[1] > 0
    Xs = X.copy()
    Xs[5::97, 1] = np.nan          # non-finite features, also in coordinates a hyperplane may not read
    Xs[7::89, d - 1] = np.inf
    Xs[11::83, :] = 0.0
    with np.errstate(all="ignore"):
        ref = oracle.Forest(t).score(Xs, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)
    assert_parity(F.score_device(torch.from_numpy(Xs).cuda(), want_parts=True), ref)
    monkeypatch.setenv("IFB_TC_SCALE", "1e30")      # every visit through the exact path on the stored terms
    assert_parity(F.score_device(colmajor_cuda(Xs[:2048]), want_parts=True), [a[:2048] for a in ref])
    monkeypatch.delenv("IFB_TC_SCALE")
    monkeypatch.setenv("IFB_EXT_NO_TC", "1")        # the CUDA-core kernels agree bit for bit
    old = [a.cpu().numpy() for a in F.score_device(colmajor_cuda(Xs), want_parts=True)]
    monkeypatch.delenv("IFB_EXT_NO_TC")
    new = [a.cpu().numpy() for a in F.score_device(colmajor_cuda(Xs), want_parts=True)]
    for a, b in zip(old, new):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("env", [{"IFB_TC_CG": "2"}, {"IFB_TC_BK": "16"}, {"IFB_TC_BK": "16", "IFB_TC_CLUSTER": "4"},
                                 {"IFB_TC_CLUSTER": "1"}])
@pytest.mark.parametrize("n,d,T", [(20_000, 64, 30), (3_000, 1024, 8)])
def test_kernel_variants_agree_with_the_oracle(nat, oracle, dev, monkeypatch, env, n, d, T):
    """The measured alternatives stay correct: the cta_group::2 pair (one tcgen05.mma for the 2 x 128 rows of a
    two-CTA cluster, half of the hyperplane tile per CTA), the 32-byte-row operand ring, other cluster sizes."""
    X = synth_mixture(n, d, 8100 + d)
    t = oracle.fit_forest(X[:min(n, 20_000)], T, 256, random_seed=9, ext_level=d - 1)
    F = nat.NativeForest.from_tables(t)
    assert F.ext_tc_info()[1] > 0
    Xs = X.copy()
    Xs[3::101, 2] = np.nan
    Xs[9::53, :] = 0.0
    with np.errstate(all="ignore"):
        ref = oracle.Forest(t).score(Xs, threads=8, want_parts=True)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)


def test_tc_special_values(nat, oracle, dev):
    n, d = 8192, 64
    X = synth_mixture(n, d, 77)
    te = oracle.fit_forest(X, 12, 256, random_seed=8, ext_level=d - 1)
    Xs = X.copy()
    Xs[3::64, 5] = 0.0
    Xs[7::64, 9] = -0.0
    Xs[11::64, 1] = np.float32(1e-42)
    Xs[13::64, 2] = np.float32(3e-38)
    Xs[17::64, 3] = np.float32(2e38)
    Xs[19::64, 4] = np.inf
    Xs[23::64, 6] = np.nan
    Xs[29::64, :] = 0.0
    Xs[31::64, :] *= np.float32(1e-30)
    Xs[37::64, :] *= np.float32(1e25)
    F = nat.NativeForest.from_tables(te)
    assert F.ext_tc_info()[1] > 0
    with np.errstate(all="ignore"):
        ref = oracle.Forest(te).score(Xs, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(Xs), want_parts=True), ref)


def test_tc_partial_sums_accumulate_across_forest_shards(nat, oracle, dev):
    """ifb_score_partial_device on two tree shards (tree-sharded multi-GPU layout) == one forest, bit for bit in the depths."""
    n, d, T = 9_000, 64, 24
    X = synth_mixture(n, d, 9)
    t = oracle.fit_forest(X, T, 256, random_seed=6, ext_level=d - 1)
    ref = oracle.Forest(t).score(X, threads=8, want_parts=True)

    def shard(a, b):
        no = t["node_off"]
        sl = slice(int(no[a]), int(no[b]))
        h0, h1 = int(t["hp_off"][no[a]]), int(t["hp_off"][no[b]])
        return dict(extended=True, num_trees=b - a, num_samples=256, total_num_features=d,
                    node_off=(no[a:b + 1] - no[a]).astype(np.int32), left=t["left"][sl], right=t["right"][sl],
                    num_instances=t["num_instances"][sl], offset=t["offset"][sl],
                    hp_off=(t["hp_off"][int(no[a]):int(no[b]) + 1] - h0), hp_idx=t["hp_idx"][h0:h1], hp_w=t["hp_w"][h0:h1])

    Xd = colmajor_cuda(X)
    ps = torch.zeros(n, dtype=torch.float32, device="cuda")
    ds = torch.zeros(n, dtype=torch.int32, device="cuda")
    for a, b in ((0, 10), (10, T)):
        nat.NativeForest.from_tables(shard(a, b)).score_partial_device(Xd, ps, ds)
    assert np.array_equal(ds.cpu().numpy(), ref[1])
    assert np.array_equal(ps.cpu().numpy(), ref[2])      # same sequential order: shard 0's sum carried into shard 1
    sc = nat.finalize_scores_device(ps, T, 256).cpu().numpy()
    assert np.max(np.abs(sc - ref[0]) / ref[0]) <= 1e-12


def test_hyperplanes_wider_than_the_staging_limit_use_the_cuda_core_kernels(nat, oracle, dev):
    """k > 1536 does not get a tensor-core layout (the row-preparation tile would not fit shared memory); the wide
    CUDA-core kernel scores it, same parity bar."""
    n, d, T = 700, 2000, 3
    X = synth_mixture(n, d, 6000 + d)
    t = oracle.fit_forest(X, T, 256, random_seed=4, ext_level=d - 1)
    F = nat.NativeForest.from_tables(t)
    assert F.ext_tc_info() == (0, 0)
    ref = oracle.Forest(t).score(X, threads=8, want_parts=True)
    assert_parity(F.score_device(colmajor_cuda(X), want_parts=True), ref)
    Fg = nat.fit_device(colmajor_cuda(X), nat.FitParams(T, 256, d, 0, 4, 1, d - 1, 0, 0))    # device-resident hyperplanes
    assert Fg.ext_tc_info() == (0, 0)
    assert_parity(Fg.score_device(colmajor_cuda(X), want_parts=True), oracle.Forest(Fg.export()).score(X, threads=8, want_parts=True))
