"""Small end-to-end invocation of every kernel family, meant to be run under compute-sanitizer."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
nat = g.load_package()._native
O = g.load_oracle()
rng = np.random.default_rng(0)

def cm(X):
    return torch.from_numpy(np.ascontiguousarray(X.T)).cuda().t()

if "--fit-only" in sys.argv:   # the tree builder alone (shared-memory staged, scratch staged, extended)
    for (n, d, T, ext, ns) in ((3000, 32, 6, -1, 256), (1500, 128, 4, -1, 256), (700, 300, 3, -1, 256), (900, 5, 3, -1, 700),
                               (2000, 64, 3, 63, 256), (1200, 12, 3, 3, 128), (600, 300, 2, 299, 128)):
        X = rng.standard_normal((n, d)).astype(np.float32)
        X[::9, 0] = 0.25
        for colmajor in (True, False):
            Xd = cm(X) if colmajor else torch.from_numpy(X).cuda()
            tb = nat.fit_device(Xd, nat.FitParams(T, ns, d, 0, 1, 1, ext, 0, 0)).export()
            ref = O.fit_forest(X, T, ns, random_seed=1, ext_level=ext)
            for k in ("node_off", "left", "right", "num_instances"):
                assert np.array_equal(tb[k], ref[k]), (n, d, T, ext, k)
        print("fit ok", n, d, T, ext, ns, flush=True)
    print("sanitize_run (fit only) done")
    sys.exit(0)

for (n, d, T, ext) in ((3000, 32, 20, -1), (1500, 128, 30, -1), (700, 900, 6, -1), (2500, 8, 10, 7), (2000, 64, 6, 63),
                       (600, 256, 4, 255), (1200, 12, 5, 3)):
    X = rng.standard_normal((n, d)).astype(np.float32)
    prm = nat.FitParams(T, min(256, n), d, 0, 1, 1, ext, 0, 0)
    F = nat.fit_device(cm(X), prm)                                  # fit kernel
    tb = F.export()
    ref = O.Forest(tb).score(X, want_parts=True)
    s, ds, ps = F.score_device(cm(X), want_parts=True)              # scoring kernels (col-major)
    s2 = F.score_device(torch.from_numpy(X).cuda())                 # row-major (transpose / generic)
    s3 = F.score_host(X)                                            # host pipeline
    torch.cuda.synchronize()
    assert np.array_equal(ds.cpu().numpy(), ref[1]), (n, d, T, ext)
    assert np.max(np.abs(s.cpu().numpy() - ref[0]) / ref[0]) < 1e-12
    assert np.max(np.abs(s2.cpu().numpy() - ref[0]) / ref[0]) < 1e-12 and np.max(np.abs(s3 - ref[0]) / ref[0]) < 1e-12
    thr, frac = nat.quantile_device(s, 0.9)                         # radix select
    lab = nat.predict_device(s, thr)
    print("ok", n, d, T, ext, flush=True)
# measured alternatives of the tensor-core kernel (cta_group::2 pair, 32-byte-row ring) and the CUDA-core fallback whose
# per-tree blobs are built on first use from a device-fitted forest
X = rng.standard_normal((1500, 64)).astype(np.float32)
F = nat.fit_device(cm(X), nat.FitParams(5, 256, 64, 0, 1, 1, 63, 0, 0))
ref = O.Forest(F.export()).score(X, want_parts=True)
for env in ({"IFB_TC_CG": "2"}, {"IFB_TC_BK": "16"}, {"IFB_EXT_NO_TC": "1"}):
    os.environ.update(env)
    s, ds, ps = F.score_device(cm(X), want_parts=True)
    torch.cuda.synchronize()
    for k in env:
        del os.environ[k]
    assert np.array_equal(ds.cpu().numpy(), ref[1]) and np.array_equal(ps.cpu().numpy(), ref[2]), env
    print("variant ok", env, flush=True)

# the opt-in rank-word standard kernel (score_std_rank.cu): bulk-copy tiles, a ragged tail, plain-load tiles
os.environ["IFB_STD_RANK"] = "1"
for (n, d, T) in ((2600, 32, 20), (1537, 7, 9)):
    X = rng.standard_normal((n, d)).astype(np.float32)
    tb = O.fit_forest(X, T, 256, random_seed=3)
    X[::11, min(1, d - 1)] = np.nan
    F = nat.NativeForest.from_tables(tb)
    assert F.std_rank_chunks(d) >= 1
    ref = O.Forest(tb).score(X, want_parts=True)
    ps = torch.zeros(n, dtype=torch.float32, device="cuda")
    F.score_partial_device(cm(X), ps)
    s = F.score_device(cm(X))
    buf = torch.zeros(d * (n + 3) + 8, dtype=torch.float32, device="cuda")      # unaligned: plain loads
    view = buf[1:1 + d * (n + 3)].view(d, n + 3)[:, :n]
    view.copy_(torch.from_numpy(np.ascontiguousarray(X.T)))
    s2 = F.score_device(view.t())
    torch.cuda.synchronize()
    assert np.array_equal(ps.cpu().numpy(), ref[2]) and np.array_equal(s.cpu().numpy(), s2.cpu().numpy())
    assert np.max(np.abs(s.cpu().numpy() - ref[0]) / ref[0]) < 1e-12
    print("rank ok", n, d, T, flush=True)
del os.environ["IFB_STD_RANK"]
print("sanitize_run done")
