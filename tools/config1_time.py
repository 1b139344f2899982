"""BASELINE configs[0]: IsolationForest fit + transform, 1k x 10 synthetic Gaussian, 100 trees, maxSamples 256."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); O = g.load_oracle()
rng = np.random.default_rng(1001)
X = rng.standard_normal((1000, 10))
pkg.IsolationForest().fit(X).transform(X)          # warm-up (CUDA context, pools)
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    m = pkg.IsolationForest().setNumEstimators(100).setMaxSamples(256).setRandomSeed(1).fit(X)
    out = m.transform(X)
    ts.append(time.perf_counter() - t0)
print(f"GPU host-mirror fit+transform 1000x10: median {1e3*np.median(ts):.2f} ms (min {1e3*min(ts):.2f})")
X32 = X.astype(np.float32)
tc = []
for _ in range(5):
    t0 = time.perf_counter()
    tb = O.fit_forest(X32, 100, 256, random_seed=1)
    s = O.Forest(tb).score(X32, threads=1)
    tc.append(time.perf_counter() - t0)
print(f"CPU port fit+transform 1000x10 (1 thread): median {1e3*np.median(tc):.2f} ms")
ref = O.Forest(m.tables() | {"num_samples": 256}).score(X32)
print("scores match oracle on the GPU-built forest:", float(np.max(np.abs(out.outlierScore - ref) / ref)))
