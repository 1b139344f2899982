"""Developer timing of the host-mirror paths (fit / transform through libifb200_host.so)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
rng = np.random.default_rng(0)
n, d = 4_000_000, 32
X = rng.standard_normal((n, d))                      # f64, what Spark Vectors hold
t0 = time.perf_counter(); m = pkg.IsolationForest().setRandomSeed(1).fit(X); t1 = time.perf_counter()
print(f"fit {n}x{d} f64 (100 trees, contamination 0): {t1 - t0:.3f} s")
for rep in range(3):
    t0 = time.perf_counter(); out = m.transform(X); t1 = time.perf_counter()
    print(f"transform f64 {n}x{d}: {t1 - t0:.3f} s = {n / (t1 - t0):.3e} rows/s")
X32 = X.astype(np.float32)
for rep in range(2):
    t0 = time.perf_counter(); out = m.transform(X32); t1 = time.perf_counter()
    print(f"transform f32 {n}x{d}: {t1 - t0:.3f} s = {n / (t1 - t0):.3e} rows/s")
t0 = time.perf_counter(); m2 = pkg.IsolationForest().setRandomSeed(1).setContamination(0.02).fit(X32); t1 = time.perf_counter()
print(f"fit with contamination 0.02 (threshold pass) f32: {t1 - t0:.3f} s  thr={m2.getOutlierScoreThreshold():.6f}")
import torch
nat = pkg._native
for (nn, dd, T, ext) in ((2_000_000, 128, 512, -1), (1_000_000, 64, 200, 63), (1_000_000, 32, 100, -1), (200_000, 1024, 256, 1023)):
    Xd = torch.randn(dd, nn, device="cuda").t()
    prm = nat.FitParams(T, 256, dd, 0, 1, 1, ext, 0, 0)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); F = nat.fit_device(Xd, prm); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"ifb_fit_device {nn}x{dd} T={T} ext={ext}: {1e3 * (t1 - t0):.1f} ms  nodes={F.info().num_nodes}")
    del Xd
