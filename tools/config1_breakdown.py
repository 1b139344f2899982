"""Where BASELINE config 1's (fit + transform, 1,000 x 10, 100 trees) time goes: estimator fit, model transform, and the
bare C-ABI calls underneath (developer tool)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package(); nat = pkg._native
rng = np.random.default_rng(1001)
X = rng.standard_normal((1000, 10))
pkg.IsolationForest().fit(X).transform(X)
def med(f, n=30):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))
est = lambda: pkg.IsolationForest().setNumEstimators(100).setMaxSamples(256).setRandomSeed(1)
m = est().fit(X)
print(f"estimator.fit      {med(lambda: est().fit(X)):.3f} ms")
print(f"model.transform    {med(lambda: m.transform(X)):.3f} ms")
X32 = np.ascontiguousarray(X.astype(np.float32))
Xd = torch.from_numpy(X32).cuda()
prm = nat.FitParams(100, 256, 10, 0, 1, 1, -1, 0, 0)
def fitdev():
    f = nat.fit_device(Xd, prm); torch.cuda.synchronize(); return f
F = fitdev()
print(f"ifb_fit_device     {med(fitdev):.3f} ms")
print(f"ifb_score_host     {med(lambda: F.score_host(X32)):.3f} ms")
s = torch.empty(1000, dtype=torch.float64, device='cuda')
def sd():
    F.score_device(Xd, scores=s); torch.cuda.synchronize()
print(f"ifb_score_device   {med(sd):.3f} ms")
def h2d():
    torch.from_numpy(X32).cuda(); torch.cuda.synchronize()
print(f"H2D 40 KB (torch)  {med(h2d):.3f} ms")
os.environ["IFB_FIT_TIMING"] = "1"
fitdev()
