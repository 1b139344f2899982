"""Developer timing helper (NOT bench.py): times the scoring kernels on synthetic data with CUDA events.
Uses the oracle only to produce a forest (dev convenience); never imported by the product or bench.py."""
import argparse
import sys
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--d", type=int, default=32)
ap.add_argument("--trees", type=int, default=100)
ap.add_argument("--ext", type=int, default=-1)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--host", action="store_true")
a = ap.parse_args()

nat = g.load_package()._native
O = g.load_oracle()
gen = torch.Generator(device="cuda").manual_seed(1002)
Xt = torch.randn(a.d, a.n, device="cuda", generator=gen)
Xt[:, : a.n // 2] += 3.0 / np.sqrt(a.d)
Xt[:, -a.n // 50:] *= 4.0
X = Xt.t()
fit_rows = np.ascontiguousarray(X[:: max(1, a.n // 8192)][:8192].cpu().numpy())
tables = O.fit_forest(fit_rows, a.trees, 256, random_seed=1, ext_level=a.ext)
F = nat.NativeForest.from_tables(tables)
info = F.info()
print(f"forest: trees={info.num_trees} nodes={info.num_nodes} depth={info.max_depth} nnz={info.max_nnz}")
scores = torch.empty(a.n, dtype=torch.float64, device="cuda")
for _ in range(2):
    F.score_device(X, scores=scores)
torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    F.score_device(X, scores=scores)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
bytes_alg = a.n * (4 * a.d + 8)
print(f"device-resident: {ms:.3f} ms/pass  {a.n / ms * 1e3:.3e} rows/s  {bytes_alg / ms / 1e6:.1f} GB/s algorithmic "
      f"(all: {['%.3f' % t for t in ts]})")
sub = X[::9973]
ref = O.Forest(tables).score(np.ascontiguousarray(sub.cpu().numpy()), threads=8)
print("max rel vs oracle on sample:", float(np.max(np.abs(scores[::9973].cpu().numpy() - ref) / ref)))
if a.host:
    pb = nat.PinnedBuffer((a.d, a.n), np.float32)
    pb.array[...] = Xt.cpu().numpy()
    hs = nat.PinnedBuffer((a.n,), np.float64)
    Xh = pb.array.T
    for _ in range(2):
        t0 = time.perf_counter()
        nat.check(nat.lib().ifb_score_host(F.handle, pb.array.ctypes.data, a.n, a.d, a.n, 0, hs.array.ctypes.data, None, None))
        dt = time.perf_counter() - t0
        print(f"host e2e: {dt * 1e3:.1f} ms  {a.n / dt:.3e} rows/s  h2d {a.n * a.d * 4 / dt / 1e9:.1f} GB/s")
    print("host == device:", bool(np.array_equal(hs.array, scores.cpu().numpy())))
