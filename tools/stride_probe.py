"""Does the column stride (ld) of a device-resident column-major matrix change the scoring rate?
Scores the same 2M x 128 rows once as a compact matrix and once as a view into a 100M-row allocation."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
nat = graft.load_package()._native
dev = torch.device("cuda", 0)
d, T, ns = 128, int(os.environ.get("T", 512)), 256
n = 2_000_000
big_n = int(os.environ.get("BIG", 100_000_000))
g = torch.Generator(device=dev).manual_seed(7)
big = torch.randn(d, big_n, device=dev, generator=g)
forest = nat.fit_device(big[:, :1 << 20].t(), nat.FitParams(T, ns, d, 0, 1, 1, -1, 0, 0))
def timeit(X, reps=5):
    s = torch.empty(X.shape[0], dtype=torch.float64, device=dev)
    forest.score_device(X, scores=s); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): forest.score_device(X, scores=s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, s
compact = big[:, :n].contiguous()
t_c, s_c = timeit(compact.t())
t_v, s_v = timeit(big[:, :n].t())
t_mid, _ = timeit(big[:, 50_000_000:50_000_000 + n].t()) if big_n >= 52_000_000 else (float("nan"), None)
print(f"T={T} compact ld={n}: {t_c:.3f} ms | view ld={big_n}: {t_v:.3f} ms | view mid: {t_mid:.3f} ms | equal={bool(torch.equal(s_c, s_v))}")
t_all, _ = timeit(big.t(), reps=2)
print(f"all {big_n} rows: {t_all:.2f} ms = {big_n / t_all * 1e3:.3e} rows/s (2M-row rate would give {t_c * big_n / n:.1f} ms)")
t0 = time.perf_counter(); f2 = nat.fit_device(big.t(), nat.FitParams(T, ns, d, 0, 1, 1, -1, 0, 0)); torch.cuda.synchronize()
print(f"fit on {big_n} rows: {(time.perf_counter() - t0) * 1e3:.1f} ms")
