import os, sys, time, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as graft, synthdata
nat = graft.load_package()._native
for (rows, d, T, ext) in [(1 << 17, 1024, 256, 1023), (1 << 20, 64, 200, 63), (1 << 20, 128, 512, -1), (1 << 20, 32, 100, -1)]:
    X = synthdata.matrix_torch(torch, rows, d, 4242, 'cuda')
    prm = nat.FitParams(T, 256, d, 0, 1, 1, ext, 0, 0)
    nat.fit_device(X, prm); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); f = nat.fit_device(X, prm); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    os.environ["IFB_FIT_TIMING"] = "1"
    f = nat.fit_device(X, prm); torch.cuda.synchronize()
    del os.environ["IFB_FIT_TIMING"]
    print(f"fit rows={rows} d={d} T={T} ext={ext}: call min {min(ts):.2f} ms median {sorted(ts)[2]:.2f} ms", flush=True)
