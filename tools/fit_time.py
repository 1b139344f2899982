"""Phase timing of ifb_fit_device (IFB_FIT_TIMING=1) for the BASELINE shapes."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
nat = graft.load_package()._native
dev = torch.device("cuda", 0)
for (rows, d, T, ext) in [(1 << 20, 32, 100, -1), (1 << 22, 128, 512, -1), (1 << 22, 128, 64, -1), (1 << 20, 64, 200, 63)]:
    g = torch.Generator(device=dev).manual_seed(7)
    X = torch.randn(d, rows, device=dev, generator=g).t()
    prm = nat.FitParams(T, 256, d, 0, 1, 1, ext, 0, 0)
    nat.fit_device(X, prm); torch.cuda.synchronize()
    os.environ["IFB_FIT_TIMING"] = "1"
    print(f"--- rows={rows} d={d} T={T} ext={ext}", file=sys.stderr, flush=True)
    t0 = time.perf_counter(); f = nat.fit_device(X, prm); torch.cuda.synchronize()
    print(f"total {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)
    del os.environ["IFB_FIT_TIMING"]
