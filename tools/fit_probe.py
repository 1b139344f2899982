"""One warm-up fit and one profiled fit of a BASELINE-shaped forest (used under ncu -k regex:fit_kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
nat = graft.load_package()._native
d, T, ext = int(os.environ.get("D", 128)), int(os.environ.get("T", 512)), int(os.environ.get("EXT", -1))
g = torch.Generator(device="cuda").manual_seed(7)
X = torch.randn(d, 1 << 21, device="cuda", generator=g).t()
for _ in range(2):
    f = nat.fit_device(X, nat.FitParams(T, 256, d, 0, 1, 1, ext, 0, 0))
torch.cuda.synchronize()
print("nodes", f.info().num_nodes)
