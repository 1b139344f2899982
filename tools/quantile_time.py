"""Times ifb_quantile_device (8-pass MSD radix select, csrc/epilogue.cu) on 10M and 100M scores and checks the value
against torch.sort.  Run under gpurun; the output is kept under profiles/."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

nat = g.load_package()._native
for n in (10_000_000, 100_000_000):
    gen = torch.Generator(device="cuda").manual_seed(5)
    s = torch.rand(n, device="cuda", generator=gen, dtype=torch.float64) * 0.6 + 0.2
    q = 0.98
    v, frac = nat.quantile_device(s, q)
    rank = min(max(int(-(-q * n // 1)), 1), n)
    ref = float(torch.sort(s).values[rank - 1])
    assert v == ref, (v, ref)
    for _ in range(3):
        nat.quantile_device(s, q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        nat.quantile_device(s, q)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(f"ifb_quantile_device n={n}: {ms:.3f} ms per call (blocking, value on the host), "
          f"{9 * n * 8 / ms / 1e6:.0f} GB/s over its 9 passes, exact order statistic == torch.sort, "
          f"observed fraction >= value {frac:.6f}", flush=True)
