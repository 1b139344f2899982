"""Counter-based synthetic rows shared by bench.py (both arms), the tests and the tools.

SURVEY.md section 8(d) asks for stateless, counter-based inputs so that the CPU oracle and the GPU regenerate identical
values for any row subset.  Every value depends only on (seed, row, column):

    z1 = splitmix64(row * d + column, seed),  z2 = splitmix64(z1 ^ K)
    g  = float32(2 * (sum of the eight 16-bit fields of z1, z2) - 8 * 65535) * C        (Irwin-Hall(8): mean 0, std 1)
    component(row) from the top 24 bits of a second hash of the row index:
        49 %: x = g          49 %: x = g + 3 / sqrt(d)          2 %: x = 4 * g            (BASELINE's mixture)

Only wrap-around 64-bit integer arithmetic, one exact int -> f32 conversion and at most two IEEE f32 operations are
used, so numpy (uint64) and torch (int64, CPU or CUDA) produce the same bits.  Box-Muller would need log/cos, whose
last bits differ between libm and CUDA; the sum of eight uniforms is the bit-reproducible stand-in for the Gaussian.
"""
from __future__ import annotations

import numpy as np

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB
_K = 0xD1B54A32D192ED03
_STD = 2.0 * np.sqrt(8.0 * (65536.0 ** 2 - 1.0) / 12.0)
C_F32 = np.float32(1.0 / _STD)
T1 = int(0.49 * (1 << 24))
T2 = int(0.98 * (1 << 24))


def _s64(c: int) -> int:
    """Two's-complement view of a 64-bit constant (torch has no uint64 arithmetic)."""
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


# ---- numpy -------------------------------------------------------------------------------------------------------------
def _mix_np(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
    return z ^ (z >> np.uint64(31))


def rows_numpy(rows, d: int, seed: int) -> np.ndarray:
    """(len(rows) x d) float32, C order, for arbitrary global row indices."""
    rows = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = np.uint64((seed * _GOLD) & _M64)
        ctr = rows[:, None] * np.uint64(d) + np.arange(d, dtype=np.uint64)[None, :]
        z1 = _mix_np(ctr * np.uint64(_GOLD) + key)
        z2 = _mix_np(z1 ^ np.uint64(_K))
        isum = np.zeros(z1.shape, np.int64)
        for z in (z1, z2):
            for sh in (0, 16, 32, 48):
                isum += ((z >> np.uint64(sh)) & np.uint64(0xFFFF)).astype(np.int64)
        g = (2 * isum - 8 * 65535).astype(np.float32) * C_F32
        zr = _mix_np(_mix_np(rows * np.uint64(_GOLD) + np.uint64(((seed + 1) * _C2) & _M64)))
        u = (zr >> np.uint64(40)).astype(np.int64)
    shift = np.float32(3.0 / np.sqrt(d))
    x = g
    c1 = (u >= T1) & (u < T2)
    x[c1] = x[c1] + shift
    c2 = u >= T2
    x[c2] = x[c2] * np.float32(4.0)
    return x


def matrix_numpy(n: int, d: int, seed: int, row0: int = 0, block: int = 1 << 16) -> np.ndarray:
    out = np.empty((n, d), np.float32)
    for b in range(0, n, block):
        e = min(n, b + block)
        out[b:e] = rows_numpy(np.arange(row0 + b, row0 + e, dtype=np.uint64), d, seed)
    return out


# ---- torch ---------------------------------------------------------------------------------------------------------------
def _lsr(torch, z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix_t(torch, z):
    z = (z ^ _lsr(torch, z, 30)) * _s64(_C1)
    z = (z ^ _lsr(torch, z, 27)) * _s64(_C2)
    return z ^ _lsr(torch, z, 31)


def matrix_torch(torch, n: int, d: int, seed: int, device, row0: int = 0, cols_per_pass: int | None = None):
    """Column-major (n x d) float32 view (memory layout [d][n]) of global rows [row0, row0 + n), generated on `device`."""
    xt = torch.empty((d, n), dtype=torch.float32, device=device)
    rows = torch.arange(row0, row0 + n, dtype=torch.int64, device=device)
    zr = _mix_t(torch, _mix_t(torch, rows * _s64(_GOLD) + _s64((seed + 1) * _C2)))
    u = _lsr(torch, zr, 40)
    c1 = ((u >= T1) & (u < T2))
    c2 = (u >= T2)
    del zr, u
    shift = float(np.float32(3.0 / np.sqrt(d)))
    base = rows * d
    key = _s64(seed * _GOLD)
    if cols_per_pass is None:
        cols_per_pass = max(1, min(d, (1 << 26) // max(n, 1)))
    for c0 in range(0, d, cols_per_pass):
        c1_ = min(d, c0 + cols_per_pass)
        cols = torch.arange(c0, c1_, dtype=torch.int64, device=device)
        ctr = base[None, :] + cols[:, None]
        z1 = _mix_t(torch, ctr * _s64(_GOLD) + key)
        z2 = _mix_t(torch, z1 ^ _s64(_K))
        isum = torch.zeros_like(z1)
        for z in (z1, z2):
            for sh in (0, 16, 32, 48):
                isum += _lsr(torch, z, sh) & 0xFFFF if sh else z & 0xFFFF
        g = (2 * isum - 8 * 65535).to(torch.float32) * float(C_F32)
        g = torch.where(c1[None, :], g + shift, g)
        g = torch.where(c2[None, :], g * 4.0, g)
        xt[c0:c1_] = g
        del ctr, z1, z2, isum, g
    return xt.t()
