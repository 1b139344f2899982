"""Prototype (numpy, CPU) of the round-2 idea for score_std_kernel named in DESIGN.md 6b(1): replace the f32
threshold compare of a visit by an integer rank compare so that a node fits ONE 32-bit word.

Exactness argument.  The walk only ever asks `(double)x < t`, which equals `x < c` with c = ceil32(t) (the f32
threshold the current kernel already uses).  For feature f let C_f be the sorted distinct c of all nodes that split on
f, rt(c) its 0-based index, and rx(x) = #{c' in C_f : c' <= x} (np.searchsorted(..., side="right")).  Then
    x < c  <=>  every c' <= x is < c  <=>  rx(x) <= rt(c)  <=>  rx(x) < rt(c) + 1,
and NaN (never `<`) gets rx = len(C_f) + 1, larger than any stored rank.  A row tile is binned once per feature
(a binary search, ~log2|C_f| shared-memory reads) and a visit then needs the node word and the row's 16-bit rank:
2 shared-memory loads instead of 3.

This file is a design artefact: it proves the transform bit-exact against the oracle on random forests and prints the
shared-memory load budget per row for a given forest.  It is NOT product code and nothing imports it but its test.
"""
from __future__ import annotations

import numpy as np


def ceil32(t: np.ndarray) -> np.ndarray:
    """Smallest f32 >= t (t f64, not NaN): `(double)x < t  <=>  x < ceil32(t)` for every f32 x."""
    with np.errstate(over="ignore"):
        c = t.astype(np.float32)
    low = c.astype(np.float64) < t
    c[low] = np.nextafter(c[low], np.float32(np.inf))
    return c


class RankedForest:
    def __init__(self, tables: dict, d: int):
        assert not tables["extended"]
        self.t = tables
        self.d = d
        left, feat = tables["left"], tables["feature"]
        internal = left != -1
        c = np.zeros(len(left), np.float32)
        c[internal] = ceil32(tables["threshold"][internal])
        self.cuts = []                      # per feature: sorted distinct f32 cut points
        self.node_rank = np.zeros(len(left), np.int64)
        for f in range(d):
            sel = internal & (feat == f)
            cf = np.unique(c[sel])
            self.cuts.append(cf)
            self.node_rank[sel] = np.searchsorted(cf, c[sel], side="left") + 1   # rt + 1
        self.max_rank = max((len(cf) for cf in self.cuts), default=0) + 1

    def bin_rows(self, X: np.ndarray) -> np.ndarray:
        """rx per (row, feature); NaN -> len(C_f) + 1."""
        R = np.empty(X.shape, np.int64)
        for f in range(self.d):
            col = X[:, f]
            r = np.searchsorted(self.cuts[f], col, side="right")
            r[np.isnan(col)] = len(self.cuts[f]) + 1
            R[:, f] = r
        return R

    def depth_and_leaf(self, X: np.ndarray):
        """Integer depth sums and the leaf node reached per (row, tree), walking on ranks only."""
        t = self.t
        R = self.bin_rows(X)
        n, T = len(X), t["num_trees"]
        depth = np.zeros(n, np.int64)
        self.tree_depth = np.zeros((n, T), np.int64)     # depth of the reached leaf per (row, tree)
        leaves = np.zeros((n, T), np.int64)
        rows = np.arange(n)
        for k in range(T):
            base = int(t["node_off"][k])
            node = np.full(n, base, np.int64)
            active = t["left"][node] != -1
            while active.any():
                a = np.nonzero(active)[0]
                nd = node[a]
                go_left = R[a, t["feature"][nd]] < self.node_rank[nd]
                node[a] = base + np.where(go_left, t["left"][nd], t["right"][nd])
                depth[a] += 1
                self.tree_depth[a, k] += 1
                active = t["left"][node] != -1
            leaves[:, k] = node
        return depth, leaves

    def budget(self, trees_walk_levels: float = 6.0) -> dict:
        """Shared-memory loads per row: today's 3 per visit vs 2 per visit + binning."""
        T = self.t["num_trees"]
        bin_loads = sum(int(np.ceil(np.log2(len(cf) + 1))) for cf in self.cuts)
        now = 3 * trees_walk_levels * T
        ranked = 2 * trees_walk_levels * T + bin_loads
        return {"now": now, "ranked": ranked, "binning": bin_loads, "saving": 1.0 - ranked / now,
                "max_rank": self.max_rank, "rank_bits": int(np.ceil(np.log2(self.max_rank + 1)))}


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as g
    O = g.load_oracle()
    rng = np.random.default_rng(0)
    for (d, T) in ((32, 100), (128, 512), (128, 64)):
        X = rng.standard_normal((1 << 15, d)).astype(np.float32)
        rf = RankedForest(O.fit_forest(X, T, 256, random_seed=1), d)
        print(f"d={d} trees={T}: {rf.budget()}")
