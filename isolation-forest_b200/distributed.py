"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.

Two layouts:
  * rows sharded, forest replicated -- what the reference itself does (broadcast trees, row-parallel transform,
    IF/IsolationForestModel.scala:129-142): no data-path collective at all;
  * trees sharded (BASELINE.json config 4): every rank scores ALL rows against its slice of the ensemble,
    one all-reduce(sum) of the per-row f32 path-length sums (and, optionally, the exact int32 depth sums),
    then the 2^(-E/c) epilogue with the FULL ensemble size.  Fit under the same layout needs no collective
    beyond gathering the finished node tables, because a tree depends only on (seed, P, tree id, data).
"""
from __future__ import annotations

import numpy as np


def tree_shard(num_trees: int, rank: int, world: int):
    """Contiguous, balanced slice [t0, t1) of the ensemble owned by `rank`."""
    return rank * num_trees // world, (rank + 1) * num_trees // world


def row_shard(num_rows: int, rank: int, world: int):
    return rank * num_rows // world, (rank + 1) * num_rows // world


_TABLE_KEYS_STD = ("left", "right", "feature", "threshold", "num_instances")
_TABLE_KEYS_EXT = ("left", "right", "num_instances", "offset")


def merge_tables(shards: list) -> dict:
    """Concatenate forest-table shards (in rank order) into one forest."""
    first = shards[0]
    ext = bool(first["extended"])
    out = {k: first[k] for k in ("extended", "num_samples", "total_num_features") if k in first}
    out["num_trees"] = int(sum(int(s["num_trees"]) for s in shards))
    offs, base = [np.zeros(1, np.int32)], 0
    for s in shards:
        offs.append((np.asarray(s["node_off"][1:], np.int64) + base).astype(np.int32))
        base += int(s["node_off"][-1])
    out["node_off"] = np.concatenate(offs)
    for k in (_TABLE_KEYS_EXT if ext else _TABLE_KEYS_STD):
        out[k] = np.concatenate([np.asarray(s[k]) for s in shards])
    if ext:
        hoffs, hbase = [np.zeros(1, np.int64)], 0
        for s in shards:
            hoffs.append(np.asarray(s["hp_off"][1:], np.int64) + hbase)
            hbase += int(s["hp_off"][-1])
        out["hp_off"] = np.concatenate(hoffs)
        out["hp_idx"] = np.concatenate([np.asarray(s["hp_idx"]) for s in shards])
        out["hp_w"] = np.concatenate([np.asarray(s["hp_w"]) for s in shards])
    return out


def gather_tables(local_tables: dict, group=None) -> dict:
    """all_gather of the per-rank table shards (KBs..MBs) -> the full forest on every rank."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    shards = [None] * world
    dist.all_gather_object(shards, local_tables, group=group)
    return merge_tables(shards)


def fit_tree_sharded(X, fit_params, group=None):
    """Each rank builds trees [t0, t1) of the ensemble on its own GPU; returns (local NativeForest, full tables)."""
    import torch.distributed as dist

    from . import _native as nat

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t0, t1 = tree_shard(fit_params.num_estimators, rank, world)
    p = nat.FitParams(fit_params.num_estimators, fit_params.num_samples, fit_params.num_features, fit_params.bootstrap,
                      fit_params.random_seed, fit_params.num_partitions, fit_params.extension_level, t0, t1)
    local = nat.fit_device(X, p)
    return local, gather_tables(local.export(), group)


def score_tree_sharded(local_forest, X, total_num_trees: int, num_samples: int, group=None, want_depth=False):
    """Tree-sharded transform: partial sums -> all_reduce -> scores (identical on every rank)."""
    import torch
    import torch.distributed as dist

    from . import _native as nat

    n = X.shape[0]
    psum = torch.zeros(n, dtype=torch.float32, device=X.device)
    dsum = torch.zeros(n, dtype=torch.int32, device=X.device) if want_depth else None
    local_forest.score_partial_device(X, psum, dsum)
    dist.all_reduce(psum, group=group)
    if want_depth:
        dist.all_reduce(dsum, group=group)
    scores = nat.finalize_scores_device(psum, total_num_trees, num_samples)
    return (scores, dsum, psum) if want_depth else scores


def hybrid_groups(world: int, tree_shards: int):
    """Hybrid rows x trees layout: ranks [g*S, (g+1)*S) form row group g and split the ensemble S ways.

    Every rank must call this (torch.distributed.new_group is collective); returns (my group, row group index,
    number of row groups).  S == world is pure tree sharding, S == 1 is the collective-free row sharding."""
    import torch.distributed as dist

    assert world % tree_shards == 0, (world, tree_shards)
    rank = dist.get_rank()
    mine = None
    for g in range(world // tree_shards):
        grp = dist.new_group(list(range(g * tree_shards, (g + 1) * tree_shards)))
        if rank // tree_shards == g:
            mine = grp
    return mine, rank // tree_shards, world // tree_shards


def score_tree_sharded_rs(local_forest, X, total_num_trees: int, num_samples: int, group=None, psum=None, out=None,
                          scores=None):
    """Tree-sharded transform with a REDUCE-SCATTER: every rank of `group` ends up with the scores of its own
    contiguous row slice only (rows are padded to a multiple of the group size).  Returns (scores_local, r0, r1)."""
    import torch
    import torch.distributed as dist

    from . import _native as nat

    n = X.shape[0]
    S, r = dist.get_world_size(group), dist.get_rank(group)
    per = (n + S - 1) // S
    if psum is None:
        psum = torch.empty(per * S, dtype=torch.float32, device=X.device)
    psum.zero_()
    local_forest.score_partial_device(X, psum[:n])
    if out is None:
        out = torch.empty(per, dtype=torch.float32, device=X.device)
    dist.reduce_scatter_tensor(out, psum, group=group)
    r0, r1 = min(n, r * per), min(n, (r + 1) * per)
    scores = nat.finalize_scores_device(out[: r1 - r0], total_num_trees, num_samples, scores=scores)
    return scores, r0, r1


class ScatterContext:
    """Peer-memory buffers for the FUSED tree-sharded transform (include/ifb200.h: ifb_score_scatter_device).

    Rank o owns rows [cut[o], cut[o+1]) and exposes, through CUDA IPC, TWO buffers of world * rows_o floats
    (alternated from call to call) plus `world` 32-bit flags; every rank maps every peer's allocation once.
    A transform is then: one scoring kernel per rank whose epilogue stores each row's partial path-length sum into
    the owner's buffer over NVLink, a device-side flag barrier (system-scope release/acquire on the peers' flags),
    and a rank-ordered sum + score epilogue on the owner.  No collective library on the data path.

    Why two buffers: rank B may start step k+1's scatter while rank A still reads step k's partials; with the
    buffers alternating, the next write into a buffer (step k+2) happens after every rank has passed step k+1's
    barrier, i.e. after every rank finished reading step k."""

    def __init__(self, n_rows: int, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _native as nat

        self.nat, self.group = nat, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.cuda.current_device()
        self.n_rows = n_rows
        self.cuts = [row_shard(n_rows, r, self.world)[0] for r in range(self.world)] + [n_rows]
        self.rows_local = self.cuts[self.rank + 1] - self.cuts[self.rank]
        rows = [self.cuts[o + 1] - self.cuts[o] for o in range(self.world)]
        # allocation layout of rank o: [flags: 256 B][buffer 0: world*rows_o floats][buffer 1: same]
        self._buf_bytes = [max(16, (self.world * r * 4 + 255) & ~255) for r in rows]
        p = C.c_void_p()
        nat.check(nat.lib().ifb_device_alloc(self.device, 256 + 2 * self._buf_bytes[self.rank], C.byref(p)))
        self.local_ptr = p
        torch.cuda.synchronize()
        zero = torch.zeros(64, dtype=torch.int32)
        nat.check(nat.lib().ifb_copy_to_device(self.device, p, C.c_void_p(zero.data_ptr()), 256))
        handle = C.create_string_buffer(64)
        nat.check(nat.lib().ifb_ipc_export(self.device, p, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.bases = []
        for o in range(self.world):
            if o == self.rank:
                self.bases.append(p.value)
            else:
                q = C.c_void_p()
                nat.check(nat.lib().ifb_ipc_open(self.device, C.create_string_buffer(handles[o], 64), C.byref(q)))
                self.bases.append(q.value)
        self._flag_arr = (C.c_void_p * self.world)(*self.bases)
        self._peer_arr = [(C.c_void_p * self.world)(*[self.bases[o] + 256 + b * self._buf_bytes[o] for o in range(self.world)])
                          for b in (0, 1)]
        self._cut_arr = (C.c_int64 * (self.world + 1))(*self.cuts)
        self.epoch = 0
        dist.barrier(group=group)   # every rank has zeroed its flags before anyone signals

    def score(self, local_forest, X, total_num_trees: int, num_samples: int, scores_local=None):
        """Returns this rank's slice of the scores (rows cut[rank] .. cut[rank+1])."""
        import ctypes as C

        import torch

        nat = self.nat
        n, d, ld, layout = nat.NativeForest._layout_of(tuple(X.shape), tuple(X.stride()))
        assert n == self.n_rows and layout == nat.COL_MAJOR
        self.epoch += 1
        b = self.epoch & 1
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nat.check(nat.lib().ifb_score_scatter_device(local_forest.handle, C.c_void_p(X.data_ptr()), n, d, ld, layout,
                                                     self.world, self.rank, self._cut_arr, self._peer_arr[b], st))
        nat.check(nat.lib().ifb_peer_signal_device(self.device, self.world, self.rank, self._flag_arr, self.epoch, st))
        nat.check(nat.lib().ifb_peer_wait_device(self.device, self.world, C.c_void_p(self.bases[self.rank]), self.epoch, st))
        if scores_local is None:
            scores_local = torch.empty(self.rows_local, dtype=torch.float64, device=X.device)
        local_buf = C.c_void_p(self.bases[self.rank] + 256 + b * self._buf_bytes[self.rank])
        nat.check(nat.lib().ifb_finalize_gathered_device(self.device, local_buf, self.world, self.rows_local,
                                                         total_num_trees, num_samples,
                                                         C.c_void_p(scores_local.data_ptr()), st))
        return scores_local

    def close(self):
        import ctypes as C

        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for o, q in enumerate(self.bases):
            if o != self.rank:
                self.nat.lib().ifb_ipc_close(self.device, C.c_void_p(q))
        dist.barrier(group=self.group)
        self.nat.lib().ifb_device_free(self.device, self.local_ptr)
        self.bases = []
