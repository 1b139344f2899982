"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: shard arithmetic, table gathering, and the
all-reduce of per-row partial path sums.  The per-shard partial sums come from the oracle here (no GPU); the
GPU kernels' partial sums are checked against the same oracle in tests/test_score_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, synth_mixture


def test_shard_arithmetic(pkg):
    from isolation_forest_b200 import distributed as D

    for T, W in ((100, 8), (512, 8), (7, 3), (3, 4), (1, 2)):
        cuts = [D.tree_shard(T, r, W) for r in range(W)]
        assert cuts[0][0] == 0 and cuts[-1][1] == T
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
        assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    assert [D.row_shard(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]


def _slice(tables, t0, t1):
    nb, ne = tables["node_off"][t0], tables["node_off"][t1]
    s = dict(tables)
    s.update(num_trees=t1 - t0, node_off=(tables["node_off"][t0:t1 + 1] - nb).astype(np.int32))
    keys = ["left", "right", "num_instances"] + (["offset"] if tables["extended"] else ["feature", "threshold"])
    for k in keys:
        s[k] = tables[k][nb:ne]
    if tables["extended"]:
        hb, he = tables["hp_off"][nb], tables["hp_off"][ne]
        s["hp_off"] = tables["hp_off"][nb:ne + 1] - hb
        s["hp_idx"], s["hp_w"] = tables["hp_idx"][hb:he], tables["hp_w"][hb:he]
    return s


@pytest.mark.parametrize("ext", [-1, 5])
def test_merge_tables_roundtrip(pkg, oracle, ext):
    from isolation_forest_b200 import distributed as D

    X = synth_mixture(3000, 6, 3)
    t = oracle.fit_forest(X, 17, 128, random_seed=2, ext_level=ext)
    parts = [_slice(t, *D.tree_shard(17, r, 3)) for r in range(3)]
    m = D.merge_tables(parts)
    for k in t:
        if isinstance(t[k], np.ndarray):
            assert np.array_equal(m[k], t[k]), k
    assert m["num_trees"] == 17


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.load_package()
    O = g.load_oracle()
    from isolation_forest_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = synth_mixture(4000, 8, 17)
        T, ns = 21, 256
        full = O.fit_forest(X, T, ns, random_seed=5)
        t0, t1 = D.tree_shard(T, rank, world)
        local = _slice(full, t0, t1)                       # what fit_tree_sharded's local build would hold
        merged = D.gather_tables(local)                    # all_gather_object over gloo
        for k in full:
            if isinstance(full[k], np.ndarray):
                assert np.array_equal(merged[k], full[k]), k
        # tree-sharded transform: per-rank partial sums (oracle), all-reduce, epilogue
        _, dsum, psum = O.Forest(dict(local, num_samples=ns)).score(X, want_parts=True)
        tp, td = torch.from_numpy(psum.copy()), torch.from_numpy(dsum.copy())
        dist.all_reduce(tp)
        dist.all_reduce(td)
        ref, rd, rp = O.Forest(full).score(X, want_parts=True)
        assert np.array_equal(td.numpy(), rd)              # integer part: exact and order-free
        e = tp.numpy() / np.float32(T)
        z = -e / O.avg_path_length(ns)
        scores = np.power(2.0, z.astype(np.float64))
        assert np.max(np.abs(scores - ref) / ref) < 1e-6   # f32 sum order differs across shards: <= ~1e-7 rel
        # row-sharded: no collective, bit-identical rows
        r0, r1 = D.row_shard(len(X), rank, world)
        assert np.array_equal(O.Forest(full).score(X[r0:r1]), ref[r0:r1])
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world2_tree_sharding(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def _hybrid_worker(rank, world, port, tmp):
    """Hybrid rows x trees layout on 4 gloo ranks: 2 row groups x 2 tree shards; the reduce-scatter inside a group is
    emulated with the group's all-reduce (gloo has no reduce_scatter_tensor) and the slice arithmetic of
    distributed.score_tree_sharded_rs."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.load_package()
    O = g.load_oracle()
    from isolation_forest_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = 2
        group, gidx, G = D.hybrid_groups(world, S)
        assert (gidx, G) == (rank // S, world // S) and dist.get_world_size(group) == S
        assert dist.get_rank(group) == rank % S
        X = synth_mixture(3001, 8, 23)                     # odd row count: slices are padded
        T, ns = 19, 256
        full = O.fit_forest(X, T, ns, random_seed=7)
        ref = O.Forest(full).score(X)
        g0, g1 = D.row_shard(len(X), gidx, G)              # rows of my row group
        t0, t1 = D.tree_shard(T, rank % S, S)              # my slice of the ensemble
        _, _, psum = O.Forest(dict(_slice(full, t0, t1), num_samples=ns)).score(X[g0:g1], want_parts=True)
        ng = g1 - g0
        per = (ng + S - 1) // S
        buf = torch.zeros(per * S)
        buf[:ng] = torch.from_numpy(psum.copy())
        dist.all_reduce(buf, group=group)                  # sum over the group's tree shards only
        r = rank % S
        lo, hi = min(ng, r * per), min(ng, (r + 1) * per)  # my slice of the group's rows
        e = buf[lo:hi].numpy() / np.float32(T)
        sc = np.power(2.0, (-e / O.avg_path_length(ns)).astype(np.float64))
        assert np.max(np.abs(sc - ref[g0 + lo:g0 + hi]) / ref[g0 + lo:g0 + hi]) < 1e-6
        covered = torch.zeros(len(X))
        covered[g0 + lo:g0 + hi] = 1
        dist.all_reduce(covered)                           # every row is finalised by exactly one rank
        assert torch.all(covered == 1)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gloo_world4_hybrid_rows_x_trees(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_hybrid_worker, args=(4, port, str(tmp_path)), nprocs=4, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1", "ok2", "ok3"]
