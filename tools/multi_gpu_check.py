"""Run under torchrun on >=2 GPUs: tree-sharded fit+transform and row-sharded transform vs one-GPU results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
nat = pkg._native
from isolation_forest_b200 import distributed as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = None
for ext, d, T, n in ((-1, 32, 100, 2_000_000), (15, 16, 40, 300_000), (63, 64, 24, 200_000)):
    gen = torch.Generator(device=dev).manual_seed(99)
    X = torch.randn(d, n, device=dev, generator=gen).t()            # identical on every rank
    prm = nat.FitParams(T, 256, d, 0, 1, 1, ext, 0, 0)
    local_forest, tables = D.fit_tree_sharded(X, prm)
    full = nat.NativeForest.from_tables(dict(tables, num_samples=256, total_num_features=d), device=local)
    whole = nat.fit_device(X, prm).export()                          # the un-sharded build on this GPU
    for k in ("node_off", "left", "right", "num_instances"):
        assert np.array_equal(tables[k], whole[k]), k                # shards reassemble the same forest
    s_ref, d_ref, p_ref = full.score_device(X, want_parts=True)
    s_sh, d_sh, p_sh = D.score_tree_sharded(local_forest, X, T, 256, want_depth=True)
    assert torch.equal(d_sh, d_ref), "depth sums must be exact under tree sharding"
    rel = float(((s_sh - s_ref).abs() / s_ref).max())
    assert rel < 1e-6, rel
    r0, r1 = D.row_shard(n, rank, world)
    if ext < 0:
        ctx = D.ScatterContext(n)
        s_fused = ctx.score(local_forest, X, T, 256)
        torch.cuda.synchronize()
        relf = float(((s_fused - s_ref[r0:r1]).abs() / s_ref[r0:r1]).max())
        assert relf < 1e-6, relf
        s_fused2 = ctx.score(local_forest, X, T, 256)          # rank-ordered sums: bitwise reproducible
        assert torch.equal(s_fused, s_fused2)
        ctx.close()
        if rank == 0:
            print(f"fused scatter ok: max rel {relf:.2e}", flush=True)
    # the same layout with the collective inside libifb200.so (ifb_comm_init / ifb_score_sharded: what a JVM would bind)
    if comm is None:
        uid = [nat.NativeComm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = nat.NativeComm(local, world, rank, uid[0])
    s_lib, b0, b1 = comm.score_sharded(local_forest, X, T, nat.SHARD_ALLREDUCE)
    torch.cuda.synchronize()
    rel_ar = float(((s_lib - s_ref).abs() / s_ref).max())
    assert (b0, b1) == (0, n) and rel_ar < 1e-6, rel_ar
    s_rs, b0, b1 = comm.score_sharded(local_forest, X, T, nat.SHARD_REDUCE_SCATTER)
    torch.cuda.synchronize()
    rel_rs = float(((s_rs - s_ref[b0:b1]).abs() / s_ref[b0:b1]).max())
    assert rel_rs < 1e-6, rel_rs
    if rank == 0:
        print(f"ifb_score_sharded ok: all-reduce max rel {rel_ar:.2e}, reduce-scatter slice [{b0},{b1}) max rel {rel_rs:.2e}",
              flush=True)
    # hybrid rows x trees: groups of 2 tree shards, each group owns a contiguous row range
    if world % 2 == 0 and world > 2:
        grp, g, G = D.hybrid_groups(world, 2)
        g0, g1 = D.row_shard(n, g, G)
        t0, t1 = D.tree_shard(T, rank % 2, 2)
        lf = nat.fit_device(X, nat.FitParams(T, 256, d, 0, 1, 1, ext, t0, t1))
        s_h, h0, h1 = D.score_tree_sharded_rs(lf, X[g0:g1], T, 256, group=grp)
        torch.cuda.synchronize()
        rel_h = float(((s_h - s_ref[g0 + h0:g0 + h1]).abs() / s_ref[g0 + h0:g0 + h1]).max())
        assert rel_h < 1e-6, rel_h
        if rank == 0:
            print(f"hybrid {G}x2 reduce-scatter ok: max rel {rel_h:.2e}", flush=True)
    s_rows = full.score_device(X[r0:r1])
    assert torch.equal(s_rows, s_ref[r0:r1]), "row sharding must be bit-identical"
    if rank == 0:
        print(f"multi-gpu check ok: world={world} ext={ext} d={d} T={T} n={n} tree-shard max rel {rel:.2e}", flush=True)
if comm is not None:
    comm.close()
dist.destroy_process_group()
