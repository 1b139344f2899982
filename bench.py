#!/usr/bin/env python
"""bench.py -- rows scored/sec of the isolation-forest scoring hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3             # native arm (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3        # reference arm: CPU port on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of IsolationForestModel.transform's hot path over one batch of synthetic rows:
BASELINE.json configs[1] = 10M x 32 f32 rows, 100 trees, maxSamples 256, standard forest (per GPU).

Data contract: rows come from the counter-based generator in synthdata.py, so the CPU oracle, the GPU
and BOTH bench arms see bit-identical rows for any row subset; the forest is fitted with seed 1 on the same training
matrix on both sides (the GPU builder and the oracle builder produce bit-identical tables, tests/test_fit_gpu.py), and
both arms print the forest's checksum.

Printed (rank 0, ONE JSON line):
  value        whole-job rows/s with inputs resident in HBM (CUDA events, max over ranks), rows sharded over the ranks
               with the forest replicated (what the reference does; no data-path collective => weak scaling)
  e2e          the same through ifb_score_host with pinned HOST buffers (H2D + kernels + D2H inside the timed region)
  parity       the scores the timed configuration produces against the CPU oracle on the SAME rows and forest
               (run fails with rc != 0 when it is off)
  roofline     algorithmic bytes N*(4d+8) per launch / event time vs the measured HBM peak
  cpu_baseline the CPU oracle port on a bounded sample on this box's host cores
  tree_sharded (N > 1) the layout BASELINE's north star names -- numEstimators split over the GPUs, per-row path sums
               reduced over NVLink -- on rank 0's rows (strong scaling): NCCL all-reduce, NCCL reduce-scatter, the
               fused peer-memory scatter, and the hybrid rows x trees layouts, each with its own parity check
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402
import synthdata  # noqa: E402

WORKLOADS = {
    # name: (rows per GPU, features, trees, numSamples, extensionLevel)
    "config2": (10_000_000, 32, 100, 256, -1),
    "config3": (10_000_000, 64, 200, 256, 63),
    "config1": (1_000, 10, 100, 256, -1),
    "config5": (1_000_000, 1024, 256, 256, 1023),   # BASELINE configs[4] per-GPU shape (named there on 4 GPUs)
    # BASELINE configs[3]: fit+transform, every step builds the forest on the step's rows and scores them
    "config4": (100_000_000, 128, 512, 256, -1),
}
SEEDS = {"config1": 1001, "config2": 1002, "config3": 1003, "config4": 1004, "config5": 1005}
TRAIN_SEED = 4242
FIT_IN_STEP = {"config4"}
TRAIN_ROWS = 1 << 20   # rows of the (rank-independent) training matrix the forest is fitted on
FALLBACK_HBM_GBS = 6650.0


def train_rows_for(n, d, ns):
    rows = TRAIN_ROWS if n >= TRAIN_ROWS else max(n, ns)
    if d >= 512:
        rows = min(rows, 1 << 17)   # the builder only samples numEstimators * numSamples rows anyway
    return rows


def workload_label(wl_name, n, d, T, ns, ext):
    """config.workload, identical for the native and the reference arm."""
    return (f"{wl_name}: " + ("IsolationForest.fit + " if wl_name in FIT_IN_STEP else "") +
            f"IsolationForestModel.transform {n}x{d} f32 per GPU, {T} trees, maxSamples={ns}" +
            (f", extensionLevel={ext}" if ext >= 0 else ""))


def forest_sha(tables):
    h = hashlib.sha1()
    for k in ("node_off", "left", "right", "num_instances", "feature", "threshold", "offset", "hp_off", "hp_idx", "hp_w"):
        if k in tables and tables[k] is not None:
            h.update(np.ascontiguousarray(tables[k]).tobytes())
    return h.hexdigest()[:16]


def data_note(wl_name, d):
    return {"generator": "synthdata.py: counter-based splitmix64, Irwin-Hall(8) Gaussian stand-in, BASELINE mixture "
                         "49/49/2 %; identical bits on CPU and GPU", "rows_seed": SEEDS[wl_name], "train_seed": TRAIN_SEED}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(ln) for ln in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def usable_cores():
    """(threads to use, note): CPU affinity and cgroup CPU quota rather than the raw logical-CPU count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    note = f"{n} logical CPUs in the affinity mask"
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            q = float(quota) / float(period)
            note += f", cgroup quota {q:.1f} CPUs (threads sized to the quota)"
            n = min(n, max(1, int(np.ceil(q))))
    except Exception:
        pass
    return max(1, n), note


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per launch from the committed ncu capture of this workload (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


def parity_of(got, ref, what):
    """got/ref = (scores f64, depth sums i32 or None, path sums f32 or None) on the same rows."""
    gs, gd, gp = got
    rs, rd, rp = ref
    out = {"rows": int(len(rs)), "against": what, "max_rel": float(np.max(np.abs(gs - rs) / rs)) if len(rs) else 0.0}
    if gd is not None and rd is not None:
        out["depth_sums_exact"] = bool(np.array_equal(gd, rd))
    if gp is not None and rp is not None:
        out["path_sums_exact"] = bool(np.array_equal(gp, rp))
    out["ok"] = bool(out["max_rel"] <= 1e-5 and out.get("depth_sums_exact", True))
    return out


def run_reference(args, wl_name, wl):
    """Reference arm: the reference's CPU algorithm (oracle port; no JVM/Spark in this image) on host cores,
    on a prefix of the SAME rows and the SAME forest as the native arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d, T, ns, ext = wl
    if args.rows:
        n = args.rows
    O = graft.load_oracle()
    cores, cores_note = usable_cores()
    train = synthdata.matrix_numpy(train_rows_for(n, d, ns), d, TRAIN_SEED)
    tables = O.fit_forest(train, T, ns, random_seed=1, ext_level=ext)
    del train
    forest = O.Forest(tables)
    # bounded sample per step: the whole --steps/--warmup run is sized to about 75 s of CPU work
    probe_rows = min(n, 20_000 if d >= 512 else 100_000)
    Xp = synthdata.matrix_numpy(probe_rows, d, SEEDS[wl_name])
    forest.score(Xp, threads=cores)
    t0 = time.perf_counter(); forest.score(Xp, threads=cores); rate = probe_rows / (time.perf_counter() - t0)
    per_step_s = min(4.0, 75.0 / max(1, args.steps + args.warmup))
    rows = int(min(n, max(20_000, rate * per_step_s)))
    X = Xp[:rows] if rows <= probe_rows else synthdata.matrix_numpy(rows, d, SEEDS[wl_name])
    for _ in range(args.warmup):
        forest.score(X, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        forest.score(X, threads=cores)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    line = {
        "impl": "reference", "metric": "rows scored/sec", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 features / f64 thresholds+scores", "data": "synthetic",
        "config": {"workload": workload_label(wl_name, n, d, T, ns, ext),
                   "note": "no JVM/Spark in this image: C port of the reference algorithm (oracle/ifb_oracle.c), "
                           "pthreads over all host cores, bounded sample per step (transform only)",
                   "data": data_note(wl_name, d),
                   "forest": {"nodes": int(tables["node_off"][-1]), "sha": forest_sha(tables),
                              "fit": "oracle CPU builder, seed 1 (bit-identical to the GPU builder's tables)"},
                   "rows": f"first {rows} rows of the native arm's rank-0 matrix"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{rows} rows x {d} features per step", "cores_note": cores_note},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_native(args, wl_name, wl):
    import torch
    import torch.distributed as dist

    n, d, T, ns, ext = wl
    if args.rows:
        n = args.rows
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = graft.load_package()
    nat = pkg._native
    fit_in_step = wl_name in FIT_IN_STEP
    seed = SEEDS[wl_name]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warm):
        """(ms per step, max over ranks) of step_fn, CUDA events around `steps` calls between barriers."""
        for _ in range(warm):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---- setup (untimed): forest from the product's own GPU fit on a rank-independent training matrix ----
    tr = train_rows_for(n, d, ns)
    train = synthdata.matrix_torch(torch, tr, d, TRAIN_SEED, dev)
    prm = nat.FitParams(T, ns, d, 0, 1, 1, ext, 0, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    forest = nat.fit_device(train, prm)
    torch.cuda.synchronize()
    fit_ms = (time.perf_counter() - t0) * 1e3
    del train
    X = synthdata.matrix_torch(torch, n, d, seed, dev, row0=rank * n)   # rows [rank*n, (rank+1)*n) of ONE global matrix
    holder = {"forest": forest}
    phase_events = []
    scores = torch.empty(n, dtype=torch.float64, device=dev)

    def step():
        if fit_in_step:     # Estimator.fit on this step's rows, then transform
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            holder["forest"] = nat.fit_device(X, prm)
            ev[1].record()
            phase_events.append(ev)
        holder["forest"].score_device(X, scores=scores)
        if fit_in_step:
            ev[2].record()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nat.kernel_launch_count(reset=True)
    phase_events.clear()
    ms_step = timed(step, args.steps, 0)
    launches = nat.kernel_launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ms_step * args.steps
    phases = None
    if phase_events:
        phases = {"fit_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in phase_events])),
                  "transform_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in phase_events]))}
    value = n * world * args.steps / (ms_total / 1e3)

    # ---- e2e: the call a Spark task would make: host buffers in, host scores out ----------------------
    e2e = None
    if n * d * 4 <= (8 << 30):
        hx = nat.PinnedBuffer((d, n), np.float32)              # column-major rows x features
        hs = nat.PinnedBuffer((n,), np.float64)
        torch.from_numpy(hx.array).copy_(X.t())                # fill the pinned staging buffer (untimed)
        torch.cuda.synchronize()
        import ctypes as C
        args_host = (holder["forest"].handle, C.c_void_p(hx.array.ctypes.data), n, d, n, nat.COL_MAJOR,
                     C.c_void_p(hs.array.ctypes.data), None, None)
        for _ in range(2):
            nat.check(nat.lib().ifb_score_host(*args_host))
        esteps = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            nat.check(nat.lib().ifb_score_host(*args_host))
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * esteps / float(tt.item()), "unit": "rows/s", "h2d_bytes_per_step": n * d * 4,
               "d2h_bytes_per_step": n * 8, "steps": esteps,
               "path": "ifb_score_host: pinned host col-major f32 -> 3-stream chunked H2D/score/D2H -> host f64"}
        e2e["matches_device_path"] = bool(np.array_equal(hs.array, scores.cpu().numpy()))
        # what bounds e2e: the bare pinned-host -> device copy rate of this rank while every rank copies at once
        # (min / max over ranks; a drop from N = 1 names the shared host-side path -- root complex / host memory -- not the GPU)
        hview = torch.from_numpy(hx.array)
        dbuf = torch.empty_like(X.t())
        dbuf.copy_(hview, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            dbuf.copy_(hview, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        gbs = 3 * n * d * 4 / (c0.elapsed_time(c1) / 1e3) / 1e9
        lo = torch.tensor([gbs], dtype=torch.float64, device=dev)
        hi = lo.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        e2e["h2d_copy_gbs_per_rank"] = {"min": float(lo.item()), "max": float(hi.item()),
                                        "note": "bare cudaMemcpyAsync of the same pinned matrix, all ranks concurrently"}
        del dbuf, hview
        hx.free(); hs.free()

    # ---- parity of what was just timed, against the CPU oracle on the same rows and the same forest (rank 0) ----
    forest = holder["forest"]
    parity = cpu_baseline = None
    tables = forest.export() if rank == 0 else None
    ref_full = None
    if rank == 0 and not args.no_cpu:
        O = graft.load_oracle()
        cores, cores_note = usable_cores()
        oforest = O.Forest(tables)
        got = [t.cpu().numpy() for t in forest.score_device(X, want_parts=True)]   # once, outside the timing
        assert np.array_equal(got[0], scores.cpu().numpy()), "want_parts launch and timed launch disagree"
        # (a) a strided sample over the whole matrix, regenerated on the CPU from the counters
        ns_rows = min(n, 1 << 14)
        idx = (np.arange(ns_rows, dtype=np.int64) * (n // ns_rows)) if n >= ns_rows else np.arange(n)
        Xs = synthdata.rows_numpy(idx.astype(np.uint64), d, seed)
        assert np.array_equal(Xs[:64], X[torch.from_numpy(idx[:64]).to(dev)].cpu().numpy()), \
            "CPU and GPU generators disagree"
        ref_s = oforest.score(Xs, threads=cores, want_parts=True)
        parity = parity_of([g[idx] for g in got], ref_s, "oracle on a strided sample regenerated on the CPU")
        # (b) the bounded prefix the CPU baseline is timed on
        if world == 1:
            target = 12.0
            probe = min(n, 20_000 if d >= 512 else 50_000)
            Xp = np.ascontiguousarray(X[:probe].cpu().numpy())
            t0 = time.perf_counter(); oforest.score(Xp, threads=cores); rate = probe / max(time.perf_counter() - t0, 1e-6)
            rows = int(min(n, max(probe, rate * target), 4_000_000 if d <= 64 else 200_000))
            Xc = np.ascontiguousarray(X[:rows].cpu().numpy())
            t0 = time.perf_counter()
            ref_c = oforest.score(Xc, threads=cores, want_parts=True)
            dtc = time.perf_counter() - t0
            cpu_baseline = {"value": rows / dtc, "unit": "rows/s", "cores": cores, "kind": "port",
                            "sample": f"first {rows} rows of the same matrix, same forest, {dtc:.1f} s",
                            "cores_note": cores_note}
            p2 = parity_of([g[:rows] for g in got], ref_c, "oracle on the cpu_baseline prefix")
            parity = {"rows": parity["rows"] + p2["rows"], "against": "CPU oracle: strided sample + cpu_baseline prefix, "
                      "same rows, same forest", "max_rel": max(parity["max_rel"], p2["max_rel"]),
                      "depth_sums_exact": parity["depth_sums_exact"] and p2["depth_sums_exact"],
                      "path_sums_exact": parity["path_sums_exact"] and p2["path_sums_exact"],
                      "ok": parity["ok"] and p2["ok"]}
        ref_full = got
        del got

    # ---- N > 1: the tree-sharded layouts on rank 0's rows (strong scaling) ----
    tree_sharded = None
    if world > 1 and not args.no_tree_sharded:
        tree_sharded = run_tree_sharded(args, torch, dist, pkg, nat, dev, rank, world, wl_name, (n, d, T, ns, ext), X,
                                        scores, timed, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    kernel_ms = ms_total / args.steps                    # one step == one launch of the dominant kernel
    alg_bytes = n * (4 * d + 8)
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    info = forest.info()
    kp, ncols = forest.ext_tc_info() if ext >= 0 else (0, 0)
    kernel_name = ("score_ext_tc_kernel (tcgen05)" if ncols else "score_ext_*") if ext >= 0 else "score_std_kernel"
    line = {
        "metric": "rows scored/sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 features, f32 path sums, f64 scores" + (", f64 hyperplane dots (fp16x2-split tcgen05 filter)" if ext >= 0 else ""),
        "data": "synthetic",
        "config": {"workload": workload_label(wl_name, n, d, T, ns, ext),
                   "parallelism": f"rows sharded x{world}, forest replicated, no data-path collective",
                   "l2": f"inputs ({n * d * 4 / 1e9:.2f} GB/GPU) larger than L2; no flush needed",
                   "data": data_note(wl_name, d),
                   "forest": {"nodes": int(info.num_nodes), "max_depth": int(info.max_depth), "sha": forest_sha(tables),
                              "fit": "ifb_fit_device (this repo's GPU builder), seed 1", "fit_ms": round(fit_ms, 2)}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(wl_name), "traffic_source": "ncu capture committed under profiles/ (constant, "
                     "not measured in this run)" if ncu_traffic(wl_name) else None, "peak_source": peak_src,
                     "kernel": kernel_name, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                     "note": "per GPU; one step = one launch of the dominant kernel"
                             + (" (+ the row-preparation launch of the tensor-core path)" if ncols else "")
                             + (" (+ the fit launch and its host-side table assembly)" if fit_in_step else "")},
    }
    if phases:
        line["phases"] = phases
        line["roofline"].update(kernel_ms=phases["transform_ms"],
                                achieved=alg_bytes / (phases["transform_ms"] / 1e3) / 1e9,
                                frac=alg_bytes / (phases["transform_ms"] / 1e3) / 1e9 / peak)
    if e2e:
        line["e2e"] = e2e
    if parity:
        line["parity"] = parity
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if tree_sharded:
        line["tree_sharded"] = tree_sharded
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    bad = (parity is not None and not parity["ok"]) or (e2e is not None and not e2e["matches_device_path"])
    if tree_sharded:
        bad = bad or any(isinstance(v, dict) and v.get("parity_ok") is False for v in tree_sharded.values())
    if bad:
        print("bench.py: PARITY FAILURE (see the parity / tree_sharded objects)", file=sys.stderr)
        sys.exit(3)


def run_tree_sharded(args, torch, dist, pkg, nat, dev, rank, world, wl_name, wl, X, scores0, timed, barrier):
    """Strong-scaling measurements of the tree-sharded layouts on rank 0's rows; every variant is compared with rank
    0's single-GPU scores of the same rows (the rows-sharded result above).  Returns the `tree_sharded` object."""
    from isolation_forest_b200 import distributed as D

    n, d, T, ns, ext = wl
    fit_in_step = wl_name in FIT_IN_STEP
    seed = SEEDS[wl_name]
    steps = max(3, min(args.steps, 20))
    # reference: rank 0's rows-sharded scores of rows [0, n) -- broadcast so that every rank can check its own slice
    ref = scores0.clone()
    dist.broadcast(ref, src=0)
    if rank != 0:
        del X
        torch.cuda.empty_cache()
        X = synthdata.matrix_torch(torch, n, d, seed, dev, row0=0)   # every rank: the SAME rows as rank 0
    tr = train_rows_for(n, d, ns)
    train = None if fit_in_step else synthdata.matrix_torch(torch, tr, d, TRAIN_SEED, dev)
    out = {"rows": n, "steps": steps, "scaling": "strong",
           "note": "every rank scores rows [0, n) of rank 0 against its slice of the ensemble; rows/s = n / ms_per_step; "
                   "under pure tree sharding every GPU reads all rows, so the aggregate HBM fraction is capped at 1/G"}

    def rel_to_ref(sc, r0, r1):
        m = float(((sc - ref[r0:r1]).abs() / ref[r0:r1]).max()) if r1 > r0 else 0.0
        t = torch.tensor([m], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def variant(name, S, mode):
        """S tree shards per row group (S == world: pure tree sharding); mode: allreduce | reduce_scatter | fused."""
        group, g, G = D.hybrid_groups(world, S) if S != world else (None, 0, 1)
        gr = rank % S
        r0g, r1g = D.row_shard(n, g, G)                       # rows of my row group
        Xg = X[r0g:r1g]
        ng = r1g - r0g
        t_lo, t_hi = D.tree_shard(T, gr, S)
        prm = nat.FitParams(T, ns, d, 0, 1, 1, ext, t_lo, t_hi)
        hold = {"f": nat.fit_device(Xg if fit_in_step else train, prm)}
        per = (ng + S - 1) // S
        psum = torch.zeros(per * S, dtype=torch.float32, device=dev)
        part = torch.empty(per, dtype=torch.float32, device=dev)
        ctx = D.ScatterContext(ng, group=group) if mode == "fused" else None
        lr0, lr1 = (D.row_shard(ng, gr, S) if mode == "fused" else (min(ng, gr * per), min(ng, (gr + 1) * per)))
        sc_local = torch.empty(ng if mode == "allreduce" else lr1 - lr0, dtype=torch.float64, device=dev)
        ev = []

        def step():
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            if fit_in_step:
                hold["f"] = nat.fit_device(Xg, prm)
            e[1].record()
            if mode == "fused":
                ctx.score(hold["f"], Xg, T, ns, scores_local=sc_local)
            elif mode == "allreduce":
                psum.zero_()
                hold["f"].score_partial_device(Xg, psum[:ng])
                e[2].record()
                dist.all_reduce(psum, group=group)
                nat.finalize_scores_device(psum[:ng], T, ns, scores=sc_local)
            else:
                psum.zero_()
                hold["f"].score_partial_device(Xg, psum[:ng])
                e[2].record()
                dist.reduce_scatter_tensor(part, psum, group=group)
                nat.finalize_scores_device(part[: lr1 - lr0], T, ns, scores=sc_local)
            ev.append(e)

        ms = timed(step, steps, 3)
        ev = ev[-steps:]
        res = {"tree_shards": S, "row_groups": G, "mode": mode, "ms_per_step": ms, "rows_per_s": n / (ms / 1e3),
               "trees_per_gpu": t_hi - t_lo, "rows_per_gpu": ng}
        if fit_in_step:
            res["fit_ms"] = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
        if mode != "fused":
            torch.cuda.synchronize()
            res["partial_kernel_ms"] = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
            buf = psum if mode == "allreduce" else None

            def coll():
                if mode == "allreduce":
                    dist.all_reduce(buf, group=group)
                else:
                    dist.reduce_scatter_tensor(part, psum, group=group)
            res["collective_ms"] = timed(coll, steps, 2)
            res["collective_bytes"] = int(per * S * 4)
        # parity against rank 0's one-GPU scores of the same rows (f32 sum order differs: <= 1e-6 expected)
        if fit_in_step:
            res["parity"] = "not comparable: every step refits on the group's rows (checked by tools/multi_gpu_check.py)"
        else:
            if mode == "allreduce":
                rel = rel_to_ref(sc_local, r0g, r1g)
            else:
                rel = rel_to_ref(sc_local, r0g + lr0, r0g + lr1)
            res["max_rel_vs_one_gpu"] = rel
            res["parity_ok"] = bool(rel <= 1e-5)
            if mode == "allreduce" and S == world:
                ds = torch.zeros(ng, dtype=torch.int32, device=dev)
                ps = torch.zeros(ng, dtype=torch.float32, device=dev)
                hold["f"].score_partial_device(Xg, ps, ds)
                dist.all_reduce(ds, group=group)
                full = nat.fit_device(train, nat.FitParams(T, ns, d, 0, 1, 1, ext, 0, 0))
                _, d1, _ = full.score_device(Xg, want_parts=True)
                res["depth_sums_exact"] = bool(torch.equal(ds, d1))
                res["parity_ok"] = res["parity_ok"] and res["depth_sums_exact"]
        if ctx is not None:
            ctx.close()
        out[name] = res
        barrier()

    variant("nccl_allreduce", world, "allreduce")
    variant("nccl_reduce_scatter", world, "reduce_scatter")
    if ext < 0:
        variant("fused_scatter", world, "fused")
    for S in (4, 2):
        if S < world and world % S == 0:
            variant(f"hybrid_{world // S}x{S}_reduce_scatter", S, "reduce_scatter")
            if ext < 0:
                variant(f"hybrid_{world // S}x{S}_fused", S, "fused")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="config2")
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the oracle legs (parity + cpu_baseline)")
    ap.add_argument("--no-tree-sharded", action="store_true", help="N > 1: skip the tree-sharded layouts")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_native(args, args.workload, wl)


if __name__ == "__main__":
    main()
