#!/usr/bin/env python
"""bench.py -- rows scored/sec of the isolation-forest scoring hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 3             # native arm (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3        # reference arm: CPU port on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of IsolationForestModel.transform's hot path over one batch of synthetic rows:
BASELINE.json configs[1] = 10M x 32 f32 rows, 100 trees, maxSamples 256, standard forest (per GPU; with
N GPUs the rows are sharded, the forest is replicated, there is no data-path collective => weak scaling).

Printed (rank 0, ONE JSON line): value = whole-job rows/s with inputs resident in HBM (CUDA events, max over
ranks); e2e = the same through ifb_score_host with pinned HOST buffers (H2D + kernels + D2H inside the timed
region); roofline = algorithmic bytes N*(4d+8) per launch / event time vs the measured HBM peak;
cpu_baseline = the CPU oracle port on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
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

WORKLOADS = {
    # name: (rows per GPU, features, trees, numSamples, extensionLevel)
    "config2": (10_000_000, 32, 100, 256, -1),
    "config3": (10_000_000, 64, 200, 256, 63),
    "config1": (1_000, 10, 100, 256, -1),
    "config5": (1_000_000, 1024, 256, 256, 1023),   # BASELINE configs[4] per-GPU shape (named there on 4 GPUs)
    # BASELINE configs[3]: fit+transform, every step builds the forest on the step's rows and scores them
    # (use --shard trees / trees-fused under torchrun for the 8-GPU tree-sharded layout it names)
    "config4": (100_000_000, 128, 512, 256, -1),
}
FIT_IN_STEP = {"config4"}
TRAIN_ROWS = 1 << 20   # rows of the (rank-independent) training matrix the forest is fitted on
FALLBACK_HBM_GBS = 6650.0


def workload_label(wl_name, n, d, T, ns, ext):
    """config.workload, identical for the native and the reference arm."""
    return (f"{wl_name}: " + ("IsolationForest.fit + " if wl_name in FIT_IN_STEP else "") +
            f"IsolationForestModel.transform {n}x{d} f32 per GPU, {T} trees, maxSamples={ns}" +
            (f", extensionLevel={ext}" if ext >= 0 else ""))


def mixture_torch(torch, n, d, seed, device):
    """BASELINE's synthetic Gaussian mixture, generated on the device as a column-major (n x d) view."""
    g = torch.Generator(device=device).manual_seed(seed)
    xt = torch.randn(d, n, device=device, generator=g)
    z = torch.rand(n, device=device, generator=g)
    xt += ((z >= 0.49) & (z < 0.98)).to(xt.dtype) * (3.0 / np.sqrt(d))
    xt *= 1.0 + 3.0 * (z >= 0.98).to(xt.dtype)
    return xt.t()


def mixture_numpy(n, d, seed):
    rng = np.random.default_rng(seed)
    z = rng.random(n)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x[(z >= 0.49) & (z < 0.98)] += np.float32(3.0 / np.sqrt(d))
    x[z >= 0.98] *= np.float32(4.0)
    return x


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


def cpu_port_rate(O, forest, X, threads, target_s):
    """rows/s of the CPU oracle port on a bounded sample of X (calibrated to about target_s seconds)."""
    probe = min(len(X), 50_000)
    t0 = time.perf_counter()
    forest.score(X[:probe], threads=threads)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    rows = int(min(len(X), max(probe, rate * target_s)))
    t0 = time.perf_counter()
    forest.score(X[:rows], threads=threads)
    dt = time.perf_counter() - t0
    return rows / dt, rows, dt


def run_reference(args, wl_name, wl):
    """Reference arm: the reference's CPU algorithm (oracle port; no JVM/Spark in this image) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d, T, ns, ext = wl
    O = graft.load_oracle()
    cores, cores_note = usable_cores()
    train = mixture_numpy(min(TRAIN_ROWS, 1 << 18), d, 4242)
    tables = O.fit_forest(train, T, ns, random_seed=1, ext_level=ext)
    forest = O.Forest(tables)
    # bounded sample per step: the whole --steps/--warmup run is sized to about 75 s of CPU work
    probe_rows = min(n, 100_000)
    Xp = mixture_numpy(probe_rows, d, 1002)
    forest.score(Xp, threads=cores)
    t0 = time.perf_counter(); forest.score(Xp, threads=cores); rate = probe_rows / (time.perf_counter() - t0)
    per_step_s = min(4.0, 75.0 / max(1, args.steps + args.warmup))
    rows = int(min(n, max(20_000, rate * per_step_s)))
    X = mixture_numpy(rows, d, 1002)
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
                           "pthreads over all host cores, bounded sample per step (transform only)"},
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
    nat = graft.load_package()._native
    tree_sharded = args.shard in ("trees", "trees-fused") and world > 1
    fused = args.shard == "trees-fused" and world > 1

    # ---- setup (untimed): forest from the product's own GPU fit on a rank-independent training matrix ----
    train_rows = TRAIN_ROWS if n >= TRAIN_ROWS else max(n, ns)
    if d >= 512:
        train_rows = min(train_rows, 1 << 17)   # the builder only samples numEstimators * numSamples rows anyway
    train = mixture_torch(torch, train_rows, d, 4242, dev)
    t_lo, t_hi = (rank * T // world, (rank + 1) * T // world) if tree_sharded else (0, 0)
    prm = nat.FitParams(T, ns, d, 0, 1, 1, ext, t_lo, t_hi)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    forest = nat.fit_device(train, prm)
    torch.cuda.synchronize()
    fit_ms = (time.perf_counter() - t0) * 1e3
    del train
    X = mixture_torch(torch, n, d, 1002 + (0 if tree_sharded else rank), dev)
    fit_in_step = wl_name in FIT_IN_STEP
    holder = {"forest": forest}
    phase_events = []
    scores = torch.empty(n, dtype=torch.float64, device=dev)
    psum = torch.zeros(n, dtype=torch.float32, device=dev) if tree_sharded else None
    ctx = None
    if fused:
        from isolation_forest_b200 import distributed as D
        ctx = D.ScatterContext(n)
        scores = torch.empty(ctx.rows_local, dtype=torch.float64, device=dev)

    def step():
        forest = holder["forest"]
        if fit_in_step:     # Estimator.fit on this step's rows (own tree slice when trees are sharded), then transform
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            forest = holder["forest"] = nat.fit_device(X, prm)
            ev[1].record()
            phase_events.append(ev)
        if fused:
            ctx.score(forest, X, T, ns, scores_local=scores)   # kernel scatters partial sums into peer memory
        elif tree_sharded:
            psum.zero_()
            forest.score_partial_device(X, psum)
            dist.all_reduce(psum)                       # NCCL sum of per-row path-length sums over NVLink
            nat.finalize_scores_device(psum, T, ns, scores=scores)
        else:
            forest.score_device(X, scores=scores)
        if fit_in_step:
            ev[2].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    nat.kernel_launch_count(reset=True)
    phase_events.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = nat.kernel_launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    rows_job = n * (1 if tree_sharded else world)
    phases = None
    if phase_events:
        phases = {"fit_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in phase_events])),
                  "transform_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in phase_events]))}
    value = rows_job * args.steps / (ms_total / 1e3)

    # ---- e2e: the call a Spark task would make: host buffers in, host scores out ----------------------
    e2e = None
    if not tree_sharded and n * d * 4 <= (8 << 30):
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
        same = bool(np.array_equal(hs.array, scores.cpu().numpy()))
        e2e["matches_device_path"] = same
        hx.free(); hs.free()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    kernel_ms = ms_total / args.steps                    # one step == one launch of the dominant kernel
    alg_bytes = n * (4 * d + 8)
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    info = holder["forest"].info()
    line = {
        "metric": "rows scored/sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "strong" if tree_sharded else "weak", "vs_baseline": None,
        "dtype": "f32 features, f32 path sums, f64 scores" + (", f64 hyperplane dots" if ext >= 0 else ""),
        "data": "synthetic",
        "config": {"workload": workload_label(wl_name, n, d, T, ns, ext),
                   "parallelism": (f"trees sharded x{world}, partial sums scattered into NVLink peer memory by the scoring kernel"
                                   if fused else f"trees sharded x{world} + NCCL all-reduce of path sums" if tree_sharded else
                                   f"rows sharded x{world}, forest replicated, no data-path collective"),
                   "l2": f"inputs ({n * d * 4 / 1e9:.2f} GB/GPU) larger than L2; no flush needed",
                   "forest": {"nodes": int(info.num_nodes), "max_depth": int(info.max_depth),
                              "fit": "ifb_fit_device (this repo's GPU builder), seed 1", "fit_ms": round(fit_ms, 2)}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(wl_name), "peak_source": peak_src,
                     "kernel": "score_ext_*" if ext >= 0 else "score_std_kernel",
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
                     "note": "per GPU; one step = one launch of the dominant kernel"
                             + (" (+ the fit launch and its host-side table assembly)" if fit_in_step else "")},
    }
    if phases:
        line["phases"] = phases
        # the roofline object describes the scoring kernel alone
        line["roofline"].update(kernel_ms=phases["transform_ms"],
                                achieved=alg_bytes / (phases["transform_ms"] / 1e3) / 1e9,
                                frac=alg_bytes / (phases["transform_ms"] / 1e3) / 1e9 / peak)
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        O = graft.load_oracle()
        cores, cores_note = usable_cores()
        tables = forest.export()
        sample = np.ascontiguousarray(X[: min(n, 4_000_000 if d <= 64 else 200_000)].cpu().numpy())
        forest = holder["forest"]
        rate, rows, dt = cpu_port_rate(O, O.Forest(tables), sample, cores, target_s=12.0)
        line["cpu_baseline"] = {"value": rate, "unit": "rows/s", "cores": cores, "kind": "port",
                                "sample": f"first {rows} rows of the same matrix, same forest, {dt:.1f} s",
                                "cores_note": cores_note}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="config2")
    ap.add_argument("--shard", choices=["rows", "trees", "trees-fused"], default="rows")
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_native(args, args.workload, wl)


if __name__ == "__main__":
    main()
