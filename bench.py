#!/usr/bin/env python
"""bench.py -- rows/s on the 1e8-row f64 add+sum hot path (BASELINE.json `metric`), one process per GPU.

    python bench.py --gpus 1 --steps 50 --warmup 5              # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus 1 --steps 5 ...     # the reference's CPU path (C restatement, host cores)
    torchrun --nproc-per-node N bench.py --gpus N ...           # N GPUs, weak scaling (1e8 rows per GPU)

A "step" is one pass of the hot path over one batch of synthetic input:
    c = ScalarFunctions::add(a, b)   (25 chunks x 4e6 rows of Float64, no nulls)   -> 24 B/row through HBM
    s = AggregateFunctions::sum(c)                                                  ->  8 B/row through HBM
`value` = rows/s with a and b already resident in HBM (c is materialised in HBM every step, s comes back to
the host every step; at N > 1 the per-GPU partial sums are combined with one NCCL all-reduce per step).
`e2e`   = the same step through the public API with HOST buffers: a and b start in pinned host memory and
are copied to the device inside the timed region, c is copied back to pinned host memory, s to the host.
Prints exactly one JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = 100_000_000
CHUNK = 4_000_000
SEED = 20260924
METRIC = "rows/s on 1e8-row f64 add+sum"
WORKLOAD = "1e8 rows x 2 Float64 cols (25 chunks x 4e6 rows, no nulls): c = a + b, then sum(c)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples = []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self, t0: float, t1: float):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nv
        win = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        seen = set()
        for _, _, r in win:
            for bit, name in names.items():
                if r & bit:
                    seen.add(name)
        return {"sm_mhz": float(np.median([s[1] for s in win])) if win else None, "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(seen), "samples": len(win)}


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process (and therefore its pinned host buffers, by first touch) to the NUMA node the GPU hangs off.
    Pure placement: host<->device copies then do not cross the inter-socket link.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path (C restatement in oracle/, see oracle/oracle.h for provenance)

def cpu_step_factory(rows: int):
    from oracle import pyoracle as orc

    n_chunks = max(1, rows // CHUNK)
    lens = [CHUNK] * n_chunks
    a = [orc.generate(orc.F64, 0, -1e3, 1e3, SEED, 0, i * CHUNK, n) for i, n in enumerate(lens)]
    b = [orc.generate(orc.F64, 0, -1e3, 1e3, SEED, 1, i * CHUNK, n) for i, n in enumerate(lens)]
    threads = max(1, min(os.cpu_count() or 1, n_chunks))
    views_a, views_b = orc._views(a), orc._views(b)
    outs, bufs = orc._alloc_outs(orc.F64, lens)  # output buffers allocated once (arrow allocates per call)
    L = orc.lib()
    import ctypes as C

    res = np.zeros(1)
    some = C.c_int32(0)
    out_views = (orc.View * n_chunks)()
    for i, (v, _) in enumerate(bufs):
        out_views[i].values = v.ctypes.data
        out_views[i].validity = None
        out_views[i].len = lens[i]
        out_views[i].offset = 0
        out_views[i].null_count = 0

    def step():
        # ScalarFunctions::add: rayon par_iter over chunks (scalar.rs:28-31) -> one thread per chunk up to nproc
        st = L.orc_col_binary(orc.ADD, orc.F64, n_chunks, views_a, n_chunks, views_b, outs, threads)
        assert st == 0
        # AggregateFunctions::sum: chunks in order on the calling thread, sequential fold (aggregate.rs:82-93)
        st = L.orc_aggregate(orc.SUM, orc.F64, n_chunks, out_views, res.ctypes.data, C.byref(some))
        assert st == 0
        return float(res[0])

    step._keep = (a, b, bufs, views_a, views_b, outs, out_views)  # the ctypes views hold raw pointers into these
    return step, n_chunks * CHUNK, threads


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    total_steps = args.steps + args.warmup
    rows = int(min(ROWS, max(CHUNK, (60 * 2.5e8 / max(total_steps, 1)) // CHUNK * CHUNK)))
    step, rows, threads = cpu_step_factory(rows)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = f"{rows} rows ({rows // CHUNK} chunks x {CHUNK}) per step, {args.steps} steps; add on {threads} threads (one per chunk, rayon mirror), sum on 1 thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": rows, "note": "CPU path of the reference: C restatement of rust-dataframe @ a8310afd + arrow-rs~2.0 semantics (oracle/oracle.c); the Rust reference cannot be built in this image"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm

def run_gpu_arm(args, rank: int, world: int, local: int):
    import rust_dataframe_b200 as rdf
    from rust_dataframe_b200 import native as N

    ctx = rdf.Context(local)
    lens = [CHUNK] * (ROWS // CHUNK)
    row0 = rank * ROWS  # every rank owns its own 1e8-row shard of the Vec<RecordBatch> (weak scaling)
    a = rdf.Column.generate(rdf.F64, lens, 0, -1e3, 1e3, seed=SEED, col_id=0, row0=row0, ctx=ctx)
    b = rdf.Column.generate(rdf.F64, lens, 0, -1e3, 1e3, seed=SEED, col_id=1, row0=row0, ctx=ctx)
    ctx.synchronize()

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

    def combine(s: float, count: int):
        """The one exchange step of the path: NCCL all-reduce of the per-GPU partial aggregates (blocking form)."""
        if world == 1:
            return s, count
        return combine_finish(combine_start(s, count))

    ring = [torch.zeros(2, dtype=torch.float64, device=f"cuda:{local}") for _ in range(8)] if world > 1 else []
    ring_pos = [0]

    def combine_start(s: float, count: int):
        """Enqueue the all-reduce of this step's partials; the result is read one step later."""
        t = ring[ring_pos[0] % len(ring)]
        ring_pos[0] += 1
        t.copy_(torch.tensor([s, float(count)], dtype=torch.float64))
        return t, dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)

    def combine_finish(handle):
        t, work = handle
        work.wait()
        r = t.cpu()
        return float(r[0]), int(r[1])

    def step_two_call():
        """The reference's call sequence, blocking: c = add(a, b); s = sum(c) (two passes over c)."""
        c = a.add(b)
        s = c.sum()
        c.free()
        return combine(float(s), ROWS)

    inflight = []
    DEPTH = 1 if world == 1 else int(os.environ.get("BDF_BENCH_DEPTH", "3"))  # steps kept in flight by the host (the NCCL combine of step i runs under the next ones)

    pending_combine = []

    def retire():
        c, fut = inflight.pop(0)
        r = fut.result()
        c.free()
        if world == 1:
            return float(r["sum"]), int(r["count"])
        pending_combine.append(combine_start(float(r["sum"]), int(r["count"])))  # overlaps the next step
        return combine_finish(pending_combine.pop(0)) if len(pending_combine) > 1 else None

    def drain():
        last = None
        while inflight:
            last = retire() or last
        while pending_combine:
            last = combine_finish(pending_combine.pop(0))
        return last

    def step():
        """c = a + b materialised in HBM with sum(c) folded into the same pass (K5); the host keeps one step in
        flight: step i's scalar is fetched after step i+1 has been enqueued."""
        inflight.append(a.binary_agg_async(N.ADD, b))
        return retire() if len(inflight) > DEPTH else None

    def barrier():
        ctx.synchronize()
        if world > 1:
            import torch

            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    drain()
    # secondary figure: the unfused, blocking two-call sequence (what a caller of add() then sum() gets)
    for _ in range(3):
        step_two_call()
    barrier()
    two_steps = max(3, min(args.steps, 20))
    ctx.timer_start()
    for _ in range(two_steps):
        step_two_call()
    two_call_ms = ctx.timer_stop() / two_steps
    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile_read()  # drop warm-up records
    ctx.profile_enable(True)
    barrier()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    ctx.timer_start()
    last = None
    for _ in range(args.steps):
        last = step() or last
    last = drain() or last
    ms = ctx.timer_stop()
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count() - launches0
    ctx.profile_enable(False)
    records = ctx.profile_read()
    if sampler.ok and len([s for s in sampler.samples if t_wall0 <= s[0] <= t_wall1]) < 3:
        t_extra = time.perf_counter()
        while time.perf_counter() - t_extra < 0.15:  # same load, untimed, only to observe the clocks
            step()
        drain()
        t_wall1_clk = time.perf_counter()
    else:
        t_wall1_clk = t_wall1
    sampler.stop_flag = True
    sampler.join(timeout=1)
    clocks = sampler.summary(t_wall0, t_wall1_clk)

    if world > 1:
        import torch

        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- e2e: host buffers in pinned memory, copies inside the timed region ----
    e2e = None if args.skip_e2e else run_e2e(args, ctx, rdf, lens, a, b, world, local, combine, dist)

    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    bins = [r for r in records if r["kernel"] == "binary"]
    reds = [r for r in records if r["kernel"] == "reduce"]
    add_ms = float(np.mean([r["ms"] for r in bins])) if bins else None
    red_ms = float(np.mean([r["ms"] for r in reds])) if reds else None
    add_bytes = bins[0]["bytes"] if bins else 24 * ROWS
    achieved = add_bytes / (add_ms * 1e-3) / 1e9 if add_ms else None
    roofline = {
        "bound": "hbm", "kernel": "k_binary<double,ADD> (one launch over 25 chunks)", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": add_bytes, "avg_launch_ms": add_ms, "frac_of_8TBs_nominal": achieved / 8000.0 if achieved else None,
        "sum_kernel": {"kernel": "k_reduce<double>", "avg_launch_ms": red_ms, "algorithmic_bytes_per_launch": reds[0]["bytes"] if reds else None,
                       "achieved": (reds[0]["bytes"] / (red_ms * 1e-3) / 1e9) if red_ms else None},
        "step_GBs_algorithmic_24B_per_row": 24 * ROWS / (ms / args.steps * 1e-3) / 1e9,
        "two_call_unfused": {"ms_per_step": two_call_ms, "rows_per_s": ROWS * world / (two_call_ms * 1e-3),
                             "note": "blocking add() then sum(): 32 B/row, two kernels + a host sync per step"},
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            roofline["traffic"] = json.load(f).get("k_binary_f64_add_bytes_per_launch")

    cpu = None
    if (world == 1 or rank == 0) and not args.skip_cpu:
        cpu = run_cpu_baseline()

    value = ROWS * world * args.steps / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": ROWS, "chunks": len(lens), "parallelism": f"shard{world}",
                   "l2": "inputs (2.4 GB working set per step) are larger than the 126 MB L2; no flush needed",
                   "fused": "sum(c) is computed by the add kernel while c is written (c is still materialised): 24 B/row",
                   "host_pipelining": f"{DEPTH} step(s) in flight: step i's scalar is read (and combined across ranks) after step i+{DEPTH} is enqueued",
                   "collective": "none" if world == 1 else "1 NCCL all-reduce of the partial (sum,count) per step, enqueued asynchronously and read one step later"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "check": {"sum": last[0], "count": last[1]},
    }
    print(json.dumps(line), flush=True)


def run_e2e(args, ctx, rdf, lens, dev_a, dev_b, world, local, combine, dist):
    """Same step through the public API with host buffers (pinned), H2D + D2H inside the timed region."""
    from rust_dataframe_b200 import native as N

    steps = max(3, min(args.steps, 8))
    # host copies of the same synthetic columns, in pinned memory (what the Rust shim would hand over as Arrow buffers)
    bufs_a = N.alloc_outputs(rdf.F64, lens, ctx, pinned=True)
    bufs_b = N.alloc_outputs(rdf.F64, lens, ctx, pinned=True)
    host_a = dev_a.download(into=bufs_a)
    host_b = dev_b.download(into=bufs_b)
    out_bufs = N.alloc_outputs(rdf.F64, lens, ctx, pinned=True)

    def step():
        ca, cb = rdf.Column.upload_many([host_a, host_b], ctx=ctx, asynchronous=True)  # a0,b0,a1,b1,... over PCIe
        cc, fut = ca.binary_agg_async(N.ADD, cb)      # per upload group, as soon as its chunks have landed
        cc.download_begin(out_bufs)                   # D2H of group g overlaps H2D of group g+1
        r = fut.result()
        cc.download_end(out_bufs)
        for col in (ca, cb, cc):
            col.free()
        return combine(float(r["sum"]), int(r["count"]))

    def barrier():
        ctx.synchronize()
        if world > 1:
            import torch

            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()

    for _ in range(2):
        step()
    barrier()
    ctx.timer_start()
    last = None
    for _ in range(steps):
        last = step()
    ms = ctx.timer_stop()
    barrier()
    if world > 1:
        import torch

        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h2d = 2 * 8 * ROWS
    d2h = 8 * ROWS + 8
    return {"value": ROWS * world * steps / (ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "check_sum": last[0] if last else None,
            "ms_per_step": ms / steps, "steps": steps,
            "api": "Column.upload_many([a,b]) [pinned host, async] -> binary_agg_async(ADD) -> download_begin/end(c) [pinned host] + scalar",
            "pcie_GBs": (h2d + d2h) / (ms / steps * 1e-3) / 1e9}


def run_cpu_baseline():
    """Oracle ('port' of the reference's CPU path) timed on this box's host cores, bounded sample."""
    try:
        step, rows, threads = cpu_step_factory(ROWS)
        step()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 8 and reps < 40):
            step()
            reps += 1
        dt = time.perf_counter() - t0
        res_pa = None
        try:  # secondary reference point, NOT the reference: Arrow C++ via pyarrow on the same workload
            import pyarrow as pa
            import pyarrow.compute as pc

            a_np, b_np, _, _, _, _, _ = step._keep
            ca = pa.chunked_array([pa.array(x.values) for x in a_np])
            cb = pa.chunked_array([pa.array(x.values) for x in b_np])
            pc.sum(pc.add(ca, cb))
            t1 = time.perf_counter()
            pc.sum(pc.add(ca, cb))
            res_pa = {"rows_per_s": rows / (time.perf_counter() - t1), "note": "pyarrow %s pc.add + pc.sum (Arrow C++, differs from arrow-rs on f64 div-by-zero, casts, sum order)" % pa.__version__}
        except Exception as e:  # pragma: no cover
            res_pa = {"rows_per_s": None, "note": f"unavailable: {e}"}
        return {"value": rows * reps / dt, "unit": "rows/s", "cores": threads, "kind": "port", "arrow_cpp_pyarrow": res_pa,
                "sample": f"{reps} passes over the full {rows}-row workload; add on {threads} threads (one per chunk, like rayon), sum sequential on 1 thread; host has {os.cpu_count()} cpus",
                "ms_per_step": dt / reps * 1e3}
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": "rows/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup(args.gpus)
    bind_to_gpu_numa_node(local)
    try:
        run_gpu_arm(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
