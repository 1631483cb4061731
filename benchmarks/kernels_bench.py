#!/usr/bin/env python
"""Per-kernel roofline table for the SURVEY 8(d) configurations (device-resident data, CUDA-event timing on
the library's compute stream via its launch records).  Not the headline benchmark (that is bench.py); this
is the evidence that every kernel of the path, not only f64 add, runs near the HBM roofline.

    python benchmarks/kernels_bench.py [--rows 100000000] [--json gpurun_out/kernels.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import rust_dataframe_b200 as rdf  # noqa: E402
from rust_dataframe_b200 import native as N  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(ctx, fn, reps=12):
    """Run fn() reps times with profiling on; return per-kernel-kind (median ms, bytes, rows)."""
    for _ in range(3):
        out = fn()
        for o in out if isinstance(out, (list, tuple)) else [out]:
            if isinstance(o, rdf.Column):
                o.free()
    ctx.synchronize()
    ctx.profile_read()
    ctx.profile_enable(True)
    for _ in range(reps):
        out = fn()
        for o in out if isinstance(out, (list, tuple)) else [out]:
            if isinstance(o, rdf.Column):
                o.free()
    ctx.profile_enable(False)
    recs = ctx.profile_read()
    per_call = max(1, len(recs) // reps)
    by = {}
    for i, r in enumerate(recs):   # several launches of the same kind per call (filter: count+scan, scatter) stay separate
        key = r["kernel"] if per_call == 1 or sum(1 for q in recs[:per_call] if q["kernel"] == r["kernel"]) == 1 else f'{r["kernel"]}#{i % per_call}'
        by.setdefault(key, []).append(r)
    return {k: (float(np.median([r["ms"] for r in v])), v[0]["bytes"], v[0]["rows"]) for k, v in by.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ctx = rdf.default_context()
    chunk = 4_000_000
    lens = [chunk] * (args.rows // chunk)
    G = rdf.Column.generate
    pk = peak()
    rows_out = []

    def report(name, config, res, kind):
        ms, nbytes, rows = res[kind]
        gbs = nbytes / ms / 1e6
        rows_out.append({"case": name, "config": config, "kernel": kind, "ms": ms, "rows": rows, "bytes_per_row": nbytes / rows,
                         "GBs": gbs, "frac_of_measured_peak": gbs / pk, "frac_of_8TBs": gbs / 8000.0, "rows_per_s": rows / ms * 1e3})
        print(f"{name:44s} {ms:8.4f} ms  {nbytes / rows:7.3f} B/row  {gbs:8.1f} GB/s  {gbs / pk:5.3f} of measured  {rows / ms * 1e3:10.3e} rows/s", flush=True)

    # ---- config 1/metric: f64 add, sum ----
    a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0)
    b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
    c3 = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=2)
    d = G(rdf.F64, lens, 1, col_id=3)
    report("add f64 (no nulls)", "metric", timed(ctx, lambda: a.add(b)), "binary")
    report("add f64 + fused sum (K5)", "metric", timed(ctx, lambda: a.binary_agg(N.ADD, b)[0]), "binary")
    report("sum f64 (no nulls)", "metric", timed(ctx, lambda: a.sum()), "reduce")
    # ---- config 2: chain add, mul, div, sin ----
    report("multiply f64", "cfg2", timed(ctx, lambda: a.multiply(c3)), "binary")
    report("divide f64 (zero check fused)", "cfg2", timed(ctx, lambda: a.divide(d)), "binary")
    g = a.divide(d)
    report("sin f64, |x| <~ 1e3", "cfg2", timed(ctx, lambda: g.sin()), "unary")
    report("cos f64", "cfg2", timed(ctx, lambda: g.cos()), "unary")
    report("tan f64", "cfg2", timed(ctx, lambda: g.tan()), "unary")
    report("abs f64", "cfg2", timed(ctx, lambda: g.abs()), "unary")
    g.free()
    # config 2, second run: |g| up to 1e12 -- beyond the fast path of the trig argument reduction
    big = G(rdf.F64, lens, 0, -1e12, 1e12, col_id=14)
    report("sin f64, |x| up to 1e12 (slow-path reduction)", "cfg2-slow", timed(ctx, lambda: big.sin()), "unary")
    report("cos f64, |x| up to 1e12", "cfg2-slow", timed(ctx, lambda: big.cos()), "unary")
    report("tan f64, |x| up to 1e12", "cfg2-slow", timed(ctx, lambda: big.tan()), "unary")
    big.free()
    bn = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4, null_mod=10)
    dn = G(rdf.F64, lens, 1, col_id=5, null_mod=10)
    report("add f64, 10% nulls on one input", "cfg2-nulls", timed(ctx, lambda: a.add(bn)), "binary")
    report("add f64, 10% nulls on both inputs", "cfg2-nulls", timed(ctx, lambda: bn.add(dn)), "binary")
    report("divide f64, 10% nulls on divisor only", "cfg2-nulls", timed(ctx, lambda: a.divide(dn)), "binary")
    report("divide f64, 10% nulls on both", "cfg2-nulls", timed(ctx, lambda: bn.divide(dn)), "binary")
    report("sin f64, 10% nulls", "cfg2-nulls", timed(ctx, lambda: bn.sin()), "unary")
    # ---- N3: the whole config-2 chain in one pass (40 B/row) vs the four materialising launches (88 B/row) ----
    prog = [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)]
    report("N3 fused sin(((a+b)*c)/d)", "cfg2-fused", timed(ctx, lambda: rdf.eval_expr([a, b, c3, d], prog)), "expr")
    report("N3 fused ((a+b)*c)/d (arithmetic only)", "cfg2-fused", timed(ctx, lambda: rdf.eval_expr([a, b, c3, d], prog[:3])), "expr")
    report("N3 fused chain + sum/count of h (one pass)", "cfg2-fused", timed(ctx, lambda: rdf.eval_expr_agg([a, b, c3, d], prog)[0]), "expr")
    report("N3 sum(sin(((a+b)*c)/d)), h never written", "cfg2-fused", timed(ctx, lambda: rdf.eval_expr_agg([a, b, c3, d], prog, materialise=False)[1] and None), "expr")
    report("N3 fused chain, 10% nulls on b and d", "cfg2-fused", timed(ctx, lambda: rdf.eval_expr([a, bn, c3, dn], prog)), "expr")

    def unfused():
        e = a.add(b); f = e.multiply(c3); e.free(); g2 = f.divide(d); f.free(); h = g2.sin(); g2.free()
        return h
    res = timed(ctx, unfused)
    total = sum(v[0] for v in res.values())
    print(f"{'N3 reference: same chain, 4 launches':44s} {total:8.4f} ms  (sum of {sorted(res)})", flush=True)
    rows_out.append({"case": "cfg2 chain unfused (4 launches)", "config": "cfg2-fused", "kernel": "binary x3 + unary", "ms": total, "rows": args.rows})
    for col in (c3, d, bn, dn):
        col.free()
    # f32 trig
    f32 = G(rdf.F32, lens, 0, -1e3, 1e3, col_id=6)
    report("sin f32", "extra", timed(ctx, lambda: f32.sin()), "unary")
    f32b = G(rdf.F32, lens, 0, -1e3, 1e3, col_id=16)   # two DISTINCT inputs: a column added to itself moves 2w, not 3w, bytes per row
    report("add f32", "extra", timed(ctx, lambda: f32.add(f32b)), "binary")
    f32.free(); f32b.free()
    # ---- config 3: int64 aggregates with 10% nulls ----
    i64n = G(rdf.I64, lens, 3, col_id=7, null_mod=10)
    i64 = G(rdf.I64, lens, 3, col_id=8)
    report("sum/min/max/count i64, 10% nulls (4-in-1)", "cfg3", timed(ctx, lambda: i64n.aggregate_all()), "reduce")
    report("sum/min/max/count i64, no nulls", "cfg3", timed(ctx, lambda: i64.aggregate_all()), "reduce")
    report("add i64, 10% nulls", "cfg5", timed(ctx, lambda: i64n.add(i64)), "binary")
    report("add i64 + fused 4-in-1 aggregate", "cfg5", timed(ctx, lambda: i64n.binary_agg(N.ADD, i64)[0]), "binary")
    # ---- config 4: cast chain ----
    i32n = G(rdf.I32, lens, 2, col_id=9, null_mod=10)
    report("cast i32 -> f64, 10% nulls", "cfg4", timed(ctx, lambda: i32n.cast(rdf.F64)), "cast")
    report("cast i64 -> f64", "cfg5", timed(ctx, lambda: i64.cast(rdf.F64)), "cast")
    report("cast f64 -> i32 (fallible)", "cfg4", timed(ctx, lambda: a.cast(rdf.I32)), "cast")
    report("cast f64 -> f32", "extra", timed(ctx, lambda: a.cast(rdf.F32)), "cast")
    report("cast i64 -> i8 (fallible, narrow out)", "extra", timed(ctx, lambda: i64.cast(rdf.I8)), "cast")
    i8 = i64.cast(rdf.I8)
    i8b = i64n.cast(rdf.I8)
    report("add i8", "extra", timed(ctx, lambda: i8.add(i8b)), "binary")
    report("sum/min/max/count i8", "extra", timed(ctx, lambda: i8.aggregate_all()), "reduce")
    i8v = G(rdf.I8, lens, 2, col_id=17, null_mod=10)
    i16v = G(rdf.I16, lens, 2, col_id=18, null_mod=10)
    u16 = G(rdf.U16, lens, 2, col_id=19)
    report("sum/min/max/count i8 full range, 10% nulls", "extra", timed(ctx, lambda: i8v.aggregate_all()), "reduce")
    report("sum/min/max/count i16 full range, 10% nulls", "extra", timed(ctx, lambda: i16v.aggregate_all()), "reduce")
    report("sum/min/max/count u16, no nulls", "extra", timed(ctx, lambda: u16.aggregate_all()), "reduce")
    # the same kernels over as many BYTES as the f64 rows move (1e8 i8 rows are 0.11 GB: 17 us at the copy peak, launch-sized)
    big_lens = [32_000_000] * max(1, args.rows * 8 // 32_000_000)
    i8big = G(rdf.I8, big_lens, 2, col_id=27, null_mod=10)
    report("sum/min/max/count i8 10% nulls, 8x the rows", "extra", timed(ctx, lambda: i8big.aggregate_all()), "reduce")
    i8big.free()
    i16big = G(rdf.I16, big_lens[: len(big_lens) // 2], 2, col_id=28, null_mod=10)
    report("sum/min/max/count i16 10% nulls, 4x the rows", "extra", timed(ctx, lambda: i16big.aggregate_all()), "reduce")
    i16big.free()
    report("sum/min/max/count i32 full range, 10% nulls", "extra", timed(ctx, lambda: i32n.aggregate_all()), "reduce")
    for col in (i8b, i8v, i16v, u16):
        col.free()
    # ---- N2 (next row): BooleanFilter compare + ChunkedArray::filter ----
    report("compare f64 > scalar", "N2", timed(ctx, lambda: a.gt(0.0)), "compare")
    report("compare f64 > f64", "N2", timed(ctx, lambda: a.gt(b)), "compare")
    report("compare i32 (10% nulls) > scalar, cast fused", "N2", timed(ctx, lambda: i32n.gt(0.0)), "compare")
    report("compare i64 > f64 column, cast fused", "N2", timed(ctx, lambda: i64.gt(a)), "compare")
    m_half = a.gt(0.0)
    m_rare = a.gt(980.0)
    an = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=11, null_mod=10)
    for name, fn in (("filter f64, 50% kept", lambda: a.filter(m_half)), ("filter f64, 1% kept", lambda: a.filter(m_rare)),
                     ("filter f64 10% nulls, 50% kept", lambda: an.filter(m_half)), ("filter i32 10% nulls, 50% kept", lambda: i32n.filter(m_half))):
        res = timed(ctx, fn)
        keys = sorted(k for k in res if k.startswith("filter"))
        report(name + ": count+scan", "N2", res, keys[0])
        report(name + ": scatter", "N2", res, keys[1])
    # ---- DataFrame::sort: lexsort_to_indices (stable LSD radix, 32 B/row per executed pass) + take ----
    report("sort indices f64 (8 passes)", "sort", timed(ctx, lambda: rdf.sort_indices([(a, False)]), reps=4), "sort")
    report("sort indices i64 10% nulls, |x| < 2^40 (6+1 passes)", "sort", timed(ctx, lambda: rdf.sort_indices([(i64n, True)]), reps=4), "sort")
    report("sort indices i32 10% nulls (4+1 passes)", "sort", timed(ctx, lambda: rdf.sort_indices([(i32n, False)]), reps=4), "sort")
    report("sort indices i8 then f64 (two criteria)", "sort", timed(ctx, lambda: rdf.sort_indices([(i8, False), (a, True)]), reps=4), "sort")
    sidx = rdf.sort_indices([(a, False)])
    report("take f64 by the sort indices (random gather)", "sort", timed(ctx, lambda: b.take(sidx), reps=4), "take")
    report("take i32 10% nulls by the sort indices", "sort", timed(ctx, lambda: i32n.take(sidx), reps=4), "take")
    # ---- group-by aggregate: sort the key, gather key and value, head bitmap, two compactions, one warp per group ----
    def group_total(key, val, what):
        res = timed(ctx, lambda: _free_group(rdf.group_aggregate(key, [val])), reps=3)
        total = sum(v[0] for v in res.values())
        print(f"{what:44s} {total:8.4f} ms  (" + ", ".join(f"{k} {v[0]:.3f}" for k, v in sorted(res.items())) + ")", flush=True)
        rows_out.append({"case": what, "config": "group", "kernel": "sort + take + filter + group", "ms": total, "rows": args.rows, "parts": {k: v[0] for k, v in res.items()}})

    def _free_group(r):
        keys, res = r
        keys.free()
        for d in res:
            for c in d.values():
                if c is not None:
                    c.free()
        return None

    kf = G(rdf.F64, lens, 0, -100.0, 100.0, col_id=22)
    small = kf.cast(rdf.I8)                                        # 200 distinct keys: every group is a hot key (k_group_big)
    kf.free()
    group_total(small, a, "group-by i8 key (200 groups), sum f64")
    group_total(i32n, i64n, "group-by i32 key (~1e8 groups), i64 4-in-1")
    small.free()
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"rows": args.rows, "peak_GBs_measured": pk, "results": rows_out}, f, indent=1)


if __name__ == "__main__":
    main()
