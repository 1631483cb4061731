"""take (random gather) and a few other kernels under the current L2 fetch-granularity hint (env BDF_L2_FETCH=32|64|128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0); b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
i32 = G(rdf.I32, lens, 2, col_id=9, null_mod=10)
idx = rdf.sort_indices([(a, False)])
m = a.gt(0.0)


def timed(fn, reps=5):
    r = fn()
    if hasattr(r, "free"):
        r.free()
    ctx.synchronize(); ctx.profile_read(); ctx.profile_enable(True)
    for _ in range(reps):
        r = fn()
        if hasattr(r, "free"):
            r.free()
    ctx.profile_enable(False)
    recs = ctx.profile_read()
    return {k: round(sum(r["ms"] for r in recs if r["kernel"] == k) / reps, 4) for k in sorted({r["kernel"] for r in recs})}


print("BDF_L2_FETCH =", os.environ.get("BDF_L2_FETCH", "default"))
print("take f64:", timed(lambda: b.take(idx)))
print("take i32 nulls:", timed(lambda: i32.take(idx)))
print("add f64:", timed(lambda: a.add(b)))
print("sum f64:", timed(lambda: a.aggregate_all_async()))
print("filter 50%:", timed(lambda: a.filter(m)))
print("sort f64:", timed(lambda: rdf.sort_indices([(a, False)]), reps=2))
