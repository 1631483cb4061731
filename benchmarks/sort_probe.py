"""DataFrame::sort rows only, (a short version of the sort rows of kernels_bench.py, for A/B runs of the radix pass): 1e8 rows, CUDA events on the library's stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0)
i64n = G(rdf.I64, lens, 3, col_id=7, null_mod=10)
i32n = G(rdf.I32, lens, 2, col_id=9, null_mod=10)


def timed(fn, reps=4):
    fn().free()
    ctx.synchronize(); ctx.profile_read(); ctx.profile_enable(True)
    for _ in range(reps):
        fn().free()
    ctx.profile_enable(False)
    recs = ctx.profile_read()
    return sum(r["ms"] for r in recs if r["kernel"] == "sort") / reps


print("sort f64            %.3f ms" % timed(lambda: rdf.sort_indices([(a, False)])))
print("sort i64 10%% nulls  %.3f ms" % timed(lambda: rdf.sort_indices([(i64n, True)])))
print("sort i32 10%% nulls  %.3f ms" % timed(lambda: rdf.sort_indices([(i32n, False)])))
