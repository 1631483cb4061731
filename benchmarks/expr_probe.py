"""The fused-expression rows of the kernel table under the current BDF_EXPR_BULK setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_dataframe_b200 as rdf
from rust_dataframe_b200 import native as N

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0); b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
c3 = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=2); d = G(rdf.F64, lens, 1, col_id=3)
bn = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4, null_mod=10); dn = G(rdf.F64, lens, 1, col_id=5, null_mod=10)
prog = [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)]


def timed(fn, reps=8):
    r = fn()
    if hasattr(r, "free"):
        r.free()
    ctx.synchronize(); ctx.profile_read(); ctx.profile_enable(True)
    for _ in range(reps):
        r = fn()
        if hasattr(r, "free"):
            r.free()
    ctx.profile_enable(False)
    recs = [x for x in ctx.profile_read() if x["kernel"] == "expr"]
    return round(float(np.median([x["ms"] for x in recs])), 4)


print("BDF_EXPR_BULK =", os.environ.get("BDF_EXPR_BULK", "default(1)"))
print("sin chain         :", timed(lambda: rdf.eval_expr([a, b, c3, d], prog)))
print("arithmetic chain  :", timed(lambda: rdf.eval_expr([a, b, c3, d], prog[:3])))
print("chain + sum       :", timed(lambda: rdf.eval_expr_agg([a, b, c3, d], prog)[0]))
print("sum only (no out) :", timed(lambda: rdf.eval_expr_agg([a, b, c3, d], prog, materialise=False)[1] and None))
print("chain, 10% nulls  :", timed(lambda: rdf.eval_expr([a, bn, c3, dn], prog)))
print("a+b               :", timed(lambda: rdf.eval_expr([a, b], [(N.ADD, 0, 1)])))
