#!/usr/bin/env python
"""N3 timing: the config-2 chain fused (bdf_eval_expr_dev) vs its four materialising launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rust_dataframe_b200 as rdf  # noqa: E402
from rust_dataframe_b200 import native as N  # noqa: E402
from kernels_bench import timed  # noqa: E402

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0); b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
c = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=2); d = G(rdf.F64, lens, 1, col_id=3)
bn = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4, null_mod=10); dn = G(rdf.F64, lens, 1, col_id=5, null_mod=10)
i32n = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=9, null_mod=10).cast(rdf.I32)   # |x| < 1e3 like column a, 10 % nulls
prog = [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)]
def unfused_typed():
    f = i32n.cast(rdf.F64); e = f.add(b); f.free(); g = e.multiply(c); e.free(); h = g.divide(d); g.free(); r = h.sin(); h.free()
    return r


for name, fn in (("add only", lambda: rdf.eval_expr([a, b], [(N.ADD, 0, 1)])),
                 ("(a+b)*c", lambda: rdf.eval_expr([a, b, c], [(N.ADD, 0, 1), (N.MUL, 3, 2)])),
                 ("((a+b)*c)/d", lambda: rdf.eval_expr([a, b, c, d], prog[:3])),
                 ("sin(((a+b)*c)/d)", lambda: rdf.eval_expr([a, b, c, d], prog)),
                 ("sin chain, 10% nulls b,d", lambda: rdf.eval_expr([a, bn, c, dn], prog)),
                 ("sin(a)", lambda: rdf.eval_expr([a], [("sin", 0)])),
                 ("sin(((i32+b)*c)/d), i32 10% nulls cast on load", lambda: rdf.eval_expr([i32n, b, c, d], prog)),
                 ("same, unfused: cast, add, mul, div, sin", lambda: unfused_typed())):
    res = timed(ctx, fn)
    if "expr" not in res:   # the unfused reference: sum of its launches
        print(f"{name:50s} {sum(v[0] for v in res.values()):8.4f} ms  ({len(res)} launches)", flush=True)
        continue
    ms, nbytes, rows = res["expr"]
    print(f"{name:50s} {ms:8.4f} ms  {nbytes / rows:6.2f} B/row  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)
