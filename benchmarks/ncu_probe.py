"""One launch of each kernel the round-1 review asked ncu evidence for (fused expressions, filter scatter, radix scatter, take),
at the 1e8-row shapes of benchmarks/kernels_bench.py.  Run under `ncu --set full -k regex:...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf
from rust_dataframe_b200 import native as N

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0); b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
c3 = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=2); d = G(rdf.F64, lens, 1, col_id=3)
prog = [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)]
for _ in range(2):
    rdf.eval_expr([a, b, c3, d], prog).free()
    rdf.eval_expr([a, b, c3, d], prog[:3]).free()
    rdf.eval_expr_agg([a, b, c3, d], prog, materialise=False)
    m = a.gt(0.0)
    a.filter(m).free()
    idx = rdf.sort_indices([(a, False)])
    b.take(idx).free()
    idx.free(); m.free()
print("ok")
