#!/usr/bin/env python
"""One launch each of a few fused expressions, for `ncu -k regex:k_expr` captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf  # noqa: E402
from rust_dataframe_b200 import native as N  # noqa: E402

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0); b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
c = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=2); d = G(rdf.F64, lens, 1, col_id=3)
rdf.eval_expr([a], [("sin", 0)])                                                    # launch 0
rdf.eval_expr([a, b, c, d], [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3)])          # launch 1
rdf.eval_expr([a, b, c, d], [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)])  # launch 2
ctx.synchronize()
