"""Runs a few single launches (for ncu captures of the weaker kernels).  usage: probe_cases.py case [case...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf
from rust_dataframe_b200 import native as N

lens = [4_000_000] * 25
G = rdf.Column.generate
ctx = rdf.default_context()
cases = sys.argv[1:] or ["div_nulls", "sin_nulls", "agg_i64_nulls", "sum_f64"]
for case in cases:
    for rep in range(2):
        if case == "div_nulls":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4, null_mod=10); b = G(rdf.F64, lens, 1, col_id=5, null_mod=10)
            a.divide(b).free()
        elif case == "add_nulls2":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4, null_mod=10); b = G(rdf.F64, lens, 1, col_id=5, null_mod=10)
            a.add(b).free()
        elif case == "div":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4); b = G(rdf.F64, lens, 1, col_id=5)
            a.divide(b).free()
        elif case == "sin_nulls":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4, null_mod=10)
            a.sin().free()
        elif case == "sin":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=4)
            a.sin().free()
        elif case == "agg_i64_nulls":
            a = G(rdf.I64, lens, 3, col_id=7, null_mod=10)
            a.aggregate_all()
        elif case == "sum_f64":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0)
            a.sum()
        elif case == "cast_i32_f64":
            a = G(rdf.I32, lens, 2, col_id=9, null_mod=10)
            a.cast(rdf.F64).free()
        elif case == "add_i64_agg":
            a = G(rdf.I64, lens, 3, col_id=7, null_mod=10); b = G(rdf.I64, lens, 3, col_id=8)
            a.binary_agg(N.ADD, b)[0].free()
        elif case == "filter_rare":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0)
            m = a.gt(980.0)
            a.filter(m).free()
        elif case == "filter_half":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0)
            m = a.gt(0.0)
            a.filter(m).free()
        elif case == "add_sum_fused":
            a = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0); b = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=1)
            a.binary_agg(N.ADD, b)[0].free()
        ctx.synchronize()
print("done")
