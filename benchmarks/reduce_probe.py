"""A few k_reduce launches for ncu (1e8 rows per column): i64 with and without nulls, f64, i32, i16, i8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf

ctx = rdf.default_context()
lens = [4_000_000] * 25
G = rdf.Column.generate
cols = [G(rdf.I64, lens, 3, col_id=7, null_mod=10), G(rdf.I64, lens, 3, col_id=8), G(rdf.F64, lens, 0, -1e3, 1e3, col_id=0),
        G(rdf.I32, lens, 2, col_id=9, null_mod=10), G(rdf.I16, lens, 2, col_id=18, null_mod=10), G(rdf.I8, lens, 2, col_id=17, null_mod=10)]
for rep in range(3):
    for c in cols:
        c.aggregate_all()
print("ok")
