"""Small ragged shapes through every kernel once (for compute-sanitizer memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_dataframe_b200 as rdf
from rust_dataframe_b200 import native as N

rng = np.random.default_rng(1)
lens = [0, 1, 33, 2047, 2049, 5000]
P = rdf.PrimitiveArray
for dtype in (rdf.I8, rdf.I16, rdf.I32, rdf.I64, rdf.U8, rdf.U16, rdf.U32, rdf.U64, rdf.F32, rdf.F64):
    npdt = rdf.NP_DTYPES[dtype]
    def col(nulls, nz=False):
        out = []
        for k, n in enumerate(lens):
            v = (rng.uniform(1, 100, n + 9) if npdt.kind == "f" else rng.integers(1, 100, n + 9)).astype(npdt)
            a = P.from_numpy(v, rng.random(n + 9) > 0.2) if nulls else P.from_numpy(v)
            a.null_count = -1 if nulls else 0
            out.append(a.slice(3 + k, n))
        return out
    a, b = col(True), col(False)
    for fn in (rdf.ScalarFunctions.add, rdf.ScalarFunctions.multiply, rdf.ScalarFunctions.divide):
        fn(a, b)
    rdf.AggregateFunctions.all([c for c in a if c.length])
    rdf.AggregateFunctions.count(a)
    for to in (rdf.I8, rdf.U32, rdf.F64, rdf.I64):
        rdf.cast(a, to)
    if dtype in (rdf.F32, rdf.F64):
        rdf.ScalarFunctions.sin(a); rdf.ScalarFunctions.tan(b); rdf.ScalarFunctions.atan2(a[3], b[3])
    ca, cb = rdf.Column.upload_many([a, b], asynchronous=True)
    cc, fut = ca.binary_agg_async(N.ADD, cb)
    fut.result(); cc.download(); ca.download()
    g = rdf.Column.generate(dtype, lens, kind=2 if npdt.kind != "f" else 0, null_mod=3)
    g.download(); g.aggregate_all()
    if dtype == rdf.F64:   # N3: fused expression with both temporaries, a shared node, a divide and a libm node
        prog = [(N.ADD, 0, 1), (N.SUB, 0, 1), (N.MUL, 2, 3), (N.MUL, 2, 2), (N.ADD, 4, 5), (N.DIV, 6, 1), ("sin", 7)]
        rdf.eval_expr([ca, cb], prog).download()
    if dtype in (rdf.F64, rdf.I16):   # DataFrame::sort: two criteria (nullable + dense), take of a numeric and a boolean column
        idx = rdf.sort_indices([(ca, True), (cb, False)])
        ca.take(idx).download(); cb.gt(50.0).take(idx).download()
    if dtype not in (rdf.I64, rdf.U64):
        rdf.AggregateFunctions.avg([c for c in a if c.length])
    # round 2: comparisons against a scalar (integer columns take k_compare_int) and a column, filter, batched multi-column reduce
    m = ca.gt(50.0); m2 = ca.le(cb)
    ca.filter(m).download(); cb.filter(m2).download()
    rdf.Column.aggregate_all_many([ca, cb, g])
    # group-by: short groups (one lane per group), medium (8 lanes), long (a warp) and one hot key (> 64 Ki rows: k_group_big)
    for n_rows, card in ((3000, 2500), (6000, 300), (9000, 7)):
        k = rdf.Column.upload([P.from_numpy(rng.integers(0, card, n_rows).astype(np.int32), rng.random(n_rows) > 0.1)])
        v = rdf.Column.upload([P.from_numpy((rng.uniform(1, 100, n_rows) if npdt.kind == "f" else rng.integers(1, 100, n_rows)).astype(npdt), rng.random(n_rows) > 0.2)])
        keys, res = rdf.group_aggregate(k, [v])
        keys.download(); res[0]["sum"].download(); res[0]["count"].download()
        if res[0]["min"] is not None:
            res[0]["min"].download(); res[0]["max"].download()
    if dtype in (rdf.I64, rdf.F64, rdf.U8):
        n_rows = 150_000
        kk = rng.integers(0, 50_000, n_rows).astype(np.int64); kk[rng.random(n_rows) < 0.6] = 7
        k = rdf.Column.upload([P.from_numpy(kk)])
        v = rdf.Column.upload([P.from_numpy((rng.uniform(1, 100, n_rows) if npdt.kind == "f" else rng.integers(1, 100, n_rows)).astype(npdt), rng.random(n_rows) > 0.2)])
        keys, res = rdf.group_aggregate(k, [v])
        keys.download(); res[0]["sum"].download()
# round 2: the fleet code path on one GPU, and a single-rank communicator (pack -> combine -> unpack on the stream)
fleet = rdf.Context.multi(1)
fa = [P.from_numpy(rng.integers(-100, 100, n).astype(np.int64), rng.random(n) > 0.3) for n in (1000, 0, 77)]
rdf.ScalarFunctions.add(fa, fa, ctx=fleet); rdf.AggregateFunctions.all([c for c in fa if c.length], ctx=fleet)
fleet.close()
solo = rdf.Context(0)
solo.comm_attach(rdf.Context.comm_unique_id(), 0, 1)
rdf.AggregateFunctions.all([c for c in fa if c.length], ctx=solo)
solo.close()
print("sanitize cases done")
