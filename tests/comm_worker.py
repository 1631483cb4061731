"""One rank of the multi-GPU parity run (launched by tests/test_comm_gpu.py through torch.distributed.run, one process
per GPU).  Every rank builds the SAME full columns from a fixed seed, keeps its shard (chunk i -> rank i mod N), runs
the collective entries of libb200df (the grouped ncclAllReduce is enqueued inside the library) and checks the result
against the CPU oracle evaluated on the FULL column.  Rank 0 prints "COMM-OK <json>" when every rank passed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20260924
CHUNK = 4_000_000


def chunks_for(rdf, name, seed, lens, null_frac):
    npdt = np.dtype(name)
    rng = np.random.default_rng(seed)
    out = []
    for n in lens:
        if npdt.kind == "f":
            v = rng.uniform(-1e3, 1e3, n).astype(npdt)
        else:
            info = np.iinfo(npdt)
            v = rng.integers(info.min, info.max, n, dtype=npdt, endpoint=True)
        out.append(rdf.PrimitiveArray.from_numpy(v, rng.random(n) >= null_frac if null_frac else None))
    return out


def float_ok(got, exact, sum_abs, n, eps):
    return abs(np.longdouble(got) - exact) <= 16 * np.log2(max(n, 2)) * eps * sum_abs + 1e-300


def main():
    import rust_dataframe_b200 as rdf
    from oracle import pyoracle as orc
    from rust_dataframe_b200 import native as N
    from rust_dataframe_b200 import parallel

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    full_size = "--full" in sys.argv
    ctx = rdf.Context(local)
    parallel.attach_communicator(ctx)
    info = ctx.comm_info()
    assert info["world"] == world and info["rank"] == rank and info["nccl_version"] > 0
    checks = 0

    # ---- 1. host-in entries on ragged shards, all ten types -------------------------------------------------------
    lens = [1000, 1, 0, 5000, 33, 2048, 7, 70001, 4096]
    for name, null_frac in [("int64", 0.1), ("uint64", 0.1), ("int8", 0.0), ("uint16", 0.3), ("int32", 0.5), ("uint8", 0.2),
                            ("int16", 0.0), ("uint32", 0.1), ("float64", 0.1), ("float32", 0.0)]:
        full = chunks_for(rdf, name, 99, lens, null_frac)
        dtype = full[0].dtype
        mine = parallel.shard(full, rank, world)
        got = rdf.AggregateFunctions.all(mine, dtype=dtype, ctx=ctx)           # collective: the aggregate of the FULL column
        want_count = int(orc.aggregate(orc.COUNT, dtype, full)[1])
        assert got["count"] == want_count and got["rows"] == sum(lens) and got["n_chunks"] == len(lens), (name, got)
        assert got["would_panic"], name   # the empty chunk makes the reference's max/min panic, whichever rank holds it
        if name.startswith("float"):
            exact, sum_abs = orc.sum_exact(dtype, full)
            assert float_ok(got["sum"], exact, sum_abs, sum(lens), 2.0 ** -53 if name == "float64" else 2.0 ** -24), (name, got["sum"], exact)
            s = rdf.AggregateFunctions.sum(mine, dtype=dtype, ctx=ctx)
            assert float_ok(s, exact, sum_abs, sum(lens), 2.0 ** -53 if name == "float64" else 2.0 ** -24)
        else:
            nonempty = [c for c in full if c.valid_mask().any()]
            for op, key in ((orc.SUM, "sum"), (orc.MIN, "min"), (orc.MAX, "max")):
                want = orc.aggregate(op, dtype, full if op == orc.SUM else nonempty)[1]
                assert int(got[key]) == int(want), (name, key, got[key], want)
            assert int(rdf.AggregateFunctions.sum(mine, dtype=dtype, ctx=ctx)) == int(orc.aggregate(orc.SUM, dtype, full)[1])
            try:
                rdf.AggregateFunctions.max(mine, dtype=dtype, ctx=ctx)
                raise AssertionError("max over a column with an empty chunk must report the reference's panic on every rank")
            except rdf.ReferencePanic:
                pass
        assert rdf.AggregateFunctions.count(mine, dtype=dtype, ctx=ctx) == want_count
        if name not in ("int64", "uint64"):
            a_got = rdf.AggregateFunctions.avg(mine, dtype=dtype, ctx=ctx)
            vals = np.concatenate([c.value_slice()[c.valid_mask()].astype(np.float64) for c in full])
            assert a_got is not None and abs(a_got - vals.mean()) <= 1e-9 * max(1.0, abs(vals.mean())) + 1e-6 * np.abs(vals).max() / max(1, len(vals)) ** 0.5, (name, a_got, vals.mean())
        checks += 1

    # ---- 2. fewer chunks than ranks: some ranks hold nothing, the calls stay collective ---------------------------
    one = chunks_for(rdf, "int64", 5, [12345], 0.2)
    mine = parallel.shard(one, rank, world)
    got = rdf.AggregateFunctions.all(mine, dtype=rdf.I64, ctx=ctx)
    assert got["n_chunks"] == 1 and not got["would_panic"]
    for op, key in ((orc.SUM, "sum"), (orc.MIN, "min"), (orc.MAX, "max"), (orc.COUNT, "count")):
        assert int(got[key]) == int(orc.aggregate(op, rdf.I64, one)[1]), key
    assert int(rdf.AggregateFunctions.max(mine, dtype=rdf.I64, ctx=ctx)) == int(orc.aggregate(orc.MAX, rdf.I64, one)[1])
    none = rdf.AggregateFunctions.all([], dtype=rdf.I64, ctx=ctx)   # Vec::new() on every rank
    assert none["n_chunks"] == 0 and none["count"] == 0 and none["min"] is None
    checks += 1

    # ---- 3. device columns: fused add + aggregate, multi-column call, futures in flight ----------------------------
    lens3 = [300_000] * 5 + [17]
    fa = chunks_for(rdf, "float64", 1, lens3, 0.0)
    fb = chunks_for(rdf, "float64", 2, lens3, 0.1)
    ia = chunks_for(rdf, "int64", 3, lens3, 0.1)
    ib = chunks_for(rdf, "int64", 4, lens3, 0.0)
    my = lambda col: parallel.shard(col, rank, world)
    ca, cb, cia, cib = (rdf.Column.upload(my(x), ctx=ctx, dtype=x[0].dtype) for x in (fa, fb, ia, ib))
    futs = []
    for _ in range(5):   # several collectives in flight on the finish stream
        futs.append(ca.binary_agg_async(N.ADD, cb))
    st, fc = orc.col_binary(orc.ADD, orc.F64, fa, fb)
    exact, sum_abs = orc.sum_exact(orc.F64, fc)
    for col, fut in futs:
        r = fut.result()
        assert r["rows"] == sum(lens3) and r["count"] == sum(c.length - c.null_count for c in fc)
        assert float_ok(r["sum"], exact, sum_abs, sum(lens3), 2.0 ** -53)
        col.free()
    first = None
    for _ in range(3):   # deterministic for a given world size
        col, r = ca.binary_agg(N.ADD, cb)
        first = first if first is not None else r["sum"]
        assert np.float64(r["sum"]).view(np.uint64) == np.float64(first).view(np.uint64)
        col.free()
    ci, ri = cia.binary_agg(N.MUL, cib)
    st, ic = orc.col_binary(orc.MUL, orc.I64, ia, ib)
    for op, key in ((orc.SUM, "sum"), (orc.MIN, "min"), (orc.MAX, "max"), (orc.COUNT, "count")):
        assert int(ri[key]) == int(orc.aggregate(op, orc.I64, ic)[1]), key
    many = rdf.Column.aggregate_all_many([cia, ca, cib, ci])
    assert [m["rows"] for m in many] == [sum(lens3)] * 4
    for m, full, dt in ((many[0], ia, orc.I64), (many[2], ib, orc.I64), (many[3], ic, orc.I64)):
        for op, key in ((orc.SUM, "sum"), (orc.MIN, "min"), (orc.MAX, "max"), (orc.COUNT, "count")):
            assert int(m[key]) == int(orc.aggregate(op, dt, full)[1]), key
    e2, s2 = orc.sum_exact(orc.F64, fa)
    assert float_ok(many[1]["sum"], e2, s2, sum(lens3), 2.0 ** -53) and many[1]["min"] is None
    fut = rdf.Column.aggregate_all_many([cia, cib], asynchronous=True)
    res = fut.result()
    assert int(res[0]["sum"]) == int(many[0]["sum"]) and int(res[1]["max"]) == int(many[2]["max"])
    # fused expression with a trailing aggregate
    _, ea = rdf.eval_expr_agg([ca, cb], [(N.ADD, 0, 1), ("sin", 2)], materialise=False)
    _, oh = orc.col_unary(orc.SIN, orc.F64, fc)
    e3, s3 = orc.sum_exact(orc.F64, oh)
    assert ea["count"] == sum(c.length - c.null_count for c in oh) and abs(np.longdouble(ea["sum"]) - e3) <= 1e-9 * float(s3)
    checks += 1

    # ---- 4. collective off: per-rank results again -------------------------------------------------------------------
    ctx.comm_collective(False)
    loc = cia.aggregate_all()
    assert loc["rows"] == sum(c.length for c in my(ia))
    assert int(loc["sum"]) == int(orc.aggregate(orc.SUM, orc.I64, my(ia))[1]) if my(ia) else True
    ctx.comm_collective(True)
    checks += 1

    # ---- 5. DivideByZero is agreed on: only ONE rank's shard holds the zero divisor --------------------------------
    num = chunks_for(rdf, "int32", 7, [4096] * world, 0.0)
    den_vals = [np.full(4096, 3, dtype=np.int32) for _ in range(world)]
    den_vals[world - 1][77] = 0
    den = [rdf.PrimitiveArray.from_numpy(v) for v in den_vals]
    cn, cd = rdf.Column.upload(my(num), ctx=ctx, dtype=rdf.I32), rdf.Column.upload(my(den), ctx=ctx, dtype=rdf.I32)
    try:
        cn.divide(cd)
        raise AssertionError(f"rank {rank}: DivideByZero not raised")
    except rdf.DivideByZero:
        pass
    den_vals[world - 1][77] = 5
    den = [rdf.PrimitiveArray.from_numpy(v) for v in den_vals]
    cd2 = rdf.Column.upload(my(den), ctx=ctx, dtype=rdf.I32)
    q = cn.divide(cd2)
    st, oq = orc.col_binary(orc.DIV, orc.I32, num, den)
    assert int(q.sum()) == int(orc.aggregate(orc.SUM, orc.I32, oq)[1])
    checks += 1

    # ---- 6. BASELINE config 3 sharded: 8 x Int64, 10 % nulls, sum/min/max/count -- identical for every N ----------
    n_chunks = 25 if full_size else 5
    chunk = CHUNK if full_size else 400_000
    full_lens = [chunk] * n_chunks
    pieces = parallel.shard_row_ranges(full_lens, rank, world)
    my_lens = [n for _, _, n in pieces]
    row0 = pieces[0][0] * chunk + pieces[0][1] if pieces else 0
    cols = [rdf.Column.generate(rdf.I64, my_lens, 2 if k == 7 else 3, col_id=30 + k, row0=row0, null_mod=10, ctx=ctx) for k in range(8)]
    got = rdf.Column.aggregate_all_many(cols)
    sep = cols[0].aggregate_all()
    assert {k: int(sep[k]) for k in ("sum", "min", "max", "count")} == {k: int(got[0][k]) for k in ("sum", "min", "max", "count")}
    for c in cols:
        c.free()
    # the same columns on ONE GPU (rank 0, collective off): integer aggregates must be identical for every world size
    want = None
    if rank == 0:
        ctx.comm_collective(False)
        want = []
        for k in range(8):
            c = rdf.Column.generate(rdf.I64, full_lens, 2 if k == 7 else 3, col_id=30 + k, row0=0, null_mod=10, ctx=ctx)
            w = c.aggregate_all()
            want.append({key: int(w[key]) for key in ("sum", "min", "max", "count", "rows")})
            c.free()
        ctx.comm_collective(True)
        for k in range(8):
            assert {key: int(got[k][key]) for key in ("sum", "min", "max", "count", "rows")} == want[k], (k, got[k], want[k])
        if not full_size:   # and against the oracle (the full-size single-GPU result is oracle-checked by tests/test_configs_gpu.py)
            for k in range(8):
                o = [orc.generate(orc.I64, 2 if k == 7 else 3, 0, 0, SEED, 30 + k, i * chunk, chunk, 10) for i in range(n_chunks)]
                for op, key in ((orc.SUM, "sum"), (orc.MIN, "min"), (orc.MAX, "max"), (orc.COUNT, "count")):
                    assert int(orc.aggregate(op, orc.I64, o)[1]) == want[k][key], (k, key)
    checks += 1

    # ---- 7. many collectives back to back on both streams (fused futures: finish stream; plain aggregates: compute stream),
    #         then the other transport of the combine: same answers
    ref_many = rdf.Column.aggregate_all_many([cia, cib])
    futs = []
    for i in range(40):
        futs.append(ca.binary_agg_async(N.ADD, cb))
        if i % 3 == 0:
            got = rdf.Column.aggregate_all_many([cia, cib])
            assert [int(g["sum"]) for g in got] == [int(g["sum"]) for g in ref_many]
    sums = set()
    for col, fut in futs:
        sums.add(np.float64(fut.result()["sum"]).view(np.uint64).item())
        col.free()
    assert len(sums) == 1
    before = ctx.comm_get_combine()
    try:
        ctx.comm_set_combine(before != "peer-memory")
        switched = True
    except rdf.ArrowError:
        switched = False   # no peer access on this box: only NCCL
    if switched:
        got = rdf.Column.aggregate_all_many([cia, cib])
        assert [int(g[k]) for g in got for k in ("sum", "min", "max", "count")] == [int(g[k]) for g in ref_many for k in ("sum", "min", "max", "count")]
        col, r = ca.binary_agg(N.ADD, cb)
        assert np.float64(r["sum"]).view(np.uint64).item() in sums   # rank-order fold in both transports: bit-identical
        col.free()
        ctx.comm_set_combine(before == "peer-memory")
    checks += 1

    ctx.comm_barrier()
    ok = ctx.comm_all_reduce([1.0], N.SUM)[0]
    assert ok == world
    if rank == 0:
        print("COMM-OK " + json.dumps({"world": world, "checks": checks, "nccl": info["nccl_version"], "collectives": ctx.comm_info()["collectives"], "combine": ctx.comm_get_combine(),
                                       "config3_sum_col0": want[0]["sum"] if want else None}), flush=True)
    ctx.comm_detach()


if __name__ == "__main__":
    main()
