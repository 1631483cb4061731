"""The lazy-evaluator mirror (rust-dataframe_b200/frame.py): Evaluate::evaluate / calculate (src/evaluation.rs:66-323)
over device columns, and the fusion pass of SURVEY 8(f) N3.  The planner is host logic (no GPU); the data tests
compare a fused evaluation with the step-by-step one (bit for bit) and with the oracle."""
from collections import OrderedDict

import numpy as np
import pytest

from helpers import assert_same_array, random_mask


def chain():
    from rust_dataframe_b200.frame import calculate
    return [calculate("add", ["a", "b"], "e"), calculate("multiply", ["e", "c"], "f"), calculate("divide", ["f", "d"], "g"),
            calculate("sine", ["g"], "h")]


def schema_of(rdf, **cols):
    return OrderedDict((k, getattr(rdf, v)) for k, v in cols.items())


def kinds(plan):
    return [k for k, _ in plan]


def test_planner_fuses_only_what_nobody_can_observe(rdf):
    F = rdf.frame
    sch = schema_of(rdf, a="F64", b="F64", c="F64", d="F64")
    plan = F.plan_fusion(sch, chain() + [F.select(["h"])])
    assert kinds(plan) == ["fused", "select"]
    fused = plan[0][1]
    assert fused.inputs == ["a", "b", "c", "d"] and fused.output == "h" and len(fused.replaced) == 4
    N = rdf.native
    assert fused.nodes == [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)]
    # nothing removes e, f, g: they are columns of the result, every Calculation stays
    assert kinds(F.plan_fusion(sch, chain())) == ["calculate"] * 4
    # drop instead of select; an unrelated sort and filter in between do not keep the intermediates alive
    plan = F.plan_fusion(sch, chain() + [F.sort([("h", True)]), F.filter_(("gt", F.col("a"), F.lit(0.0))), F.drop(["e", "f", "g"])])
    assert kinds(plan) == ["fused", "sort", "filter", "drop"]
    # ... but a filter that READS g does: the run is cut after g, which stays a real column
    plan = F.plan_fusion(sch, chain() + [F.filter_(("lt", F.col("g"), F.lit(1.0))), F.select(["h"])])
    assert kinds(plan) == ["fused", "calculate", "filter", "select"]
    assert plan[0][1].output == "g" and len(plan[0][1].replaced) == 3 and plan[1][1].function == "sine"
    # f kept by the final select: e can still be fused away into f; g is fused into h
    plan = F.plan_fusion(sch, chain() + [F.select(["f", "h"])])
    assert kinds(plan) == ["fused", "fused", "select"]
    assert [p[1].output for p in plan[:2]] == ["f", "h"] and plan[1][1].inputs == ["f", "d"]


def test_planner_type_and_size_rules(rdf):
    F = rdf.frame
    # integer arithmetic and casts are not fusable; the Float64 tail after the cast is
    sch = schema_of(rdf, i="I64", j="I64", x="F64")
    steps = [F.calculate("add", ["i", "j"], "k"), F.calculate("cast", ["k"], "kf", rdf.F64), F.calculate("multiply", ["kf", "x"], "m"),
             F.calculate("cosine", ["m"], "n"), F.select(["n"])]
    plan = F.plan_fusion(sch, steps, fuse_casts=False)
    assert kinds(plan) == ["calculate", "calculate", "fused", "select"] and plan[2][1].inputs == ["kf", "x"]
    plan = F.plan_fusion(sch, steps)
    assert kinds(plan) == ["calculate", "fused", "select"] and plan[1][1].inputs == ["k", "x"]      # the cast rides in the fused load
    assert [c.function for c in plan[1][1].replaced] == ["cast", "multiply", "cosine"] and len(plan[1][1].nodes) == 2
    # a cast to anything but Float64, or one whose result stays visible, is a real Calculation
    assert kinds(F.plan_fusion(sch, [F.calculate("cast", ["i"], "i32", rdf.I32), F.calculate("cast", ["i32"], "f", rdf.F64), F.calculate("sine", ["f"], "s"),
                                     F.select(["s", "f"])])) == ["calculate", "calculate", "calculate", "select"]
    # Float32 chains are left alone (the fused kernel is Float64)
    assert kinds(F.plan_fusion(schema_of(rdf, a="F32", b="F32"), [F.calculate("add", ["a", "b"], "e"), F.calculate("sine", ["e"], "h"), F.select(["h"])])) \
        == ["calculate", "calculate", "select"]
    # more than 12 nodes: the run is split, every piece still correct on its own
    sch = schema_of(rdf, a="F64", b="F64")
    long = [F.calculate("add", ["a", "b"], "t0")] + [F.calculate("multiply", [f"t{k}", "b"], f"t{k + 1}") for k in range(15)] + [F.select(["t15"])]
    plan = F.plan_fusion(sch, long)
    assert kinds(plan) == ["fused", "fused", "select"] and len(plan[0][1].replaced) + len(plan[1][1].replaced) == 16
    assert all(len(p[1].nodes) <= 12 for p in plan[:2])
    # more than 6 distinct inputs
    sch = schema_of(rdf, **{f"c{k}": "F64" for k in range(9)})
    wide = [F.calculate("add", ["c0", "c1"], "s1")] + [F.calculate("add", [f"s{k}", f"c{k + 1}"], f"s{k + 1}") for k in range(1, 8)] + [F.select(["s8"])]
    plan = F.plan_fusion(sch, wide)
    assert all(k in ("fused", "calculate", "select") for k in kinds(plan))
    assert all(len(p[1].inputs) <= 6 for p in plan if p[0] == "fused") and kinds(plan)[0] == "fused"
    # an intermediate that overwrites an input of the run, a name produced twice, a dead intermediate: not fused
    sch = schema_of(rdf, a="F64", b="F64")
    assert "fused" not in kinds(F.plan_fusion(sch, [F.calculate("add", ["a", "b"], "a"), F.calculate("sine", ["a"], "h"), F.select(["h"])]))
    assert "fused" not in kinds(F.plan_fusion(sch, [F.calculate("add", ["a", "b"], "e"), F.calculate("subtract", ["a", "b"], "e"), F.select(["e"])]))
    assert "fused" not in kinds(F.plan_fusion(sch, [F.calculate("divide", ["a", "b"], "e"), F.calculate("sine", ["a"], "h"), F.select(["h"])]))
    # rename and select("*")
    plan = F.plan_fusion(sch, [F.calculate("add", ["a", "b"], "e"), F.calculate("sine", ["e"], "h"), F.calculate("rename", ["h"], "out"), F.select(["*"])])
    assert kinds(plan) == ["calculate", "calculate", "calculate", "select"]      # e is still visible through "*"


def test_expr_check_and_planner_respect_the_compiler_limits(rdf):
    """bdf_expr_check runs the library's host compiler without a GPU; the planner asks it before it fuses."""
    import ctypes as C

    F, N = rdf.frame, rdf.native
    from rust_dataframe_b200.functions import _expr_nodes

    def check(n_inputs, prog):
        ni, nt = C.c_int32(-1), C.c_int32(-1)
        st = N.lib().bdf_expr_check(n_inputs, None, len(prog), _expr_nodes(prog), C.byref(ni), C.byref(nt))
        return st, ni.value, nt.value

    assert check(4, [(N.ADD, 0, 1), (N.MUL, 4, 2), (N.DIV, 5, 3), ("sin", 6)]) == (N.OK, 5, 0)        # load, add, mul, div, sin
    assert check(4, [(N.ADD, 0, 1), (N.SUB, 2, 3), (N.MUL, 4, 5)])[::2] == (N.OK, 1)                   # one side parked in a temporary
    assert check(2, [(N.ADD, 0, 1), (N.MUL, 2, 2), (N.SUB, 0, 2), (N.ADD, 3, 4)])[0] == N.OK             # a shared node
    assert check(2, [(N.ADD, 0, 2)])[0] == N.INVALID                                                   # forward reference
    assert check(2, [(N.DIV, 0, 1), (N.ADD, 0, 1)])[0] == N.INVALID                                    # dead node
    assert check(2, [(99, 0, 1)])[0] == N.INVALID and check(7, [(N.ADD, 0, 1)])[0] == N.INVALID
    three_shared = [(N.ADD, 0, 1), (N.SUB, 0, 1), (N.MUL, 0, 1), (N.MUL, 2, 3), (N.MUL, 5, 4), (N.ADD, 6, 2), (N.ADD, 7, 3), (N.ADD, 8, 4)]
    assert check(2, three_shared)[0] == N.UNSUPPORTED
    # the same shape as Calculations: the planner must not emit it as one step, and what it emits must compile
    sch = schema_of(rdf, a="F64", b="F64")
    steps = [F.calculate("add", ["a", "b"], "s1"), F.calculate("subtract", ["a", "b"], "s2"), F.calculate("multiply", ["a", "b"], "s3"),
             F.calculate("multiply", ["s1", "s2"], "t"), F.calculate("multiply", ["t", "s3"], "u"), F.calculate("add", ["u", "s1"], "v"),
             F.calculate("add", ["v", "s2"], "w"), F.calculate("add", ["w", "s3"], "z"), F.select(["z"])]
    plan = F.plan_fusion(sch, steps)
    assert sum(len(p[1].replaced) if p[0] == "fused" else 1 for p in plan if p[0] != "select") == 8
    for kind, arg in plan:
        if kind == "fused":
            assert len(arg.replaced) < 8 and check(len(arg.inputs), arg.nodes)[0] == N.OK


# ---- data -----------------------------------------------------------------------------------------------

def host_frame(rdf, rng, lens, nulls=True):
    def colf(lo, hi, nf):
        out = []
        for n in lens:
            a = rdf.PrimitiveArray.from_numpy(rng.uniform(lo, hi, n), random_mask(rng, n, nf) if nf else None)
            out.append(a)
        return out
    return OrderedDict([("a", colf(-50, 50, 0)), ("b", colf(-50, 50, 0.2 if nulls else 0)), ("c", colf(-5, 5, 0)), ("d", colf(1, 9, 0.1 if nulls else 0)),
                        ("k", [rdf.PrimitiveArray.from_numpy(rng.integers(-3, 4, n).astype(np.int32), random_mask(rng, n, 0.1)) for n in lens])])


def frames_equal(x, y):
    assert list(x.columns) == list(y.columns)
    hx, hy = x.to_host(), y.to_host()
    for name in hx:
        assert len(hx[name]) == len(hy[name]), name
        for cx, cy in zip(hx[name], hy[name]):
            assert cx.length == cy.length, name
            m = cx.valid_mask()
            assert np.array_equal(m, cy.valid_mask()), name
            vx, vy = cx.value_slice()[m], cy.value_slice()[m]
            assert np.array_equal(vx.view(f"u{vx.dtype.itemsize}"), vy.view(f"u{vy.dtype.itemsize}")), name


@pytest.mark.gpu
def test_fused_and_stepwise_evaluation_agree(rdf, ctx, oracle):
    F = rdf.frame
    rng = np.random.default_rng(8)
    lens = [0, 1, 2049, 70_001]
    host = host_frame(rdf, rng, lens)
    frame = rdf.DeviceFrame.from_host(host)
    pipelines = [
        chain() + [F.select(["a", "h"])],
        chain() + [F.filter_(("and", ("gt", F.col("h"), F.lit(-0.5)), ("not", ("lt", F.col("k"), F.lit(0.0))))), F.sort([("k", True), ("h", False)]),
                   F.drop(["e", "f", "g", "b"])],
        [F.calculate("cast", ["k"], "kf", rdf.F64), F.calculate("multiply", ["kf", "a"], "m"), F.calculate("subtract", ["m", "c"], "n"),
         F.calculate("tangent", ["n"], "t"), F.limit(1000), F.select(["k", "t"])],                      # the cast is fused too
        [F.calculate("cast", ["k"], "kf", rdf.F64), F.calculate("add", ["kf", "b"], "s"), F.select(["kf", "s"])],   # kf stays visible: no fusion
        chain() + [F.calculate("rename", ["h"], "out"), F.select(["out", "f"])],
    ]
    for steps in pipelines:
        plan = F.plan_fusion(frame.schema, steps)
        assert ("fused" in [k for k, _ in plan]) == ("kf" not in steps[-1][1])
        frames_equal(frame.evaluate(steps, fuse=True), frame.evaluate(steps, fuse=False))
        frames_equal(frame.evaluate(steps, fuse=True, fuse_casts=False), frame.evaluate(steps, fuse=False))
    # against the oracle: the first pipeline, chunk by chunk
    got = frame.evaluate(pipelines[0]).to_host()
    assert list(got) == ["a", "h"]
    for i in range(len(lens)):
        _, e = oracle.col_binary(oracle.ADD, oracle.F64, [host["a"][i]], [host["b"][i]])
        _, f = oracle.col_binary(oracle.MUL, oracle.F64, e, [host["c"][i]])
        st, g = oracle.col_binary(oracle.DIV, oracle.F64, f, [host["d"][i]])
        assert st == oracle.OK
        _, h = oracle.col_unary(oracle.SIN, oracle.F64, g)
        assert_same_array(got["h"][i], h[0], what=f"pipeline h chunk {i}", exact=False, max_ulp=3, check_payload=False)


@pytest.mark.gpu
def test_evaluator_rules_and_arrow_round_trip(rdf, ctx, tmp_path):
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.ipc

    F = rdf.frame
    rng = np.random.default_rng(4)
    n = 50_000
    t = pa.table({"a": pa.array(rng.normal(0, 10, n)), "b": pa.array(rng.normal(0, 10, n), mask=rng.random(n) < 0.1),
                  "name": pa.array([f"r{i}" for i in range(n)]), "k": pa.array(rng.integers(0, 5, n).astype(np.int8))})
    src, dst = str(tmp_path / "in.arrow"), str(tmp_path / "out.arrow")
    with pa.ipc.new_file(src, t.schema) as w:
        for b in t.to_batches(max_chunksize=12_345):
            w.write_batch(b)
    frame = rdf.DeviceFrame.from_arrow(src)
    assert list(frame.columns) == ["a", "b", "k"]                       # the Utf8 column stays with arrow's reader
    out = frame.evaluate([F.calculate("subtract", ["a", "b"], "d"), F.calculate("multiply", ["d", "d"], "sq"), F.filter_(("ge", F.col("sq"), F.lit(1.0))),
                          F.sort([("k", False), ("sq", True)]), F.select(["k", "sq"])])
    out.to_arrow(dst)
    got = pa.ipc.open_file(dst).read_all()
    d = pc.subtract(t["a"], t["b"])
    sq = pc.multiply(d, d)
    want = pa.table({"k": t["k"], "sq": sq}).filter(pc.greater_equal(sq, 1.0), null_selection_behavior="drop")
    want = want.take(pc.sort_indices(want, sort_keys=[("k", "ascending"), ("sq", "descending")], null_placement="at_end"))
    assert got.column("k").combine_chunks().equals(want.column("k").combine_chunks())
    assert got.column("sq").combine_chunks().equals(want.column("sq").combine_chunks())
    # the reference's panics, mirrored
    with pytest.raises(rdf.ReferencePanic):
        frame.evaluate([F.calculate("add", ["k", "k"], "kk")])          # Int8: "Unsupported operation"
    with pytest.raises(rdf.ReferencePanic):
        frame.evaluate([F.calculate("sine", ["k"], "s")])               # "Expecting float datatype"
    with pytest.raises(rdf.ReferencePanic):
        frame.evaluate([("group_aggregate", None)])                     # "aggregations not supported"
    with pytest.raises(rdf.ReferencePanic):
        frame.evaluate([F.calculate("add", ["a", "nope"], "x")])
    with pytest.raises(rdf.DivideByZero):
        frame.evaluate([F.calculate("subtract", ["a", "a"], "z"), F.calculate("divide", ["a", "z"], "q"), F.select(["q"])])
    assert list(frame.evaluate([F.calculate("add", ["a", "b"], "a")]).columns) == ["b", "k", "a"]   # with_column: drop + append


# ---- the planner's claim, checked on random pipelines without a GPU ---------------------------------------------------
# A tiny numpy interpreter of the transformation list (values + validity per column, one chunk) stands in for the
# device: what matters here is WHICH columns exist, in which order, with which values -- fused plan vs original list.

class _NpFrame:
    def __init__(self, cols):
        self.cols = OrderedDict(cols)      # name -> (values float64 or int64, valid bool)

    def run(self, steps):
        import math
        f = _NpFrame(self.cols)
        for kind, arg in steps:
            if kind == "calculate":
                f = f._calc(arg.function, arg.inputs, arg.output, arg.dtype)
            elif kind == "fused":
                slots = [f.cols[n] for n in arg.inputs]
                slots = [(v.astype(np.float64), m) for v, m in slots]
                for nd in arg.nodes:
                    if len(nd) == 2:
                        v, m = slots[nd[1]]
                        slots.append(({"sin": np.sin, "cos": np.cos, "tan": np.tan}[nd[0]](v), m))
                    else:
                        (a, ma), (b, mb) = slots[nd[1]], slots[nd[2]]
                        with np.errstate(all="ignore"):
                            r = [a + b, a - b, a * b, a / np.where(ma & mb, b, 1.0)][nd[0]]
                        slots.append((r, ma & mb))
                cols = OrderedDict((n, c) for n, c in f.cols.items() if n != arg.output)
                cols[arg.output] = slots[-1]
                f = _NpFrame(cols)
            elif kind == "select":
                f = _NpFrame((n, c) for n, c in f.cols.items() if "*" in arg or n in arg)
            elif kind == "drop":
                f = _NpFrame((n, c) for n, c in f.cols.items() if n not in arg)
            elif kind == "filter":
                _, (_, name), (_, thr) = arg
                v, m = f.cols[name]
                keep = m & (v.astype(np.float64) > thr)
                f = _NpFrame((n, (cv[keep], cm[keep])) for n, (cv, cm) in f.cols.items())
            elif kind == "sort":
                name, desc = arg[0]
                v, m = f.cols[name]
                key = np.where(m, -v if desc else v, np.inf)
                order = np.argsort(key, kind="stable")
                f = _NpFrame((n, (cv[order], cm[order])) for n, (cv, cm) in f.cols.items())
        return f

    def _calc(self, fn, inputs, out, dtype):
        if fn == "rename":
            return _NpFrame(((out if n == inputs[0] else n), c) for n, c in self.cols.items())
        a, ma = self.cols[inputs[0]]
        if fn == "cast":
            r, m = a.astype(np.float64), ma
        elif fn in ("sine", "cosine", "tangent"):
            r, m = {"sine": np.sin, "cosine": np.cos, "tangent": np.tan}[fn](a), ma
        else:
            b, mb = self.cols[inputs[1]]
            m = ma & mb
            with np.errstate(all="ignore"):
                if a.dtype.kind == "i":
                    r = {"add": a + b, "subtract": a - b, "multiply": a * b, "divide": a // np.where(m & (b != 0), b, 1)}[fn]
                else:
                    r = {"add": a + b, "subtract": a - b, "multiply": a * b, "divide": a / np.where(m, b, 1.0)}[fn]
        cols = OrderedDict((n, c) for n, c in self.cols.items() if n != out)
        cols[out] = (r, m)
        return _NpFrame(cols)


def test_random_pipelines_fused_plan_is_equivalent(rdf):
    F = rdf.frame
    rng = np.random.default_rng(2024)
    n = 64
    base = OrderedDict([("a", (rng.normal(0, 3, n), np.ones(n, bool))), ("b", (rng.normal(0, 3, n), rng.random(n) > 0.2)),
                        ("c", (rng.uniform(1, 4, n), np.ones(n, bool))), ("i", (rng.integers(-5, 6, n), rng.random(n) > 0.1))])
    schema = OrderedDict([("a", rdf.F64), ("b", rdf.F64), ("c", rdf.F64), ("i", rdf.I64)])
    fused_plans = 0
    for trial in range(600):
        names = {"a": rdf.F64, "b": rdf.F64, "c": rdf.F64, "i": rdf.I64}
        steps, fresh = [], 0
        for _ in range(int(rng.integers(2, 12))):
            floats = [k for k, t in names.items() if t == rdf.F64]
            r = rng.random()
            if r < 0.55 and len(floats) >= 2:
                out = f"t{fresh}" if rng.random() < 0.85 else str(rng.choice(list(names)))   # sometimes overwrite an existing column
                fresh += 1
                if rng.random() < 0.7:
                    x, y = rng.choice(floats, 2)
                    if out in (x, y) and rng.random() < 0.5:
                        out = f"t{fresh}"; fresh += 1
                    steps.append(F.calculate(str(rng.choice(["add", "subtract", "multiply"])), [str(x), str(y)], out))
                else:
                    steps.append(F.calculate(str(rng.choice(["sine", "cosine"])), [str(rng.choice(floats))], out))
                names[out] = rdf.F64
            elif r < 0.62 and "i" in names and names["i"] == rdf.I64:
                steps.append(F.calculate("cast", ["i"], f"t{fresh}", rdf.F64)); names[f"t{fresh}"] = rdf.F64; fresh += 1
            elif r < 0.72 and len(names) > 2:
                victims = [str(v) for v in rng.choice(list(names), int(rng.integers(1, 3)), replace=False)]
                steps.append(F.drop(victims))
                for v in victims:
                    names.pop(v)
            elif r < 0.80 and len(names) > 2:
                keep = [str(v) for v in rng.choice(list(names), int(rng.integers(1, len(names))), replace=False)]
                steps.append(F.select(keep))
                names = {k: t for k, t in names.items() if k in keep}
            elif r < 0.88 and floats:
                steps.append(F.filter_(("gt", F.col(str(rng.choice(floats))), F.lit(float(rng.normal())))))
            elif r < 0.94 and floats:
                steps.append(F.sort([(str(rng.choice(floats)), bool(rng.random() < 0.5))]))
            elif len(names) > 1:
                old = str(rng.choice(list(names)))
                steps.append(F.calculate("rename", [old], f"r{fresh}")); names[f"r{fresh}"] = names.pop(old); fresh += 1
        if rng.random() < 0.6 and len(names) > 1:      # the usual end of a lazy pipeline: keep a few columns, the rest is dead
            keep = [str(v) for v in rng.choice(list(names), int(rng.integers(1, 3)), replace=False)]
            if fresh and f"t{fresh - 1}" in names:
                keep.append(f"t{fresh - 1}")
            steps.append(F.select(keep))
        plan = F.plan_fusion(schema, steps)
        fused_plans += any(k == "fused" for k, _ in plan)
        want, got = _NpFrame(base).run(steps), _NpFrame(base).run(plan)
        assert list(got.cols) == list(want.cols), (trial, steps, plan)
        for name in want.cols:
            (wv, wm), (gv, gm) = want.cols[name], got.cols[name]
            assert np.array_equal(wm, gm), (trial, name)
            assert np.array_equal(wv[wm], gv[gm], equal_nan=True) if wv.dtype.kind == "f" else np.array_equal(wv[wm], gv[gm]), (trial, name, steps, plan)
    assert fused_plans > 60      # the generator does produce fusable runs (99 of 600 with this seed)
