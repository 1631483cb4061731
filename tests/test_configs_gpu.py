"""BASELINE.json `configs` 2-5 at their full sizes, as parity cases.

Data is generated on the device by the counter-based generator, which the CPU oracle reproduces bit-for-bit
(tests/test_parity_gpu.py::test_generator_matches_oracle), so any chunk can be regenerated on the host and
pushed through the oracle's restatement of the reference path.  Every config checks: sampled chunks in full
(bit-exact for arithmetic/cast, <= 3 ulp for sin), column-wide aggregates against an oracle pass over ALL
chunks (exact for integers/counts, tolerance for float sums), and the validity/null-count bookkeeping.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from helpers import assert_same_array

pytestmark = pytest.mark.gpu

SEED = 20260924
CHUNK = 4_000_000


def gen(oracle, dtype, kind, lo, hi, col, chunk_index, n=CHUNK, null_mod=0):
    return oracle.generate(dtype, kind, lo, hi, SEED, col, chunk_index * n, n, null_mod)


def float_sum_ok(got, exact, sum_abs, n, eps=2.0 ** -53):
    return abs(np.longdouble(got) - exact) <= 16 * np.log2(max(n, 2)) * eps * sum_abs


def test_config2_chain_1e8_f64(rdf, ctx, oracle):
    """1e8 rows, 4 x Float64: e=a+b; f=e*c; g=f/d; h=sin(g); variant with 10% nulls on b and d."""
    lens = [CHUNK] * 25
    C = rdf.Column
    for null_mod in (0, 10):
        a = C.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=20)
        b = C.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=21, null_mod=null_mod)
        c = C.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=22)
        d = C.generate(rdf.F64, lens, 1, col_id=23, null_mod=null_mod)
        e = a.add(b); f = e.multiply(c); g = f.divide(d); h = g.sin()
        hsum, hcount = h.sum(), h.count()
        got_h, got_g = h.download(), None
        exact, sum_abs, count = np.longdouble(0), np.longdouble(0), 0

        def ref_chunk(i):
            oa, ob = gen(oracle, oracle.F64, 0, -1e3, 1e3, 20, i), gen(oracle, oracle.F64, 0, -1e3, 1e3, 21, i, null_mod=null_mod)
            oc, od = gen(oracle, oracle.F64, 0, -1e3, 1e3, 22, i), gen(oracle, oracle.F64, 1, 0, 0, 23, i, null_mod=null_mod)
            _, oe = oracle.col_binary(oracle.ADD, oracle.F64, [oa], [ob])
            _, of = oracle.col_binary(oracle.MUL, oracle.F64, oe, [oc])
            st, og = oracle.col_binary(oracle.DIV, oracle.F64, of, [od])
            assert st == oracle.OK
            _, oh = oracle.col_unary(oracle.SIN, oracle.F64, og)
            return og[0], oh[0]

        with ThreadPoolExecutor(16) as ex:
            refs = list(ex.map(ref_chunk, range(len(lens))))
        for i, (og, oh) in enumerate(refs):
            ex_i, sa_i = oracle.sum_exact(oracle.F64, [oh])
            exact += ex_i; sum_abs += sa_i; count += oh.length - oh.null_count
            assert_same_array(got_h[i], oh, what=f"cfg2 h chunk {i} null_mod={null_mod}", exact=False, max_ulp=3, check_payload=False)
            if i in (0, 13, 24):
                if got_g is None:
                    got_g = g.download()
                assert_same_array(got_g[i], og, what=f"cfg2 g chunk {i}")
        assert hcount == count
        assert float_sum_ok(hsum, exact, sum_abs, 100_000_000)
        for col in (a, b, c, d, e, f, g, h):
            col.free()


def test_config3_int64_aggregates_8_columns(rdf, ctx, oracle):
    """1e8 rows, 8 x Int64 with 10% nulls: sum / min(intended) / max / count per column (one column full range
    to prove wrapping parity).  Exact equality with an oracle pass over all chunks."""
    lens = [CHUNK] * 25
    for k in range(8):
        kind = 2 if k == 7 else 3
        col = rdf.Column.generate(rdf.I64, lens, kind, col_id=30 + k, null_mod=10)
        fused = col.aggregate_all()
        sep = {"sum": col.sum(), "min": col.min(), "max": col.max(), "count": col.count()}

        def ref_chunk(i, k=k, kind=kind):
            o = gen(oracle, oracle.I64, kind, 0, 0, 30 + k, i, null_mod=10)
            return (int(oracle.aggregate(oracle.SUM, oracle.I64, [o])[1]), int(oracle.aggregate(oracle.MIN, oracle.I64, [o])[1]),
                    int(oracle.aggregate(oracle.MAX, oracle.I64, [o])[1]), int(oracle.aggregate(oracle.COUNT, oracle.I64, [o])[1]))

        with ThreadPoolExecutor(16) as ex:
            parts = list(ex.map(ref_chunk, range(len(lens))))
        total = sum(p[0] for p in parts) & ((1 << 64) - 1)
        total = total - (1 << 64) if total >= 1 << 63 else total  # the reference folds chunk sums with wrapping adds
        want = {"sum": total, "min": min(p[1] for p in parts), "max": max(p[2] for p in parts), "count": sum(p[3] for p in parts)}
        for key in ("sum", "min", "max", "count"):
            assert int(fused[key]) == want[key] == int(sep[key]), (k, key, fused[key], sep[key], want[key])
        assert not fused["would_panic"] and fused["rows"] == 100_000_000
        col.free()


def test_config4_cast_chain(rdf, ctx, oracle):
    """1e8 rows: x Int32 (full range, 10% nulls), y Float64: xf=cast(x,Float64); z=xf+y; w=z*y; s=sum(w)."""
    lens = [CHUNK] * 25
    x = rdf.Column.generate(rdf.I32, lens, 2, col_id=40, null_mod=10)
    y = rdf.Column.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=41)
    xf = x.cast(rdf.F64); z = xf.add(y)
    w, agg = z.binary_agg(rdf.native.MUL, y)     # fused multiply + aggregate
    s_two_pass = w.sum()
    got = {"xf": xf.download(), "w": w.download()}
    exact, sum_abs, count = np.longdouble(0), np.longdouble(0), 0

    def ref_chunk(i):
        ox, oy = gen(oracle, oracle.I32, 2, 0, 0, 40, i, null_mod=10), gen(oracle, oracle.F64, 0, -1e3, 1e3, 41, i)
        _, oxf = oracle.col_cast(oracle.I32, oracle.F64, [ox])
        _, oz = oracle.col_binary(oracle.ADD, oracle.F64, oxf, [oy])
        _, ow = oracle.col_binary(oracle.MUL, oracle.F64, oz, [oy])
        return oxf[0], ow[0]

    with ThreadPoolExecutor(16) as ex:
        refs = list(ex.map(ref_chunk, range(len(lens))))
    for i, (oxf, ow) in enumerate(refs):
        assert_same_array(got["xf"][i], oxf, what=f"cfg4 xf chunk {i}")
        assert_same_array(got["w"][i], ow, what=f"cfg4 w chunk {i}")
        e, sa = oracle.sum_exact(oracle.F64, [ow])
        exact += e; sum_abs += sa; count += ow.length - ow.null_count
    assert agg["count"] == count == w.count()
    assert float_sum_ok(agg["sum"], exact, sum_abs, 100_000_000) and float_sum_ok(s_two_pass, exact, sum_abs, 100_000_000)
    for col in (x, y, xf, z, w):
        col.free()


def test_config5_full_pipeline_256_batches(rdf, ctx, oracle):
    """Vec<RecordBatch> of 256 x 4e6 rows (1.024e9 rows; i0,i1 Int64 and f0,f1 Float64 with 5% nulls):
    f2=f0+f1; f3=f2*f0; i2=i0+i1; f4=cast(i2,Float64); f5=f3/f1; f6=sin(f5); f7=f6+f4;
    then sum/min/max/count(i2), sum(f7), count(f7).  One B200 holds the whole frame (~100 GB live)."""
    n_batches = 256
    lens = [CHUNK] * n_batches
    G = rdf.Column.generate
    i0 = G(rdf.I64, lens, 3, col_id=50, null_mod=20); i1 = G(rdf.I64, lens, 3, col_id=51, null_mod=20)
    f0 = G(rdf.F64, lens, 0, -1e3, 1e3, col_id=52, null_mod=20); f1 = G(rdf.F64, lens, 1, col_id=53, null_mod=20)
    f2 = f0.add(f1); f3 = f2.multiply(f0); f2.free()
    i2, i2_agg = i0.binary_agg(rdf.native.ADD, i1)
    f4 = i2.cast(rdf.F64)
    f5 = f3.divide(f1); f3.free()
    f6 = f5.sin(); f5.free()
    f7, f7_agg = f6.binary_agg(rdf.native.ADD, f4); f6.free()
    assert f7.n_chunks == n_batches and len(f7) == n_batches * CHUNK
    sample = [0, 100, 255]
    got_f7 = {i: None for i in sample}
    allf7 = None

    def ref_chunk(i):
        oi0, oi1 = gen(oracle, oracle.I64, 3, 0, 0, 50, i, null_mod=20), gen(oracle, oracle.I64, 3, 0, 0, 51, i, null_mod=20)
        of0, of1 = gen(oracle, oracle.F64, 0, -1e3, 1e3, 52, i, null_mod=20), gen(oracle, oracle.F64, 1, 0, 0, 53, i, null_mod=20)
        _, of2 = oracle.col_binary(oracle.ADD, oracle.F64, [of0], [of1])
        _, of3 = oracle.col_binary(oracle.MUL, oracle.F64, of2, [of0])
        _, oi2 = oracle.col_binary(oracle.ADD, oracle.I64, [oi0], [oi1])
        _, of4 = oracle.col_cast(oracle.I64, oracle.F64, oi2)
        st, of5 = oracle.col_binary(oracle.DIV, oracle.F64, of3, [of1])
        assert st == oracle.OK
        _, of6 = oracle.col_unary(oracle.SIN, oracle.F64, of5)
        _, of7 = oracle.col_binary(oracle.ADD, oracle.F64, of6, of4)
        e, sa = oracle.sum_exact(oracle.F64, of7)
        agg_i2 = (int(oracle.aggregate(oracle.SUM, oracle.I64, oi2)[1]), int(oracle.aggregate(oracle.MIN, oracle.I64, oi2)[1]),
                  int(oracle.aggregate(oracle.MAX, oracle.I64, oi2)[1]), int(oracle.aggregate(oracle.COUNT, oracle.I64, oi2)[1]))
        keep = of7[0] if i in sample else None
        return e, sa, of7[0].length - of7[0].null_count, agg_i2, keep

    with ThreadPoolExecutor(32) as ex:
        refs = list(ex.map(ref_chunk, range(n_batches)))
    exact = sum((r[0] for r in refs), np.longdouble(0)); sum_abs = sum((r[1] for r in refs), np.longdouble(0))
    f7_count = sum(r[2] for r in refs)
    i2_sum = sum(r[3][0] for r in refs) & ((1 << 64) - 1)
    i2_sum = i2_sum - (1 << 64) if i2_sum >= 1 << 63 else i2_sum
    assert int(i2_agg["sum"]) == i2_sum and int(i2_agg["min"]) == min(r[3][1] for r in refs)
    assert int(i2_agg["max"]) == max(r[3][2] for r in refs) and i2_agg["count"] == sum(r[3][3] for r in refs)
    assert i2.aggregate_all()["sum"] == i2_agg["sum"]
    assert f7_agg["count"] == f7_count == f7.count()
    # f7 = sin(..) + i2 as f64: |i2| <~ 2^41, so the sum is dominated by f4; tolerance relative to sum|x| as everywhere
    assert float_sum_ok(f7_agg["sum"], exact, sum_abs, n_batches * CHUNK)
    allf7 = f7.download()
    for i in sample:
        # sin(f5) carries <= 3 ulp of |sin| <= 1 into a sum of magnitude up to 2^41: compare f7 with absolute slack
        want = refs[i][4]
        g = allf7[i]
        assert np.array_equal(g.valid_mask(), want.valid_mask()) and g.null_count == want.null_count
        m = want.valid_mask()
        diff = np.abs(g.value_slice()[m] - want.values[m])
        assert diff.max() <= 3 * 2.0 ** -52 + np.spacing(np.abs(want.values[m])).max(), f"cfg5 f7 chunk {i}: {diff.max()}"
    for col in (i0, i1, f0, f1, i2, f4, f7):
        col.free()


def test_single_huge_chunk_and_many_tiny_chunks(rdf, ctx, oracle):
    """Two extremes of the Vec<RecordBatch> shape: one 5e7-row chunk (every tile of every kernel lands in the same
    chunk descriptor; the filter scan walks 24k tiles in one CTA) and 4000 ragged chunks of 0..3000 rows (the
    per-CTA descriptor search, tail tiles everywhere; the reference's own benchmark is 380 x 5 rows)."""
    # --- one huge chunk ---
    n = 50_000_000
    a = rdf.Column.generate(rdf.I64, [n], 3, col_id=70, null_mod=10)
    b = rdf.Column.generate(rdf.I64, [n], 3, col_id=71)
    c, agg = a.binary_agg(rdf.native.ADD, b)
    oa = oracle.generate(oracle.I64, 3, 0, 0, SEED, 70, 0, n, 10)
    ob = oracle.generate(oracle.I64, 3, 0, 0, SEED, 71, 0, n)
    _, oc = oracle.col_binary(oracle.ADD, oracle.I64, [oa], [ob])
    assert_same_array(c.download()[0], oc[0], what="huge chunk add")
    for key, op in (("sum", oracle.SUM), ("min", oracle.MIN), ("max", oracle.MAX), ("count", oracle.COUNT)):
        want = int(oracle.aggregate(op, oracle.I64, oc)[1])
        assert int(agg[key]) == want == int(c.aggregate_all()[key]), key
    mask = c.gt(0.0)
    kept = c.filter(mask)
    _, om = oracle.compare(oracle.GT, oc[0], None, scalar=0.0)
    _, ok = oracle.filter_chunk(oc[0], om)
    assert_same_array(kept.download()[0], ok, what="huge chunk filter")
    for col in (a, b, c, mask, kept):
        col.free()
    # --- thousands of tiny ragged chunks ---
    rng = np.random.default_rng(123)
    lens = [int(x) for x in rng.integers(0, 3000, 4000)]
    lens[:6] = [0, 0, 1, 2049, 0, 5]
    x = rdf.Column.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=72, null_mod=7)
    y = rdf.Column.generate(rdf.F64, lens, 1, col_id=73, null_mod=5)
    q, qagg = x.binary_agg(rdf.native.DIV, y)
    qi = q.cast(rdf.I32)
    keep = q.filter(q.lt(0.0))
    got_q, got_qi, got_keep = q.download(), qi.download(), keep.download()
    row, exact, sum_abs, count = 0, np.longdouble(0), np.longdouble(0), 0
    for i, m in enumerate(lens):
        ox = oracle.generate(oracle.F64, 0, -1e3, 1e3, SEED, 72, row, m, 7)
        oy = oracle.generate(oracle.F64, 1, 0, 0, SEED, 73, row, m, 5)
        row += m
        st, oq = oracle.col_binary(oracle.DIV, oracle.F64, [ox], [oy])
        assert st == oracle.OK
        assert_same_array(got_q[i], oq[0], what=f"tiny chunk {i} divide")
        if i % 40 == 0:
            _, oqi = oracle.col_cast(oracle.F64, oracle.I32, oq)
            assert_same_array(got_qi[i], oqi[0], what=f"tiny chunk {i} cast")
            _, om = oracle.compare(oracle.LT, oq[0], None, scalar=0.0)
            _, ok = oracle.filter_chunk(oq[0], om)
            assert_same_array(got_keep[i], ok, what=f"tiny chunk {i} filter")
        e, sa = oracle.sum_exact(oracle.F64, oq)
        exact += e; sum_abs += sa; count += oq[0].length - oq[0].null_count
    assert qagg["count"] == count == q.count()
    assert float_sum_ok(qagg["sum"], exact, sum_abs, sum(lens)) and float_sum_ok(q.sum(), exact, sum_abs, sum(lens))
