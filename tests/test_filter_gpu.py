"""N2: BooleanFilter comparisons, boolean kernels and ChunkedArray::filter on the GPU vs the oracle
(which tests/test_oracle_golden.py cross-checks against pyarrow).  Bit-exact everywhere: comparison results,
validity, kept values, kept validity, output lengths and null counts."""
import numpy as np
import pytest

from helpers import assert_same_array, random_mask

pytestmark = pytest.mark.gpu

RAGGED = [0, 1, 31, 32, 33, 2047, 2048, 2049, 10007, 70001]


def make_col(rdf, rng, dtype, lens, null_frac, sliced, lo=-50, hi=50):
    out = []
    for k, n in enumerate(lens):
        pad = (5 + 3 * k) % 23 if sliced else 0
        npdt = np.dtype(rdf.NP_DTYPES[dtype])
        if npdt.kind == "f":
            v = rng.uniform(lo, hi, n + pad + 3).astype(npdt)
            v[::97] = 0.0
        else:
            v = rng.integers(max(lo, np.iinfo(npdt).min), min(hi, np.iinfo(npdt).max), n + pad + 3).astype(npdt)
        a = rdf.PrimitiveArray.from_numpy(v, random_mask(rng, n + pad + 3, null_frac) if null_frac else None)
        a.null_count = -1 if a.validity is not None else 0
        out.append(a.slice(pad, n))
    return out


def check_bool(got, want, what):
    assert got.length == want.length, what
    gm, wm = got.valid_mask(), want.valid_mask()
    assert np.array_equal(gm, wm), f"{what}: validity"
    assert np.array_equal(got.value_bits(), want.value_bits()), f"{what}: values"  # computed under nulls too
    assert (got.null_count if got.validity is not None else 0) == int((~wm).sum()), f"{what}: null_count"
    if got.length % 8:  # zero padding bits
        assert got.values[(got.length - 1) // 8] >> (got.length % 8) == 0, f"{what}: padding bits"


@pytest.mark.parametrize("ltype,rtype", [("F64", "F64"), ("I32", "F64"), ("I64", "I8"), ("U16", "F32"), ("F32", "I64")])
def test_compare_matches_oracle(rdf, ctx, oracle, ltype, rtype):
    lt, rt = getattr(rdf, ltype), getattr(rdf, rtype)
    rng = np.random.default_rng(lt * 16 + rt)
    for nl, nr, sliced in ((0, 0, False), (0.2, 0, True), (0.2, 0.3, True)):
        a = make_col(rdf, rng, lt, RAGGED, nl, sliced)
        b = make_col(rdf, rng, rt, RAGGED, nr, sliced)
        if ltype == "F64":
            a[-1].values[a[-1].offset:a[-1].offset + 4] = [np.nan, np.inf, -np.inf, -0.0]
        ca, cb = rdf.Column.upload(a), rdf.Column.upload(b)
        for op in range(6):
            got = ca.compare(op, cb).download()
            for i, g in enumerate(got):
                st, want = oracle.compare(op, a[i], b[i])
                assert st == oracle.OK
                check_bool(g, want, f"compare op{op} {ltype},{rtype} chunk {i} nulls=({nl},{nr})")
            gs = ca.compare(op, 3.0).download()   # BooleanInput::Scalar
            for i, g in enumerate(gs):
                st, want = oracle.compare(op, a[i], None, scalar=3.0)
                check_bool(g, want, f"compare-scalar op{op} {ltype} chunk {i}")
        ca.free(); cb.free()


def test_boolean_kernels_and_uploaded_masks(rdf, ctx, oracle):
    rng = np.random.default_rng(5)
    lens = [0, 5, 4097, 300001]
    A, B = [], []
    for k, n in enumerate(lens):
        pad = 3 + 2 * k
        A.append(rdf.BooleanArray.from_numpy(rng.random(n + pad) > 0.5, rng.random(n + pad) > 0.1).slice(pad, n))
        B.append(rdf.BooleanArray.from_numpy(rng.random(n + pad + 1) > 0.3).slice(1, n))
    ca, cb = rdf.Column.upload(A), rdf.Column.upload(B)
    for name, op, oop in (("and", rdf.native.AND, oracle.AND), ("or", rdf.native.OR, oracle.OR)):
        got = (ca.logical_and(cb) if name == "and" else ca.logical_or(cb)).download()
        for i, g in enumerate(got):
            st, want = oracle.boolean(oop, A[i], B[i])
            check_bool(g, want, f"{name} chunk {i}")
    for i, g in enumerate(ca.logical_not().download()):
        st, want = oracle.boolean(oracle.NOT, A[i])
        check_bool(g, want, f"not chunk {i}")
    for i, g in enumerate(ca.download()):   # a sliced boolean column re-aligned to bit offset 0 on the way back
        assert np.array_equal(g.value_bits(), A[i].value_bits()) and np.array_equal(g.valid_mask(), A[i].valid_mask())
    with pytest.raises(rdf.UnsupportedType):
        ca.add(cb)
    with pytest.raises(rdf.UnsupportedType):
        ca.sum()


@pytest.mark.parametrize("tname", ["I8", "I16", "I32", "I64", "U8", "F32", "F64"])
def test_filter_matches_oracle(rdf, ctx, oracle, tname):
    dtype = getattr(rdf, tname)
    rng = np.random.default_rng(40 + dtype)
    for null_frac, mask_nulls, sliced, p_true in ((0, 0, False, 0.5), (0.2, 0.1, True, 0.5), (0.2, 0.1, True, 0.02), (0.1, 0, True, 0.98)):
        vals = make_col(rdf, rng, dtype, RAGGED, null_frac, sliced)
        masks = []
        for k, n in enumerate(RAGGED):
            pad = (7 * k) % 13 if sliced else 0
            m = rdf.BooleanArray.from_numpy(rng.random(n + pad) < p_true, (rng.random(n + pad) > mask_nulls) if mask_nulls else None)
            masks.append(m.slice(pad, n))
        cv, cm = rdf.Column.upload(vals), rdf.Column.upload(masks)
        out = cv.filter(cm)
        got = out.download()
        for i, g in enumerate(got):
            st, want = oracle.filter_chunk(vals[i], masks[i])
            assert st == oracle.OK
            assert out.chunk_info(i)["len"] == want.length
            assert_same_array(g, want, what=f"filter<{tname}> chunk {i} (nulls {null_frac}, mask nulls {mask_nulls}, p {p_true})")
        if tname in ("I64", "F64"):
            nonempty = [w for w in (oracle.filter_chunk(v, m)[1] for v, m in zip(vals, masks)) if w.length]
            assert out.count() == sum(w.length - w.null_count for w in nonempty)
        for col in (cv, cm, out):
            col.free()
    # all-false and all-true masks, and the error paths
    v = [rdf.PrimitiveArray.from_numpy(np.arange(5000, dtype=rdf.NP_DTYPES[dtype]) % 100)]
    cv = rdf.Column.upload(v)
    none = cv.filter(rdf.Column.upload([rdf.BooleanArray.from_numpy(np.zeros(5000, bool))])).download()[0]
    allv = cv.filter(rdf.Column.upload([rdf.BooleanArray.from_numpy(np.ones(5000, bool))])).download()[0]
    assert none.length == 0 and allv.length == 5000 and np.array_equal(allv.value_slice(), v[0].value_slice())
    with pytest.raises(rdf.ComputeError):
        cv.filter(rdf.Column.upload([rdf.BooleanArray.from_numpy(np.ones(4999, bool))]))
    with pytest.raises(rdf.ArrowError):
        cv.filter(cv)


def test_lazy_filter_pipeline_1e8(rdf, ctx, oracle):
    """The lazy pipeline step this row stands for: f = a + b; keep rows where f > 0 (BooleanFilter::Gt with a scalar);
    1e8 rows, 10% nulls on b.  Device generated, oracle regenerated per chunk."""
    from concurrent.futures import ThreadPoolExecutor

    CH, NCH = 4_000_000, 25
    lens = [CH] * NCH
    a = rdf.Column.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=60)
    b = rdf.Column.generate(rdf.F64, lens, 0, -1e3, 1e3, col_id=61, null_mod=10)
    f = a.add(b)
    mask = f.gt(0.0)
    kept = f.filter(mask)
    ksum, kcount = kept.sum(), kept.count()
    got = kept.download()

    def ref(i):
        oa = oracle.generate(oracle.F64, 0, -1e3, 1e3, 20260924, 60, i * CH, CH)
        ob = oracle.generate(oracle.F64, 0, -1e3, 1e3, 20260924, 61, i * CH, CH, 10)
        _, of = oracle.col_binary(oracle.ADD, oracle.F64, [oa], [ob])
        _, om = oracle.compare(oracle.GT, of[0], None, scalar=0.0)
        _, ok = oracle.filter_chunk(of[0], om)
        return ok

    with ThreadPoolExecutor(16) as ex:
        refs = list(ex.map(ref, range(NCH)))
    exact, sum_abs, count = np.longdouble(0), np.longdouble(0), 0
    for i, want in enumerate(refs):
        assert_same_array(got[i], want, what=f"filter pipeline chunk {i}")
        e, sa = oracle.sum_exact(oracle.F64, [want])
        exact += e; sum_abs += sa; count += want.length - want.null_count
    assert kcount == count and all(g.null_count == 0 for g in got)   # nulls never pass `f > 0`: the mask is null there
    assert abs(np.longdouble(ksum) - exact) <= 16 * np.log2(1e8) * 2.0 ** -53 * sum_abs
    assert 0.4 < sum(g.length for g in got) / 1e8 < 0.5


@pytest.mark.parametrize("ltype", ["I8", "I16", "I32", "U8", "U16", "U32"])
def test_integer_column_against_scalar_fast_path(rdf, ctx, oracle, ltype):
    """k_compare_int: `cast(x, Float64) OP s` decided in the integer domain.  Every scalar that can trip the floor / ceil / clamp
    logic: fractions either side of zero, values at and beyond the type's range, infinities, NaN, signed zero."""
    lt = getattr(rdf, ltype)
    npdt = rdf.NP_DTYPES[lt]
    info = np.iinfo(npdt)
    rng = np.random.default_rng(lt + 77)
    scalars = [0.0, -0.0, 3.0, 2.5, -2.5, -3.0, 0.49, float(info.max), float(info.min), info.max + 0.5, info.min - 0.5, float(info.max) + 1, float(info.min) - 1,
               1e10, -1e10, np.inf, -np.inf, np.nan, 126.999999, 65534.5]
    for null_frac, sliced in ((0, False), (0.25, True)):
        a = make_col(rdf, rng, lt, RAGGED, null_frac, sliced)
        edge = np.array([info.min, info.max, 0, 1, 2, 3, info.max - 1, min(info.min + 1, info.max)], dtype=npdt)
        a[-1].values[a[-1].offset:a[-1].offset + len(edge)] = edge   # the values the scalars sit next to
        ca = rdf.Column.upload(a)
        for s in scalars:
            for op in range(6):
                got = ca.compare(op, s).download()
                for i, g in enumerate(got):
                    st, want = oracle.compare(op, a[i], None, scalar=s)
                    assert st == oracle.OK
                    check_bool(g, want, f"{ltype} op{op} scalar {s!r} chunk {i} nulls={null_frac}")
        ca.free()
