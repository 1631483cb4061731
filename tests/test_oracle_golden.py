"""Pins the CPU oracle (oracle/oracle.c) before anything is compared against it.

1. Every golden vector / known-answer the REFERENCE's own tests hold for the hot path
   (SURVEY.md 8(c)): abs i32/f64, acos/cos goldens, count, avg, add on the CSV fixture, the
   par_multiply bench input.
2. Cross-checks against independent implementations on the subset where they agree with arrow-rs:
   pyarrow (Arrow C++) for wrapping add/sub/mul, integer divide, in-range casts, min/max/count;
   numpy for IEEE float arithmetic.
3. The semantics the reference inherits from arrow-rs / num-traits that NO reference test pins
   ("parity unpinned" in oracle.h): DivideByZero rule, cast -> NULL rule, sum order, sliced arrays.
   These are checked against the published rules they restate.
"""
import json
import math
import os

import numpy as np
import pytest

from helpers import ulp_distance

HERE = os.path.dirname(os.path.abspath(__file__))


class Chunk:
    """Minimal duck-typed chunk for the oracle wrapper."""

    def __init__(self, values, dtype, mask=None, offset=0, length=None):
        self.values = np.ascontiguousarray(values)
        self.dtype = dtype
        self.offset = offset
        self.length = len(self.values) - offset if length is None else length
        if mask is None:
            self.validity, self.null_count = None, 0
        else:
            self.validity = np.packbits(np.asarray(mask, dtype=bool), bitorder="little")
            self.null_count = -1


def from_list(orc, dtype, items):
    mask = [x is not None for x in items]
    vals = np.array([0 if x is None else x for x in items], dtype=orc.NP_DTYPES[dtype])
    return Chunk(vals, dtype, None if all(mask) else mask)


def to_list(arr):
    m = arr.valid_mask()
    return [arr.values[i].item() if m[i] else None for i in range(arr.length)]


# ---- 1. the reference's own goldens -------------------------------------------------------------------

def test_abs_i32_reference_golden(oracle):  # src/functions/scalar.rs:576-584
    st, out = oracle.col_unary(oracle.ABS, oracle.I32, [from_list(oracle, oracle.I32, [-5, -6, 7, -8, -0])])
    assert st == oracle.OK and to_list(out[0]) == [5, 6, 7, 8, 0]


def test_abs_f64_reference_golden(oracle):  # src/functions/scalar.rs:565-573
    st, out = oracle.col_unary(oracle.ABS, oracle.F64, [from_list(oracle, oracle.F64, [-5.2, -6.1, 7.3, -8.6, -0.0])])
    assert st == oracle.OK
    assert to_list(out[0]) == [5.2, 6.1, 7.3, 8.6, 0.0]
    assert not math.copysign(1.0, out[0].values[4]) < 0  # |-0.0| = +0.0


def test_acos_cos_reference_goldens(oracle):  # src/functions/scalar.rs:587-602
    x = from_list(oracle, oracle.F64, [-0.2, 0.25, 0.75])
    _, acos = oracle.col_unary(oracle.ACOS, oracle.F64, [x])
    _, cos = oracle.col_unary(oracle.COS, oracle.F64, [x])
    assert to_list(acos[0]) == [1.7721542475852274, 1.318116071652818, 0.7227342478134157]
    assert to_list(cos[0]) == [0.9800665778412416, 0.9689124217106447, 0.7316888688738209]


def test_count_reference_golden(oracle):  # src/functions/aggregate.rs:123-127
    st, v = oracle.aggregate(oracle.COUNT, oracle.I32, [from_list(oracle, oracle.I32, [5, 6, 7, 8, 9])])
    assert st == oracle.OK and v == 5


def test_avg_reference_goldens(oracle):  # src/functions/aggregate.rs:130-146
    a = from_list(oracle, oracle.I32, [0, 1, 2, 3, 4])
    b = from_list(oracle, oracle.I32, [5, 6, 7, 8, 9])
    assert oracle.avg(oracle.I32, [a, b]) == (oracle.OK, 4.5)
    d = from_list(oracle, oracle.I32, [0, None, 1, None, 2, 3, 4])
    assert oracle.avg(oracle.I32, [d, b]) == (oracle.OK, 4.5)


def test_par_multiply_bench_input(oracle):  # src/functions/scalar.rs:621-671: [None,200,None,-256,None]^2
    a = from_list(oracle, oracle.I32, [None, 200, None, -256, None])
    st, out = oracle.col_binary(oracle.MUL, oracle.I32, [a] * 380, [a] * 380, threads=4)
    assert st == oracle.OK and len(out) == 380
    for o in out:
        assert to_list(o) == [None, 40000, None, 65536, None] and o.null_count == 3


@pytest.fixture(scope="module")
def cities():
    with open(os.path.join(HERE, "golden", "uk_cities.json")) as f:
        return json.load(f)


def test_csv_fixture_reference_asserts(oracle, cities):
    """add / abs on the CSV fixture: the reference asserts row 0 (src/dataframe.rs:803-808, 835)."""
    ref = cities["reference_asserts"]
    lat = Chunk(np.array(cities["lat"]), oracle.F64)
    lng = Chunk(np.array(cities["lng"]), oracle.F64)
    assert lat.length == ref["n_rows"] and lat.values[0] == ref["lat_row0"] and lng.values[0] == ref["lng_row0"]
    _, s = oracle.col_binary(oracle.ADD, oracle.F64, [lat], [lng])
    assert abs(ref["lat_plus_lng_row0"] - s[0].values[0]) < 1e-4
    _, a = oracle.col_unary(oracle.ABS, oracle.F64, [lng])
    assert abs(ref["abs_lng_row0"] - a[0].values[0]) < np.finfo(np.float64).eps
    # derived values committed with the fixture (oracle regression pins; SURVEY 8(c) lists sum(lat))
    d = cities["derived"]
    assert [float(x).hex() for x in s[0].values] == d["lat_plus_lng_hex"]
    assert float(oracle.aggregate(oracle.SUM, oracle.F64, [lat])[1]).hex() == d["sum_lat_hex"]
    assert float.fromhex(d["sum_lat_hex"]) == 1948.1160980000002


# ---- 2. independent cross-checks ------------------------------------------------------------------------

INT_TYPES = ["I8", "I16", "I32", "I64", "U8", "U16", "U32", "U64"]


@pytest.mark.parametrize("tname", INT_TYPES)
def test_int_arithmetic_matches_pyarrow(oracle, tname):
    import pyarrow as pa
    import pyarrow.compute as pc

    dtype = getattr(oracle, tname)
    npdt = oracle.NP_DTYPES[dtype]
    rng = np.random.default_rng(dtype)
    info = np.iinfo(npdt)
    n = 3000
    x = rng.integers(info.min, info.max, n, dtype=npdt, endpoint=True)
    y = rng.integers(info.min, info.max, n, dtype=npdt, endpoint=True)
    y[y == 0] = 1
    if info.min < 0:
        y[y == -1] = 2  # MIN / -1 is unspecified in the reference
    ma, mb = rng.random(n) > 0.2, rng.random(n) > 0.2
    a, b = Chunk(x, dtype, ma), Chunk(y, dtype, mb)
    pa_a, pa_b = pa.array(x, mask=~ma), pa.array(y, mask=~mb)
    for op, fn in ((oracle.ADD, pc.add), (oracle.SUB, pc.subtract), (oracle.MUL, pc.multiply), (oracle.DIV, pc.divide)):
        st, out = oracle.col_binary(op, dtype, [a], [b])
        assert st == oracle.OK
        want = fn(pa_a, pa_b)  # unchecked variants wrap like arrow-rs' SIMD kernels
        m = out[0].valid_mask()
        assert np.array_equal(m, ma & mb)
        assert np.array_equal(m, ~np.asarray(want.is_null()))
        assert np.array_equal(out[0].values[m], want.drop_null().to_numpy())
        assert out[0].null_count == want.null_count
    # aggregates
    for op, want in ((oracle.MIN, pc.min(pa_a)), (oracle.MAX, pc.max(pa_a)), (oracle.COUNT, pc.count(pa_a))):
        st, v = oracle.aggregate(op, dtype, [a])
        assert st == oracle.OK and int(v) == want.as_py()
    st, v = oracle.aggregate(oracle.SUM, dtype, [a, b])
    wide = int(x[ma].astype(object).sum()) + int(y[mb].astype(object).sum())
    bits = 8 * np.dtype(npdt).itemsize
    wrapped = wide % (1 << bits)
    if info.min < 0 and wrapped >= 1 << (bits - 1):
        wrapped -= 1 << bits
    assert int(v) == wrapped  # wrapping sum == exact sum mod 2^bits


@pytest.mark.parametrize("tname", ["F32", "F64"])
def test_float_arithmetic_matches_numpy(oracle, tname):
    dtype = getattr(oracle, tname)
    npdt = oracle.NP_DTYPES[dtype]
    rng = np.random.default_rng(11)
    x = rng.uniform(-1e3, 1e3, 5000).astype(npdt)
    y = (rng.uniform(1, 2, 5000) * rng.choice([-1, 1], 5000)).astype(npdt)
    x[:4] = [0.0, -0.0, np.inf, np.nan]
    for op, fn in ((oracle.ADD, np.add), (oracle.SUB, np.subtract), (oracle.MUL, np.multiply), (oracle.DIV, np.divide)):
        st, out = oracle.col_binary(op, dtype, [Chunk(x, dtype)], [Chunk(y, dtype)])
        assert st == oracle.OK and out[0].validity is None
        with np.errstate(all="ignore"):
            want = fn(x, y)
        assert np.array_equal(out[0].values.view(f"u{npdt().itemsize}"), want.view(f"u{npdt().itemsize}"))


def test_in_range_casts_match_pyarrow(oracle):
    import pyarrow as pa
    import pyarrow.compute as pc

    rng = np.random.default_rng(3)
    names = INT_TYPES + ["F32", "F64"]
    for fname in names:
        f = getattr(oracle, fname)
        fdt = oracle.NP_DTYPES[f]
        for tname in names:
            t = getattr(oracle, tname)
            tdt = oracle.NP_DTYPES[t]
            src = rng.integers(0, 100, 257).astype(fdt)  # representable everywhere
            mask = rng.random(257) > 0.1
            st, out = oracle.col_cast(f, t, [Chunk(src, f, mask)])
            assert st == oracle.OK
            want = pc.cast(pa.array(src, mask=~mask), pa.from_numpy_dtype(tdt))
            assert np.array_equal(out[0].valid_mask(), mask)
            assert np.array_equal(out[0].values[mask], want.drop_null().to_numpy())


# ---- 3. inherited semantics no reference test pins -------------------------------------------------------

def test_divide_by_zero_rule(oracle):
    """arrow-rs divide: Err(DivideByZero) iff a VALID slot has a zero divisor -- ints and floats."""
    for dtype, zero in ((oracle.I32, 0), (oracle.F64, 0.0), (oracle.F64, -0.0), (oracle.U8, 0)):
        a = from_list(oracle, dtype, [6, 8, 10])
        assert oracle.col_binary(oracle.DIV, dtype, [a], [from_list(oracle, dtype, [2, zero, 5])])[0] == oracle.DIVIDE_BY_ZERO
        # the zero sits under a null of the divisor: ignored
        b = Chunk(np.array([2, zero, 5], dtype=oracle.NP_DTYPES[dtype]), dtype, [True, False, True])
        st, out = oracle.col_binary(oracle.DIV, dtype, [a], [b])
        assert st == oracle.OK and to_list(out[0]) == [3, None, 2]
        # ... or under a null of the dividend
        a2 = Chunk(np.array([6, 8, 10], dtype=oracle.NP_DTYPES[dtype]), dtype, [True, False, True])
        st, out = oracle.col_binary(oracle.DIV, dtype, [a2], [from_list(oracle, dtype, [2, zero, 5])])
        assert st == oracle.OK and to_list(out[0]) == [3, None, 2]
    # first error wins across chunks, later chunks irrelevant
    ok = from_list(oracle, oracle.I32, [1, 2])
    bad = from_list(oracle, oracle.I32, [1, 0])
    assert oracle.col_binary(oracle.DIV, oracle.I32, [ok, ok], [ok, bad])[0] == oracle.DIVIDE_BY_ZERO


def test_int_divide_truncates_and_wraps(oracle):
    a = from_list(oracle, oracle.I32, [7, -7, 7, -7, -2 ** 31])
    b = from_list(oracle, oracle.I32, [2, 2, -2, -2, -1])
    st, out = oracle.col_binary(oracle.DIV, oracle.I32, [a], [b])
    assert st == oracle.OK and to_list(out[0]) == [3, -3, -3, 3, -2 ** 31]


def test_length_mismatch_and_zip(oracle):
    a = from_list(oracle, oracle.I64, [1, 2, 3])
    b = from_list(oracle, oracle.I64, [1, 2])
    assert oracle.col_binary(oracle.ADD, oracle.I64, [a], [b])[0] == oracle.LENGTH_MISMATCH
    st, out = oracle.col_binary(oracle.ADD, oracle.I64, [a, a, a], [a])  # zip truncates to the shorter Vec
    assert st == oracle.OK and len(out) == 1


def test_validity_presence_and_sliced_inputs(oracle):
    x = np.arange(100, dtype=np.int64)
    mask = (np.arange(100) % 3) != 0
    full = Chunk(x, oracle.I64, mask)
    sl = Chunk(x, oracle.I64, mask, offset=13, length=50)   # ChunkedArray::slice (src/table.rs:77-95)
    plain = Chunk(x[:50].copy(), oracle.I64)
    st, out = oracle.col_binary(oracle.ADD, oracle.I64, [sl], [plain])
    assert st == oracle.OK
    assert np.array_equal(out[0].valid_mask(), mask[13:63])
    assert np.array_equal(out[0].values, x[13:63] + x[:50])      # computed under nulls too
    assert out[0].null_count == int((~mask[13:63]).sum())
    st, out = oracle.col_binary(oracle.ADD, oracle.I64, [plain], [plain])
    assert out[0].validity is None                                 # no input bitmap -> no output bitmap
    assert oracle.lib().orc_null_count(oracle._views([full])) == int((~mask).sum())


def test_cast_none_becomes_null(oracle):
    """num::cast::cast returns None -> NULL (SURVEY 8(d) config 4 parity cases)."""
    f = from_list(oracle, oracle.F64, [float("nan"), float("inf"), float("-inf"), 3e10, -3e10, -1.9, 1.9, 2147483647.9,
                                       -2147483648.9, 2147483648.0, -2147483649.0])
    st, out = oracle.col_cast(oracle.F64, oracle.I32, [f])
    assert to_list(out[0]) == [None, None, None, None, None, -1, 1, 2147483647, -2147483648, None, None]
    st, out = oracle.col_cast(oracle.I64, oracle.I32, [from_list(oracle, oracle.I64, [2 ** 31, -2 ** 31 - 1, 2 ** 31 - 1, -2 ** 31, None])])
    assert to_list(out[0]) == [None, None, 2 ** 31 - 1, -2 ** 31, None]
    st, out = oracle.col_cast(oracle.I32, oracle.U32, [from_list(oracle, oracle.I32, [-1, 0, 5])])
    assert to_list(out[0]) == [None, 0, 5]
    st, out = oracle.col_cast(oracle.F64, oracle.U8, [from_list(oracle, oracle.F64, [-0.5, -1.0, 255.9, 256.0])])
    assert to_list(out[0]) == [0, None, 255, None]
    st, out = oracle.col_cast(oracle.F32, oracle.I64, [from_list(oracle, oracle.F32, [9.223372e18, 9.2233715e18, -9.223372e18])])
    assert to_list(out[0]) == [None, 9223371487098961920, -9223372036854775808]
    st, out = oracle.col_cast(oracle.U64, oracle.I64, [from_list(oracle, oracle.U64, [2 ** 63, 2 ** 63 - 1])])
    assert to_list(out[0]) == [None, 2 ** 63 - 1]
    # int -> float always Some, round-to-nearest-even; f64 -> f32 overflow saturates to inf (Some)
    st, out = oracle.col_cast(oracle.I64, oracle.F64, [from_list(oracle, oracle.I64, [2 ** 53 + 1, 2 ** 53 + 3])])
    assert to_list(out[0]) == [float(2 ** 53), float(2 ** 53 + 4)]
    st, out = oracle.col_cast(oracle.F64, oracle.F32, [from_list(oracle, oracle.F64, [1e300, -1e300, 0.1])])
    assert out[0].null_count == 0 and to_list(out[0])[:2] == [float("inf"), float("-inf")]
    # same-type cast is a clone
    st, out = oracle.col_cast(oracle.I16, oracle.I16, [from_list(oracle, oracle.I16, [1, None, 3])])
    assert to_list(out[0]) == [1, None, 3]


def test_sum_min_max_option_rules(oracle):
    allnull = Chunk(np.array([1, 2], dtype=np.int64), oracle.I64, [False, False])
    some = from_list(oracle, oracle.I64, [5, None, -7])
    assert oracle.aggregate(oracle.SUM, oracle.I64, [allnull, some]) == (oracle.OK, -2)
    assert oracle.aggregate(oracle.SUM, oracle.I64, []) == (oracle.OK, 0)             # Some(default)
    assert oracle.aggregate(oracle.MAX, oracle.I64, []) == (oracle.OK, None)          # empty Vec -> None
    assert oracle.aggregate(oracle.MAX, oracle.I64, [some, allnull])[0] == oracle.PANIC  # unwrap() on None
    assert oracle.aggregate(oracle.MAX, oracle.I64, [some]) == (oracle.OK, 5)
    assert oracle.aggregate(oracle.MIN, oracle.I64, [some]) == (oracle.OK, -7)
    assert oracle.aggregate(oracle.MIN_AS_WRITTEN, oracle.I64, [some]) == (oracle.OK, 5)  # the reference's bug
    assert oracle.aggregate(oracle.MAX, oracle.F64, [from_list(oracle, oracle.F64, [1.0])])[0] == oracle.UNSUPPORTED
    assert oracle.aggregate(oracle.COUNT, oracle.I64, [allnull, some]) == (oracle.OK, 2)


def test_float_sum_is_sequential_fold(oracle):
    rng = np.random.default_rng(5)
    x = rng.uniform(-1e3, 1e3, 10001)
    mask = rng.random(10001) > 0.1
    acc = 0.0
    for v, m in zip(x[:6000], mask[:6000]):
        if m:
            acc += v
    acc2 = 0.0
    for v, m in zip(x[6000:], mask[6000:]):
        if m:
            acc2 += v
    st, s = oracle.aggregate(oracle.SUM, oracle.F64, [Chunk(x[:6000], oracle.F64, mask[:6000]), Chunk(x[6000:], oracle.F64, mask[6000:])])
    assert st == oracle.OK and s == (0.0 + acc) + acc2
    exact, sum_abs = oracle.sum_exact(oracle.F64, [Chunk(x, oracle.F64, mask)])
    assert abs(np.longdouble(math.fsum(x[mask])) - exact) <= 1e-9
    assert abs(np.longdouble(s) - exact) <= len(x) * 2.0 ** -53 * sum_abs


def test_trig_is_glibc(oracle):
    x = np.array([0.0, -0.0, 1e-300, 0.5, -2.5, 1e6, 1e22, np.inf, -np.inf, np.nan])
    for op, fn in ((oracle.SIN, math.sin), (oracle.COS, math.cos), (oracle.TAN, math.tan)):
        _, out = oracle.col_unary(op, oracle.F64, [Chunk(x, oracle.F64)])
        for got, v in zip(out[0].values, x):
            want = fn(v) if np.isfinite(v) else float("nan")   # Rust: inf.sin() = NaN, no error
            assert (np.isnan(got) and np.isnan(want)) or got == want
    _, out = oracle.col_unary(oracle.SIN, oracle.F64, [Chunk(x, oracle.F64, [True] * 9 + [False])])
    assert out[0].null_count == 1 and out[0].values[9] == 0.0       # null payload 0
    assert math.copysign(1, out[0].values[1]) < 0                   # sin(-0.0) = -0.0
    assert oracle.col_unary(oracle.SIN, oracle.I32, [from_list(oracle, oracle.I32, [1])])[0] == oracle.UNSUPPORTED


def test_generator_is_counter_based(oracle):
    a = oracle.generate(oracle.F64, 0, -1e3, 1e3, 20260924, 1, 0, 1000, null_mod=10)
    b = oracle.generate(oracle.F64, 0, -1e3, 1e3, 20260924, 1, 500, 500, null_mod=10)
    assert np.array_equal(a.values[500:], b.values) and np.array_equal(a.valid_mask()[500:], b.valid_mask())
    assert a.values.min() >= -1e3 and a.values.max() < 1e3 and 50 < a.null_count < 150
    d = oracle.generate(oracle.F64, 1, 0, 0, 1, 2, 0, 1000)
    assert np.all((np.abs(d.values) >= 1) & (np.abs(d.values) < 2)) and (d.values < 0).any() and (d.values > 0).any()
    i = oracle.generate(oracle.I64, 3, 0, 0, 1, 3, 0, 1000)
    assert i.values.min() >= -2 ** 40 and i.values.max() < 2 ** 40
    h = oracle.lib().orc_splitmix64(0)
    assert h == 0xE220A8397B1DCDAF  # published splitmix64 test vector (first output for seed 0)


# ---- 4. N2 (next row): BooleanFilter comparisons, boolean kernels, filter -- cross-checked with pyarrow -----------

def test_compare_bool_filter_match_pyarrow(oracle):
    import pyarrow as pa
    import pyarrow.compute as pc

    rng = np.random.default_rng(21)
    n = 3001
    x = rng.integers(-50, 50, n).astype(np.int32)
    y = rng.uniform(-50, 50, n)
    y[:5] = [np.nan, np.inf, -np.inf, 0.0, -0.0]
    x[3:5] = 0
    mx, my = rng.random(n) > 0.2, rng.random(n) > 0.1
    cx, cy = Chunk(x, oracle.I32, mx), Chunk(y, oracle.F64, my)
    pax, pay = pa.array(x, mask=~mx).cast(pa.float64()), pa.array(y, mask=~my)
    ops = [(oracle.GT, pc.greater), (oracle.GE, pc.greater_equal), (oracle.EQ, pc.equal), (oracle.NE, pc.not_equal),
           (oracle.LT, pc.less), (oracle.LE, pc.less_equal)]
    masks = []
    for op, fn in ops:
        st, m = oracle.compare(op, cx, cy)
        assert st == oracle.OK
        want = fn(pax, pay)
        assert np.array_equal(m.valid_mask(), mx & my) and m.null_count == want.null_count
        assert np.array_equal(m.value_bits()[m.valid_mask()], want.drop_null().to_numpy(zero_copy_only=False))
        st, ms = oracle.compare(op, cx, None, scalar=3.0)   # BooleanInput::Scalar broadcast
        wants = fn(pax, pa.scalar(3.0))
        assert np.array_equal(ms.value_bits()[ms.valid_mask()], wants.drop_null().to_numpy(zero_copy_only=False))
        masks.append(m)
    # boolean kernels: values op values, validity AND (arrow-rs compute::and / or / not)
    a, b = masks[0], masks[3]
    for op, npop in ((oracle.AND, np.logical_and), (oracle.OR, np.logical_or)):
        st, r = oracle.boolean(op, a, b)
        assert np.array_equal(r.valid_mask(), a.valid_mask() & b.valid_mask())
        v = r.valid_mask()
        assert np.array_equal(r.value_bits()[v], npop(a.value_bits(), b.value_bits())[v])
    st, r = oracle.boolean(oracle.NOT, a)
    assert np.array_equal(r.value_bits()[r.valid_mask()], ~a.value_bits()[a.valid_mask()])
    # filter: null mask slots count as false; kept slots keep their validity
    for values, pav in ((cx, pa.array(x, mask=~mx)), (cy, pay)):
        st, f = oracle.filter_chunk(values, a)
        pam = pa.array(a.value_bits(), mask=~a.valid_mask())
        want = pc.filter(pav, pam, null_selection_behavior="drop")
        assert f.length == len(want) and f.null_count == want.null_count
        assert np.array_equal(f.valid_mask(), ~np.asarray(want.is_null()))
        wv = want.to_numpy(zero_copy_only=False)
        got = f.values[f.valid_mask()]
        assert np.array_equal(got, wv[~np.asarray(want.is_null())].astype(got.dtype), equal_nan=True)
    st, fb = oracle.filter_chunk(b, a)   # filtering a boolean column
    sel = a.value_bits() & a.valid_mask()
    assert fb.length == int(sel.sum()) and np.array_equal(fb.value_bits(), b.value_bits()[sel])
    assert oracle.filter_chunk(cx, Chunk(np.zeros(1, np.uint8), oracle.BOOL, None, 0, 5))[0] == oracle.LENGTH_MISMATCH


# ---- DataFrame::sort: lexsort_to_indices + take ---------------------------------------------------------

def test_sort_reference_golden(oracle):  # src/dataframe.rs:963-1002 (test_sort): a descending, b ascending, nulls last
    a = from_list(oracle, oracle.I32, [1, 1, None, 3, 3, 4])
    b = from_list(oracle, oracle.U8, [9, 5, 6, 7, 4, 8])
    st, idx = oracle.lexsort_indices([([a], True), ([b], False)])
    assert st == oracle.OK
    _, sa = oracle.take([a], idx)
    _, sb = oracle.take([b], idx)
    assert to_list(sa) == [4, 3, 3, 1, 1, None]          # is_null(5), values 4,3,3,1,1
    assert to_list(sb) == [8, 4, 7, 5, 9, 6]
    assert oracle.lexsort_indices([])[0] != oracle.OK     # "Sort criteria cannot be empty"


def test_sort_and_take_match_pyarrow(oracle):
    """Stable multi-key order (nulls last whatever the direction) and take, against Arrow C++ on NaN-free data
    (Arrow C++ keeps NaN behind the numbers in descending order too; arrow-rs reverses the comparator, NaN first)."""
    import pyarrow as pa
    import pyarrow.compute as pc

    rng = np.random.default_rng(33)
    n = 5000
    k1 = rng.integers(-3, 4, n).astype(np.int16)
    k2 = np.round(rng.normal(0, 2, n)).astype(np.float64)
    k2[::17] = -0.0
    k3 = rng.integers(0, 2 ** 63, n).astype(np.uint64)
    m1, m2 = rng.random(n) > 0.2, rng.random(n) > 0.1
    split = [0, 1, 1, 1200, 5000]
    def chunks(v, dt, m=None):
        return [Chunk(v, dt, m, offset=s, length=e - s) if m is not None else Chunk(v[s:e], dt) for s, e in zip(split[:-1], split[1:])]
    c1 = [Chunk(k1, oracle.I16, m1, offset=s, length=e - s) for s, e in zip(split[:-1], split[1:])]
    c2 = [Chunk(k2, oracle.F64, m2, offset=s, length=e - s) for s, e in zip(split[:-1], split[1:])]
    c3 = [Chunk(k3[s:e], oracle.U64) for s, e in zip(split[:-1], split[1:])]
    t = pa.table({"k1": pa.array(k1, mask=~m1), "k2": pa.array(k2, mask=~m2), "k3": pa.array(k3)})
    for desc in ((False, False, False), (True, False, True), (False, True, False), (True, True, True)):
        st, idx = oracle.lexsort_indices([(c1, desc[0]), (c2, desc[1]), (c3, desc[2])])
        assert st == oracle.OK
        want = pc.sort_indices(t, sort_keys=[(f"k{i + 1}", "descending" if d else "ascending") for i, d in enumerate(desc)], null_placement="at_end")
        assert np.array_equal(idx, want.to_numpy().astype(np.uint32)), desc
    # ties only: a stable sort leaves the rows where they are
    st, idx = oracle.lexsort_indices([([Chunk(np.zeros(100, np.int8), oracle.I8)], True)])
    assert np.array_equal(idx, np.arange(100))
    # NaN is the greatest number (cmp_nans_last), -0.0 == 0.0, nulls after NaN
    f = Chunk(np.array([np.nan, 1.0, -np.inf, 0.0, -0.0, np.inf, 7.0, np.nan]), oracle.F64, [1, 1, 1, 1, 1, 1, 0, 1])
    assert oracle.lexsort_indices([([f], False)])[1].tolist() == [2, 3, 4, 1, 5, 0, 7, 6]
    assert oracle.lexsort_indices([([f], True)])[1].tolist() == [0, 7, 5, 1, 3, 4, 2, 6]
    # take with null indices and across chunks
    ix = rng.integers(0, n, 777).astype(np.uint32)
    iv = rng.random(777) > 0.3
    st, got = oracle.take(c2, ix, iv)
    want = pc.take(t["k2"], pa.array(ix, mask=~iv))
    assert to_list(got) == want.to_pylist()
    assert oracle.take(c2, np.array([n], np.uint32))[0] == oracle.PANIC
