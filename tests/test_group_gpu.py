"""Group-by aggregate (bdf_group_aggregate_dev): the operator the reference leaves as a panic! (src/evaluation.rs:73).  Checked
against the numpy restatement in oracle/pyoracle.py (itself cross-checked against pyarrow's hash aggregate below)."""
import numpy as np
import pytest


def _chunks(rdf, values, mask, lens):
    out, row = [], 0
    for n in lens:
        out.append(rdf.PrimitiveArray.from_numpy(values[row:row + n], None if mask is None else mask[row:row + n]))
        row += n
    return out


def _col_values(col):
    """One-chunk result column -> (values, valid mask, array)."""
    arrs = col.download()
    assert len(arrs) == 1
    return arrs[0].value_slice(), arrs[0].valid_mask(), arrs[0]


def test_oracle_group_aggregate_matches_pyarrow(oracle):
    import pyarrow as pa

    rng = np.random.default_rng(3)
    n = 20_000
    k = rng.integers(-50, 50, n).astype(np.int32)
    km = rng.random(n) > 0.05
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    vm = rng.random(n) > 0.2

    class Ch:
        def __init__(self, values, mask):
            self.values, self.offset, self.length, self._m = values, 0, len(values), mask

        def valid_mask(self):
            return self._m

    keys, kvalid, out = oracle.group_aggregate([Ch(k, km)], [Ch(v, vm)])
    t = pa.table({"k": pa.array(k, mask=~km), "v": pa.array(v, mask=~vm)})
    g = t.group_by("k").aggregate([("v", "sum"), ("v", "count"), ("v", "min"), ("v", "max")]).to_pydict()
    want = {kk: (s, c, mn, mx) for kk, s, c, mn, mx in zip(g["k"], g["v_sum"], g["v_count"], g["v_min"], g["v_max"])}
    assert len(want) == len(keys)
    for i in range(len(keys)):
        kk = int(keys[i]) if kvalid[i] else None
        s, c, mn, mx = want[kk]
        assert out["count"][i] == c and int(out["sum"][i]) == (s or 0)
        assert (bool(out["min"][1][i]), bool(out["max"][1][i])) == (mn is not None, mx is not None)
        if mn is not None:
            assert int(out["min"][0][i]) == mn and int(out["max"][0][i]) == mx
    assert list(keys[kvalid]) == sorted(keys[kvalid]) and (not kvalid.all()) and not kvalid[-1]   # ascending, the null key last
    # narrow values: sums wrap in T's width (AggregateFunctions::sum adds T::Native), whatever numpy's default accumulator is
    v16 = rng.integers(-32768, 32767, n).astype(np.int16)
    _, _, out16 = oracle.group_aggregate([Ch(k, km)], [Ch(v16, np.ones(n, bool))])
    assert out16["sum"].dtype == np.int16 and len(out16["sum"]) == len(keys)
    order = np.lexsort((np.where(km, k, 0), ~km))
    first = np.nonzero(np.where(km, k, 0)[order] == np.where(km, k, 0)[order][0])[0]
    assert int(out16["sum"][0]) == int(np.sum(v16[order][first[first < np.searchsorted(~km[order], True)]].astype(np.int64)).astype(np.int16))


CASES = [("int32", "int64", 200, 0.05, 0.1), ("float64", "float64", 50, 0.0, 0.2), ("int8", "int16", 300, 0.1, 0.0),
         ("uint64", "uint32", 7, 0.0, 0.5), ("float32", "int8", 1000, 0.02, 0.05), ("int64", "float32", 3, 0.3, 0.0),
         # short groups: 8 lanes per group (average 4..64 rows), one lane per group (average <= 4 rows)
         ("int32", "int32", 20_000, 0.05, 0.3), ("int64", "int64", 400_000, 0.0, 0.4), ("uint64", "float64", 100_000, 0.1, 0.1),
         ("int16", "uint8", 30_000, 0.0, 0.6)]


@pytest.mark.gpu
@pytest.mark.parametrize("kname,vname,cardinality,knull,vnull", CASES)
def test_group_aggregate_matches_the_oracle(rdf, ctx, oracle, kname, vname, cardinality, knull, vnull):
    rng = np.random.default_rng(sum(map(ord, kname + vname)))
    lens = [50_000, 1, 0, 123_457, 64]
    n = sum(lens)
    kd, vd = np.dtype(kname), np.dtype(vname)
    if kd.kind == "f":
        k = rng.integers(-cardinality // 2, cardinality // 2, n).astype(kd) * kd.type(0.5)
        k[rng.random(n) < 0.01] = np.nan            # NaN keys: one group after every number
        k[rng.random(n) < 0.01] = -0.0              # -0.0 and 0.0 are one group
    elif kd.kind == "u":
        k = (rng.integers(0, cardinality, n).astype(np.uint64) * np.uint64(2 ** 60 // max(cardinality, 1))).astype(kd)
    else:
        info = np.iinfo(kd)
        k = rng.integers(max(info.min, -cardinality // 2), min(info.max, cardinality // 2), n, endpoint=True).astype(kd)
    if vd.kind == "f":
        v = rng.uniform(-1e3, 1e3, n).astype(vd)
    else:
        info = np.iinfo(vd)
        v = rng.integers(info.min, info.max, n, dtype=vd, endpoint=True)   # full range: sums wrap
    km = rng.random(n) >= knull if knull else None
    vm = rng.random(n) >= vnull if vnull else None
    kch, vch = _chunks(rdf, k, km, lens), _chunks(rdf, v, vm, lens)
    ck, cv = rdf.Column.upload(kch, ctx=ctx), rdf.Column.upload(vch, ctx=ctx)
    keys, res = rdf.group_aggregate(ck, [cv])
    okeys, okvalid, want = oracle.group_aggregate(kch, vch)
    gk, gkv, _ = _col_values(keys)
    assert len(gk) == len(okeys) and np.array_equal(gkv, okvalid)
    if kd.kind == "f":
        both_nan = np.isnan(gk) & np.isnan(okeys)
        assert np.array_equal(np.where(both_nan | ~gkv, 0, gk) + 0.0, np.where(both_nan | ~okvalid, 0, okeys) + 0.0)
    else:
        assert np.array_equal(gk[gkv], okeys[okvalid])
    r = res[0]
    gcount, _, _ = _col_values(r["count"])
    assert np.array_equal(gcount, want["count"])
    gsum, gsv, sum_arr = _col_values(r["sum"])
    assert gsv.all() and sum_arr.validity is None          # an all-null group sums to 0, never NULL
    if vd.kind == "f":
        exact, mag = want["exact"]
        eps = 2.0 ** -53 if vd.itemsize == 8 else 2.0 ** -24
        ng = np.maximum(want["count"], 2)
        assert np.all(np.abs(gsum.astype(np.longdouble) - exact) <= 16 * np.log2(ng) * eps * mag + 1e-300)
        assert r["min"] is None and r["max"] is None       # T::Native: Ord
    else:
        assert np.array_equal(gsum, want["sum"])
        for key in ("min", "max"):
            g, gv, arr = _col_values(r[key])
            wv, wm = want[key]
            assert np.array_equal(gv, wm) and np.array_equal(g[gv], wv[wm])
            assert arr.null_count == int((~wm).sum())
    for col in (ck, cv, keys, r["sum"], r["count"], r["min"], r["max"]):
        if col is not None:
            col.free()


@pytest.mark.gpu
def test_group_aggregate_shapes(rdf, ctx, oracle):
    """Edge shapes: empty input, one group, every row its own group, several value columns at once, determinism."""
    empty = rdf.Column.upload([rdf.PrimitiveArray.from_numpy(np.zeros(0, np.int64))], ctx=ctx)
    keys, res = rdf.group_aggregate(empty, [empty])
    assert len(keys) == 0 and len(res[0]["sum"]) == 0
    n = 300_000
    rng = np.random.default_rng(1)
    one = rdf.Column.upload([rdf.PrimitiveArray.from_numpy(np.full(n, 7, np.int16))], ctx=ctx)
    vals = rng.integers(-10 ** 6, 10 ** 6, n)
    v = rdf.Column.upload([rdf.PrimitiveArray.from_numpy(vals)], ctx=ctx)
    f = rdf.Column.upload([rdf.PrimitiveArray.from_numpy(rng.uniform(-1, 1, n))], ctx=ctx)
    keys, res = rdf.group_aggregate(one, [v, f])
    assert len(keys) == 1 and int(res[0]["sum"].download()[0].value_slice()[0]) == int(vals.sum()) and int(res[0]["count"].download()[0].value_slice()[0]) == n
    first = res[1]["sum"].download()[0].value_slice()[0]
    keys2, res2 = rdf.group_aggregate(one, [v, f])
    assert np.float64(res2[1]["sum"].download()[0].value_slice()[0]).view(np.uint64) == np.float64(first).view(np.uint64)   # deterministic
    uniq = rdf.Column.upload([rdf.PrimitiveArray.from_numpy(rng.permutation(n).astype(np.int64))], ctx=ctx)
    keys3, res3 = rdf.group_aggregate(uniq, [v])
    assert len(keys3) == n and np.array_equal(keys3.download()[0].value_slice(), np.arange(n))
    perm_keys = uniq.download()[0].value_slice()
    assert np.array_equal(res3[0]["sum"].download()[0].value_slice()[perm_keys], vals)   # group of key k holds the row whose key is k
    with pytest.raises(rdf.ComputeError):
        rdf.group_aggregate(one, [empty])   # different lengths


@pytest.mark.gpu
def test_device_frame_group_aggregate(rdf, ctx):
    """DeviceFrame.group_aggregate against pandas' groupby on the same columns (and the evaluator still mirrors the reference's panic)."""
    import pandas as pd

    rng = np.random.default_rng(9)
    n = 100_000
    k = rng.integers(0, 40, n).astype(np.int64)
    x = rng.integers(-1000, 1000, n).astype(np.int32)
    y = rng.uniform(-1, 1, n)
    frame = rdf.DeviceFrame.from_host({"k": [rdf.PrimitiveArray.from_numpy(k)], "x": [rdf.PrimitiveArray.from_numpy(x)],
                                       "y": [rdf.PrimitiveArray.from_numpy(y)]}, ctx=ctx)
    g = frame.group_aggregate(["k"], [("x", "sum"), ("x", "max"), ("y", "sum"), ("y", "count")])
    assert list(g.columns) == ["k", "sum(x)", "max(x)", "sum(y)", "count(y)"]          # the schema Dataset::try_aggregate plans
    assert g.schema["count(y)"] == rdf.U32 and g.schema["sum(x)"] == rdf.I32
    g = g.to_host()
    want = pd.DataFrame({"k": k, "x": x, "y": y}).groupby("k").agg(sum_x=("x", "sum"), max_x=("x", "max"), sum_y=("y", "sum"), count_y=("y", "count"))
    assert np.array_equal(g["k"][0].value_slice(), want.index.to_numpy())
    assert np.array_equal(g["sum(x)"][0].value_slice(), want["sum_x"].to_numpy().astype(np.int32))
    assert np.array_equal(g["max(x)"][0].value_slice(), want["max_x"].to_numpy())
    assert np.allclose(g["sum(y)"][0].value_slice(), want["sum_y"].to_numpy(), rtol=0, atol=1e-9)
    assert np.array_equal(g["count(y)"][0].value_slice(), want["count_y"].to_numpy().astype(np.uint32))
    with pytest.raises(rdf.ComputeError):
        frame.group_aggregate("nope", [("x", "sum")])
    with pytest.raises(rdf.ComputeError):
        frame.group_aggregate("k", [("nope", "sum")])
    with pytest.raises(rdf.UnsupportedType):
        frame.group_aggregate("k", [("x", "avg")])
    with pytest.raises(rdf.UnsupportedType):
        frame.group_aggregate("k", [("y", "max")])                                      # T::Native: Ord
    with pytest.raises(rdf.ReferencePanic):
        frame.evaluate([("group_aggregate", None)])


@pytest.mark.gpu
def test_group_aggregate_hot_keys(rdf, ctx, oracle):
    """Skew: two hot keys (hundreds of thousands of rows each: folded by k_group_big, one CTA per 64 Ki-row segment) among many
    small groups, nullable values and keys; integers exact, floats within the sum tolerance and bit-reproducible."""
    rng = np.random.default_rng(21)
    n = 1_500_000
    k = rng.integers(0, 5000, n).astype(np.int64)
    hot = rng.random(n)
    k[hot < 0.45] = 17
    k[(hot >= 0.45) & (hot < 0.6)] = -3
    km = rng.random(n) >= 0.02
    vi = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64, endpoint=True)
    vf = rng.uniform(-1e3, 1e3, n)
    vm = rng.random(n) >= 0.1
    lens = [700_000, 800_000]
    kch, ich, fch = _chunks(rdf, k, km, lens), _chunks(rdf, vi, vm, lens), _chunks(rdf, vf, None, lens)
    ck, ci, cf = rdf.Column.upload(kch, ctx=ctx), rdf.Column.upload(ich, ctx=ctx), rdf.Column.upload(fch, ctx=ctx)
    keys, res = rdf.group_aggregate(ck, [ci, cf])
    okeys, okvalid, wi = oracle.group_aggregate(kch, ich)
    _, _, wf = oracle.group_aggregate(kch, fch)
    gk, gkv, _ = _col_values(keys)
    assert np.array_equal(gkv, okvalid) and np.array_equal(gk[gkv], okeys[okvalid])
    assert wi["count"].max() > 400_000                                   # the hot key really is hot
    assert np.array_equal(_col_values(res[0]["count"])[0], wi["count"])
    assert np.array_equal(_col_values(res[0]["sum"])[0], wi["sum"])
    for key in ("min", "max"):
        g, gv, _ = _col_values(res[0][key])
        assert np.array_equal(gv, wi[key][1]) and np.array_equal(g[gv], wi[key][0][wi[key][1]])
    fs = _col_values(res[1]["sum"])[0]
    exact, mag = wf["exact"]
    assert np.all(np.abs(fs.astype(np.longdouble) - exact) <= 16 * np.log2(np.maximum(wf["count"], 2)) * 2.0 ** -53 * mag + 1e-300)
    keys2, res2 = rdf.group_aggregate(ck, [cf])
    assert np.array_equal(_col_values(res2[0]["sum"])[0].view(np.uint64), fs.view(np.uint64))   # same bits every run


@pytest.mark.gpu
def test_group_aggregate_one_hot_key_among_singletons(rdf, ctx, oracle):
    """Almost every key is its own group (one lane per group) and ONE key owns 100k rows (deferred to k_group_big): the min/max
    validity word the hot group shares with 31 singleton groups is written by both kernels."""
    rng = np.random.default_rng(5)
    n = 600_000
    k = rng.permutation(n).astype(np.int64)
    k[rng.choice(n, 100_000, replace=False)] = 300_001
    v = rng.integers(-2 ** 40, 2 ** 40, n).astype(np.int64)
    vm = rng.random(n) >= 0.5
    lens = [n]
    kch, vch = _chunks(rdf, k, None, lens), _chunks(rdf, v, vm, lens)
    ck, cv = rdf.Column.upload(kch, ctx=ctx), rdf.Column.upload(vch, ctx=ctx)
    keys, res = rdf.group_aggregate(ck, [cv])
    okeys, okvalid, want = oracle.group_aggregate(kch, vch)
    gk, gkv, _ = _col_values(keys)
    assert np.array_equal(gk, okeys) and gkv.all()
    assert want["count"].max() > 40_000
    assert np.array_equal(_col_values(res[0]["count"])[0], want["count"])
    assert np.array_equal(_col_values(res[0]["sum"])[0], want["sum"])
    for key in ("min", "max"):
        g, gv, arr = _col_values(res[0][key])
        assert np.array_equal(gv, want[key][1]) and np.array_equal(g[gv], want[key][0][want[key][1]])
        assert arr.null_count == int((~want[key][1]).sum())


def test_oracle_group_aggregate_small_cases_against_pandas(oracle):
    """Many tiny random shapes (empty, one row, all-null keys, all-null values, several chunks) of the numpy restatement against
    pandas' groupby -- the checker of the GPU tests is itself checked on the shapes where off-by-one errors live."""
    import pandas as pd
    from hypothesis import given, settings, strategies as st

    class Ch:
        def __init__(self, values, mask):
            self.values, self.offset, self.length, self._m = values, 0, len(values), mask

        def valid_mask(self):
            return self._m

    @settings(max_examples=80, deadline=None)
    @given(st.integers(0, 40), st.integers(1, 6), st.floats(0, 1), st.floats(0, 1), st.integers(0, 2 ** 31), st.integers(1, 3))
    def check(n, card, knull, vnull, seed, n_chunks):
        rng = np.random.default_rng(seed)
        k = rng.integers(-card, card, n).astype(np.int64)
        v = rng.integers(-100, 100, n).astype(np.int32)
        km, vm = rng.random(n) >= knull, rng.random(n) >= vnull
        cuts = sorted(rng.integers(0, n + 1, n_chunks - 1).tolist()) if n_chunks > 1 else []
        bounds = [0] + cuts + [n]
        kch = [Ch(k[a:b], km[a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
        vch = [Ch(v[a:b], vm[a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
        keys, kvalid, out = oracle.group_aggregate(kch, vch)
        df = pd.DataFrame({"k": pd.array(np.where(km, k, 0), dtype="Int64").astype(object), "v": v.astype(np.int64), "ok": vm})
        df.loc[~km, "k"] = None
        df["v_valid"] = df["v"].where(df["ok"])
        want = df.groupby("k", dropna=False, sort=True).agg(s=("v_valid", "sum"), c=("ok", "sum"), mn=("v_valid", "min"), mx=("v_valid", "max"))
        assert len(keys) == len(want)
        got_keys = [int(x) if ok else None for x, ok in zip(keys, kvalid)]
        valid_keys = [x for x in got_keys if x is not None]
        assert valid_keys == sorted(valid_keys) and got_keys.count(None) <= 1 and (None not in got_keys or got_keys[-1] is None)
        rows = {(None if pd.isna(x) else int(x)): r for x, r in zip(want.index, want.itertuples(index=False))}   # pandas' own order is not ours
        assert set(rows) == set(got_keys)
        for i, key in enumerate(got_keys):
            r = rows[key]
            assert int(out["count"][i]) == int(r.c) and int(out["sum"][i]) == (0 if pd.isna(r.s) else int(r.s))
            for name, w in (("min", r.mn), ("max", r.mx)):
                vals, valid = out[name]
                assert bool(valid[i]) == (not pd.isna(w))
                if valid[i]:
                    assert int(vals[i]) == int(w)

    check()
