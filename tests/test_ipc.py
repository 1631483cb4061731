"""N4: Arrow IPC files either side of the path (DataFrame::from_arrow / to_arrow,
src/dataframe.rs:391-407, 515-525).  The library decodes and encodes the format itself; the checker is pyarrow
(an independent Arrow implementation): files written by pyarrow must read back identically through bdf_ipc_*,
files written by bdf_ipc_write* must pass pyarrow's flatbuffers verifier and read back identically.
The parser/writer tests need no GPU; the device round trips are marked gpu."""
import os

import numpy as np
import pyarrow as pa
import pyarrow.ipc
import pytest

NUMERIC = [("i8", pa.int8()), ("i16", pa.int16()), ("i32", pa.int32()), ("i64", pa.int64()), ("u8", pa.uint8()), ("u16", pa.uint16()),
           ("u32", pa.uint32()), ("u64", pa.uint64()), ("f32", pa.float32()), ("f64", pa.float64())]


def random_array(rng, typ, n, null_frac):
    if pa.types.is_boolean(typ):
        v = rng.random(n) > 0.5
    elif pa.types.is_floating(typ):
        v = rng.normal(0, 1e3, n).astype(typ.to_pandas_dtype())
    else:
        info = np.iinfo(typ.to_pandas_dtype())
        v = rng.integers(info.min, info.max, n, dtype=typ.to_pandas_dtype(), endpoint=True)
    mask = rng.random(n) < null_frac if null_frac else None
    return pa.array(v, type=typ, mask=mask)


def mixed_batches(rng, lens, null_frac=0.2):
    """Numeric + boolean columns interleaved with every kind of column the reader has to step over."""
    fields, batches = None, []
    for n in lens:
        cols, names = [], []

        def add(name, arr):
            names.append(name); cols.append(arr)
        add("s", pa.array([None if i % 7 == 0 else f"row{i}" for i in range(n)], pa.string()))
        for name, typ in NUMERIC[:5]:
            add(name, random_array(rng, typ, n, null_frac))
        add("lst", pa.array([[i, i + 1] if i % 3 else None for i in range(n)], pa.list_(pa.int32())))
        add("st", pa.array([{"x": i, "y": [float(i)]} for i in range(n)], pa.struct([("x", pa.int64()), ("y", pa.list_(pa.float64()))])))
        for name, typ in NUMERIC[5:]:
            add(name, random_array(rng, typ, n, 0.0 if name == "u32" else null_frac))
        add("dict", pa.DictionaryArray.from_arrays(pa.array([None if i % 4 == 2 else i % 2 for i in range(n)], pa.int8()), pa.array(["a", "b"])))
        add("b", random_array(rng, pa.bool_(), n, null_frac))
        add("d32", pa.array(np.arange(n, dtype=np.int32), pa.date32()))
        add("ts", pa.array(np.arange(n, dtype=np.int64), pa.timestamp("us")))
        add("fsl", pa.array([[1.0, 2.0, 3.0]] * n, pa.list_(pa.float32(), 3)))
        add("ls", pa.array([f"{i}" for i in range(n)], pa.large_string()))
        add("nul", pa.nulls(n))
        add("dec", pa.array([None if i % 5 == 0 else i for i in range(n)], pa.decimal128(12, 2)))
        add("b2", random_array(rng, pa.bool_(), n, 0.0))
        add("mp", pa.array([[("k", i)] for i in range(n)], pa.map_(pa.string(), pa.int32())))
        add("tail", random_array(rng, pa.float64(), n, null_frac))
        b = pa.record_batch(cols, names=names)
        fields = fields or b.schema
        batches.append(b)
    return fields, batches


def write_file(path, schema, batches, options=None):
    with pa.ipc.new_file(path, schema, options=options) as w:
        for b in batches:
            w.write_batch(b)


def check_view_against_pyarrow(arr, parr, what):
    """arr: PrimitiveArray/BooleanArray view from the library; parr: the pyarrow array of the same batch/column."""
    assert arr.length == len(parr), what
    want_valid = np.array(parr.is_valid().to_pylist(), dtype=bool) if len(parr) else np.zeros(0, bool)
    assert np.array_equal(arr.valid_mask(), want_valid), f"{what}: validity"
    if arr.validity is None or arr.null_count >= 0:   # -1 = "unknown" on a host-made slice
        assert (arr.null_count if arr.validity is not None else 0) == parr.null_count, f"{what}: null_count"
    if pa.types.is_boolean(parr.type):
        got = arr.value_bits()[want_valid]
        want = np.array(parr.drop_null().to_pylist(), dtype=bool)
    else:
        got = arr.value_slice()[want_valid]
        want = parr.drop_null().to_numpy(zero_copy_only=False)
        assert got.dtype == want.dtype, what
        got, want = got.view(f"u{got.dtype.itemsize}"), want.view(f"u{want.dtype.itemsize}")
    assert np.array_equal(got, want), f"{what}: values"


@pytest.mark.parametrize("variant", ["v5", "v4", "v4-legacy"])
def test_reader_matches_pyarrow(rdf, tmp_path, variant):
    rng = np.random.default_rng(7)
    schema, batches = mixed_batches(rng, [0, 1, 7, 64, 1000, 4097])
    opts = {"v5": None, "v4": pa.ipc.IpcWriteOptions(metadata_version=pa.ipc.MetadataVersion.V4),
            "v4-legacy": pa.ipc.IpcWriteOptions(metadata_version=pa.ipc.MetadataVersion.V4, use_legacy_format=True)}[variant]
    path = str(tmp_path / "mixed.arrow")
    write_file(path, schema, batches, opts)
    with rdf.IpcFile(path) as f:
        assert f.num_columns == len(schema) and f.num_batches == len(batches) and f.num_rows == sum(b.num_rows for b in batches)
        want_dtype = {name: getattr(rdf, name.upper()) for name, _ in NUMERIC}
        want_dtype.update({"b": 10, "b2": 10, "tail": rdf.F64})
        for (name, dtype, nullable), fld in zip(f.schema, schema):
            assert name == fld.name and nullable == fld.nullable
            assert dtype == want_dtype.get(name, -1), name
        for bi, b in enumerate(batches):
            assert f.batch_rows(bi) == b.num_rows
            for name in want_dtype:
                check_view_against_pyarrow(f.view(bi, name), b.column(name), f"{variant} batch {bi} column {name}")
        with pytest.raises(rdf.UnsupportedType):
            f.view(0, "s")
        with pytest.raises(rdf.ArrowError):
            f.view(len(batches), "i8")
        with pytest.raises(KeyError):
            f.view(0, "nope")


def test_reader_rejects_what_it_cannot_decode(rdf, tmp_path):
    rng = np.random.default_rng(1)
    t = pa.record_batch([random_array(rng, pa.float64(), 100, 0.1)], names=["x"])
    p = str(tmp_path / "z.arrow")
    try:
        write_file(p, t.schema, [t], pa.ipc.IpcWriteOptions(compression="lz4"))
        with pytest.raises(rdf.UnsupportedType):
            rdf.IpcFile(p)
    except (pa.ArrowNotImplementedError, pa.ArrowInvalid):
        pass   # this pyarrow build has no lz4: nothing to reject
    p = str(tmp_path / "stream.arrows")
    with pa.ipc.new_stream(p, t.schema) as w:
        w.write_batch(t)
    with pytest.raises(rdf.ArrowError):
        rdf.IpcFile(p)
    good = str(tmp_path / "good.arrow")
    write_file(good, t.schema, [t])
    raw = open(good, "rb").read()
    for cut, name in ((len(raw) - 3, "cut-magic"), (12, "tiny")):
        q = str(tmp_path / name)
        open(q, "wb").write(raw[:cut])
        with pytest.raises(rdf.ArrowError):
            rdf.IpcFile(q)
    bad = bytearray(raw)
    bad[-10:-6] = (2 ** 31 - 1).to_bytes(4, "little")   # absurd footer length
    q = str(tmp_path / "badfooter")
    open(q, "wb").write(bytes(bad))
    with pytest.raises(rdf.ArrowError):
        rdf.IpcFile(q)
    with pytest.raises(rdf.ArrowError):
        rdf.IpcFile(str(tmp_path / "does-not-exist"))
    u = pa.record_batch([pa.UnionArray.from_sparse(pa.array([0, 1], pa.int8()), [pa.array([1, 2]), pa.array(["a", "b"])])], names=["u"])
    q = str(tmp_path / "union.arrow")
    write_file(q, u.schema, [u])
    with pytest.raises(rdf.UnsupportedType):
        rdf.IpcFile(q)
    empty = str(tmp_path / "empty.arrow")
    write_file(empty, t.schema, [])
    with rdf.IpcFile(empty) as f:
        assert f.num_batches == 0 and f.num_rows == 0 and f.schema == [("x", rdf.F64, True)]


def host_columns(rdf, rng, lens, sliced):
    cols = {}
    for k, (name, _) in enumerate(NUMERIC):
        npdt = rdf.NP_DTYPES[getattr(rdf, name.upper())]
        chunks = []
        for j, n in enumerate(lens):
            pad = (3 + 5 * j + k) % 19 if sliced else 0
            v = (rng.normal(0, 100, n + pad + 2) if np.dtype(npdt).kind == "f" else rng.integers(0, 100, n + pad + 2)).astype(npdt)
            mask = (rng.random(n + pad + 2) > 0.25) if (k + j) % 3 else None
            a = rdf.PrimitiveArray.from_numpy(v, mask)
            a.null_count = -1 if mask is not None else 0
            chunks.append(a.slice(pad, n))
        cols[name] = chunks
    bools = []
    for j, n in enumerate(lens):
        pad = (2 + 3 * j) % 11 if sliced else 0
        bools.append(rdf.BooleanArray.from_numpy(rng.random(n + pad) > 0.4, (rng.random(n + pad) > 0.2) if j % 2 else None).slice(pad, n))
    cols["flag"] = bools
    return cols


def check_file_against_host(rdf, path, cols, lens):
    with pa.ipc.open_file(path) as r:   # runs pyarrow's metadata verifier
        assert r.num_record_batches == len(lens)
        assert r.schema.names == list(cols)
        for bi in range(len(lens)):
            b = r.get_batch(bi)
            b.validate(full=True)
            assert b.num_rows == lens[bi]
            for name, chunks in cols.items():
                check_view_against_pyarrow(chunks[bi], b.column(name), f"written batch {bi} column {name}")
    with rdf.IpcFile(path) as f:        # and the library reads its own files
        assert f.num_batches == len(lens) and [s[0] for s in f.schema] == list(cols)
        for bi in range(len(lens)):
            for name, chunks in cols.items():
                got, want = f.view(bi, name), chunks[bi]
                assert np.array_equal(got.valid_mask(), want.valid_mask())
                if name == "flag":
                    assert np.array_equal(got.value_bits()[want.valid_mask()], want.value_bits()[want.valid_mask()])
                else:
                    assert np.array_equal(got.value_slice()[want.valid_mask()], want.value_slice()[want.valid_mask()])


@pytest.mark.parametrize("sliced", [False, True])
def test_writer_is_read_by_pyarrow(rdf, tmp_path, sliced):
    rng = np.random.default_rng(3 + sliced)
    lens = [5, 0, 64, 1001, 4096]
    cols = host_columns(rdf, rng, lens, sliced)
    path = str(tmp_path / "out.arrow")
    rdf.write_ipc_host(path, cols)
    check_file_against_host(rdf, path, cols, lens)
    # the body of every buffer starts on a 64-byte file offset (what the device upload likes)
    with rdf.IpcFile(path) as f:
        base = np.frombuffer(open(path, "rb").read(16), np.uint8)  # noqa: F841 (keeps the file in the page cache)
        v = f.view(3, "f64")
        assert v.values.ctypes.data % 64 == 0
    rdf.write_ipc_host(path, {"only": []})   # zero batches: schema + footer only
    with pa.ipc.open_file(path) as r:
        assert r.num_record_batches == 0 and r.schema.names == ["only"]
    with pytest.raises(rdf.ComputeError):
        rdf.write_ipc_host(path, {"a": cols["i8"], "b": [c.slice(0, max(c.length - 1, 0)) for c in cols["i16"]]})
    with pytest.raises(rdf.ArrowError):
        rdf.write_ipc_host(str(tmp_path / "no" / "such" / "dir.arrow"), {"a": cols["i8"]})


# ---- device round trips -----------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_file_to_device_and_back(rdf, ctx, oracle, tmp_path):
    """from_arrow -> compute on the device -> to_arrow, checked by pyarrow on both ends."""
    rng = np.random.default_rng(12)
    lens = [0, 1, 2047, 2049, 70_001, 300_000]
    schema, batches = mixed_batches(rng, lens, null_frac=0.15)
    src = str(tmp_path / "in.arrow")
    write_file(src, schema, batches)
    with rdf.IpcFile(src) as f:
        cols = f.read()                                    # every numeric/boolean column
        assert set(cols) == {n for n, _ in NUMERIC} | {"b", "b2", "tail"}
        for name, col in cols.items():                     # device copy == file contents
            for bi, g in enumerate(col.download()):
                check_view_against_pyarrow(g, batches[bi].column(name), f"device copy batch {bi} column {name}")
        a, t = cols["f64"], cols["tail"]
        s = a.add(t)
        kept = s.filter(cols["b"])
        h = rdf.eval_expr([a, t], [(rdf.native.MUL, 0, 1), ("sin", 2)])
        with pytest.raises(rdf.UnsupportedType):
            f.read(["s"])
        lazy = f.read(["i32", "u8"], asynchronous=True)
        assert lazy["i32"].count() == sum(len(b.column("i32")) - b.column("i32").null_count for b in batches)
    out = str(tmp_path / "out.arrow")
    rdf.write_ipc(out, {"f64": a, "tail": t, "sum": s, "h": h, "b": cols["b"], "i8": cols["i8"]})
    with pa.ipc.open_file(out) as r:
        assert r.num_record_batches == len(lens)
        for bi in range(len(lens)):
            b = r.get_batch(bi)
            b.validate(full=True)
            src_b = batches[bi]
            for name in ("f64", "tail", "b", "i8"):
                assert b.column(name).equals(src_b.column(name)), f"batch {bi} column {name} changed on the round trip"
            want = pa.compute.add(src_b.column("f64"), src_b.column("tail"))
            assert b.column("sum").equals(want), f"batch {bi}: device a+b differs from pyarrow's"
            got_h, ref_h = b.column("h"), pa.compute.sin(pa.compute.multiply(src_b.column("f64"), src_b.column("tail")))
            assert got_h.is_valid().equals(ref_h.is_valid())
            gv, rv = got_h.drop_null().to_numpy(zero_copy_only=False), ref_h.drop_null().to_numpy(zero_copy_only=False)
            assert np.allclose(gv, rv, rtol=0, atol=4 * 2.0 ** -53)
    # filter output has data-dependent chunk lengths: written as its own file
    out2 = str(tmp_path / "kept.arrow")
    rdf.write_ipc(out2, {"kept": kept})
    with pa.ipc.open_file(out2) as r:
        for bi in range(len(lens)):
            sb = batches[bi]
            want = pa.compute.add(sb.column("f64"), sb.column("tail")).filter(sb.column("b"), null_selection_behavior="drop")
            assert r.get_batch(bi).column("kept").equals(want)
    with pytest.raises(rdf.ComputeError):
        rdf.write_ipc(out2, {"a": a, "kept": kept})      # chunk lengths differ: not a set of RecordBatches


@pytest.mark.gpu
def test_sharded_read_combines_to_the_whole_file(rdf, ctx, tmp_path):
    """Multi-GPU shape of from_arrow: rank r of N reads RecordBatches by the shard map, aggregate partials combine
    to the single-reader result (here the N ranks are played in one process; the combine rules are the ones
    tests/test_parallel.py checks over gloo)."""
    from rust_dataframe_b200 import parallel

    rng = np.random.default_rng(5)
    lens = [100, 0, 4097, 30_000, 5, 77_777, 2048]
    batches = [pa.record_batch([random_array(rng, pa.int64(), n, 0.2), random_array(rng, pa.float64(), n, 0.1)], names=["k", "x"]) for n in lens]
    path = str(tmp_path / "sharded.arrow")
    write_file(path, batches[0].schema, batches)
    with rdf.IpcFile(path) as f:
        whole = f.read(["k"])["k"].aggregate_all()
        for world in (2, 3):
            parts, seen = [], []
            for rank in range(world):
                mine = parallel.shard_indices(f.num_batches, rank, world, lens=[f.batch_rows(b) for b in range(f.num_batches)])
                seen += mine
                col = f.read(["k", "x"], batches=mine)["k"]
                assert [col.chunk_info(i)["len"] for i in range(len(mine))] == [lens[b] for b in mine]
                parts.append(col.aggregate_all() if mine else None)
            assert sorted(seen) == list(range(len(lens)))
            parts = [p for p in parts if p is not None and p["count"] > 0]
            assert sum(p["count"] for p in parts) == whole["count"]
            assert (sum(int(p["sum"]) for p in parts) + 2 ** 63) % 2 ** 64 - 2 ** 63 == whole["sum"]
            assert min(p["min"] for p in parts) == whole["min"] and max(p["max"] for p in parts) == whole["max"]
        with pytest.raises(rdf.ArrowError):
            f.read(["k"], batches=[len(lens)])


@pytest.mark.gpu
def test_large_file_read_1e7(rdf, ctx, tmp_path):
    """2 x Float64 x 1e7 rows in 10 batches: file -> device add + fused sum equals numpy on the file's data."""
    rng = np.random.default_rng(2)
    n, nb = 1_000_000, 10
    batches = [pa.record_batch([pa.array(rng.normal(0, 10, n)), pa.array(rng.normal(0, 10, n), mask=rng.random(n) < 0.1)], names=["a", "b"])
               for _ in range(nb)]
    path = str(tmp_path / "big.arrow")
    write_file(path, batches[0].schema, batches)
    with rdf.IpcFile(path) as f:
        cols = f.read(["a", "b"])
    c, agg = cols["a"].binary_agg(rdf.native.ADD, cols["b"])
    want = [pa.compute.add(b.column("a"), b.column("b")) for b in batches]
    assert agg["count"] == sum(len(w) - w.null_count for w in want)
    tot = sum(float(np.sum(w.drop_null().to_numpy(zero_copy_only=False).astype(np.longdouble))) for w in want)
    assert abs(agg["sum"] - tot) <= 1e-6
    for bi, g in enumerate(c.download()):
        check_view_against_pyarrow(g, want[bi], f"a+b batch {bi}")


def test_reader_survives_corrupted_files(rdf, tmp_path):
    """Every offset in the footer and the messages is bounds-checked: a damaged file is an error (or still readable),
    never a crash, and a view that is handed out stays inside the mapping."""
    rng = np.random.default_rng(99)
    schema, batches = mixed_batches(rng, [3, 40, 300])
    good = str(tmp_path / "good.arrow")
    write_file(good, schema, batches)
    raw = np.frombuffer(open(good, "rb").read(), dtype=np.uint8)
    # metadata lives in the footer (file end) and in the message headers; damage both regions, and the body a little
    footer_len = int(np.frombuffer(raw[-10:-6].tobytes(), np.int32)[0])
    regions = [(len(raw) - 10 - footer_len, len(raw) - 6), (8, 2000)]
    opened = failed = 0
    for trial in range(400):
        bad = raw.copy()
        lo, hi = regions[trial % 2]
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(lo, hi))
            bad[pos] = rng.integers(0, 256) if trial % 3 else (0xFF if bad[pos] != 0xFF else 0x7F)
        p = str(tmp_path / "bad.arrow")
        bad.tofile(p)
        try:
            with rdf.IpcFile(p) as f:
                opened += 1
                for bi in range(f.num_batches):
                    for name, dtype, _ in f.schema:
                        if dtype >= 0:
                            v = f.view(bi, name)
                            if v.length:          # touching first and last byte must be legal
                                _ = v.values[0] if dtype != 10 else v.values[(v.length - 1) // 8]
                                _ = v.value_slice()[-1] if dtype != 10 else v.values[0]
                                if v.validity is not None:
                                    _ = v.validity[(v.length - 1) // 8]
        except rdf.ArrowError:
            failed += 1
    assert opened + failed == 400 and failed > 50


def _patch_i64(raw: np.ndarray, pos: int, value: int) -> np.ndarray:
    out = raw.copy()
    out[pos:pos + 8] = np.frombuffer(np.array([value], dtype=np.int64).tobytes(), dtype=np.uint8)
    return out


def test_reader_rejects_offsets_that_wrap_int64(rdf, tmp_path):
    """Buffer and block offsets/lengths near INT64_MAX: `off + len` would wrap negative and pass a naive bound check
    (the advisor's reproduction: values-buffer offset 2^63-101 was accepted and the first read of the view segfaulted).
    Every such file must be refused at open."""
    n = 12345
    batch = pa.record_batch([pa.array(np.arange(n, dtype=np.int64)), pa.array(np.arange(n, dtype=np.float64), mask=np.arange(n) % 5 == 0)], names=["k", "x"])
    good = str(tmp_path / "good.arrow")
    write_file(good, batch.schema, [batch])
    with rdf.IpcFile(good) as f:
        assert f.num_batches == 1 and f.view(0, "k").value_slice()[-1] == n - 1
    raw = np.frombuffer(open(good, "rb").read(), dtype=np.uint8)
    data = raw.tobytes()
    vlen = np.array([n * 8], dtype=np.int64).tobytes()
    # Buffer structs {offset:int64, length:int64} of the two values buffers: find the length words in the message metadata
    hits = []
    pos = data.find(vlen)
    while pos >= 0:
        hits.append(pos)
        pos = data.find(vlen, pos + 1)
    assert len(hits) >= 2, "values-buffer lengths not found in the metadata"
    i64max = (1 << 63) - 1
    cases = []
    for h in hits:
        cases.append(_patch_i64(raw, h - 8, i64max - 100))      # offset near INT64_MAX: off + len wraps
        cases.append(_patch_i64(raw, h - 8, -(1 << 62)))         # negative offset
        cases.append(_patch_i64(raw, h, i64max - 7))             # length near INT64_MAX
        both = _patch_i64(raw, h - 8, i64max - 100)
        cases.append(_patch_i64(both, h, i64max - 100))
    # the footer's Block {offset:int64, metaDataLength:int32, pad, bodyLength:int64}: bodyLength and offset near INT64_MAX
    footer_len = int(np.frombuffer(raw[-10:-6].tobytes(), np.int32)[0])
    fstart = len(raw) - 10 - footer_len
    for off in range(fstart, len(raw) - 10 - 24 + 1):
        o, ml, _, bl = np.frombuffer(raw[off:off + 24].tobytes(), dtype=[("o", "<i8"), ("m", "<i4"), ("p", "<i4"), ("b", "<i8")])[0]
        if 8 <= o < len(raw) and 8 <= ml < 4096 and 0 < bl < len(raw) and o + ml + bl <= len(raw):
            cases.append(_patch_i64(raw, off + 16, i64max - 3))
            cases.append(_patch_i64(raw, off, i64max - 3))
            cases.append(_patch_i64(_patch_i64(raw, off, i64max - 64), off + 16, i64max - 64))
            break
    else:
        raise AssertionError("record batch block not found in the footer")
    for k, bad in enumerate(cases):
        p = str(tmp_path / f"wrap{k}.arrow")
        bad.tofile(p)
        with pytest.raises(rdf.ArrowError):
            with rdf.IpcFile(p) as f:
                for name, dtype, _ in f.schema:   # if it were accepted, touching the view must still be legal -- but it must not be
                    v = f.view(0, name)
                    _ = v.value_slice()[-1]
                raise AssertionError(f"case {k}: a file with wrapping offsets was accepted")


def test_reader_bounds_schema_work(rdf, tmp_path):
    """A schema whose nested fields all point at the same child table costs k^depth visits unless the total is bounded:
    build a Struct field whose children vector lists ITSELF many times and check that open fails fast."""
    import time

    inner = pa.struct([("x", pa.int64())])
    batch = pa.record_batch([pa.array([{"x": 1}], inner)], names=["s"])
    good = str(tmp_path / "nest.arrow")
    write_file(good, batch.schema, [batch])
    raw = np.frombuffer(open(good, "rb").read(), dtype=np.uint8).copy()
    footer_len = int(np.frombuffer(raw[-10:-6].tobytes(), np.int32)[0])
    fstart = len(raw) - 10 - footer_len
    # find a children vector of length 1 inside the footer (u32 count == 1 followed by a u32 offset to a table) and make the
    # single child offset point back at the PARENT field table (a cycle): unbounded recursion without a budget
    patched = 0
    for vec in range(fstart, len(raw) - 10 - 8, 4):
        cnt = int(np.frombuffer(raw[vec:vec + 4].tobytes(), np.uint32)[0])
        if cnt != 1:
            continue
        for target in range(fstart, len(raw) - 10 - 4, 4):
            rel = target - (vec + 4)
            if rel <= 0:
                continue
            bad = raw.copy()
            bad[vec + 4:vec + 8] = np.frombuffer(np.array([rel], dtype=np.uint32).tobytes(), dtype=np.uint8)
            p = str(tmp_path / "cycle.arrow")
            bad.tofile(p)
            t0 = time.perf_counter()
            try:
                with rdf.IpcFile(p):
                    pass
            except rdf.ArrowError:
                pass
            assert time.perf_counter() - t0 < 2.0, "schema decoding did not terminate quickly"
            patched += 1
            if patched >= 300:
                return
    assert patched > 0


def test_a_file_without_columns_on_the_path_is_an_empty_frame(rdf, tmp_path):
    """Only Utf8 columns: nothing to read, so no device call is made at all (the C entry rejects n_cols == 0) -- works without a GPU."""
    path = str(tmp_path / "strings.arrow")
    table = pa.table({"city": pa.array(["Elgin", "Stoke-on-Trent", None]), "county": pa.array(["Moray", None, "x"])})
    with pa.OSFile(path, "wb") as sink, pa.ipc.new_file(sink, table.schema) as w:
        w.write_table(table)
    with rdf.IpcFile(path) as f:
        assert [dt for _, dt, _ in f.schema] == [-1, -1]
        assert f.read() == {}
    frame = rdf.DeviceFrame.from_arrow(path)
    assert list(frame.columns) == []
