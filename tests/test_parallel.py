"""Host-side multi-GPU logic on CPU: world_size-2 gloo process groups (127.0.0.1 rendezvous).

Without a GPU the per-rank partial aggregates come from the CPU oracle (standing in for the per-GPU reduction
kernel); what is under test is the sharding map and the partial-aggregate combine -- the one collective of
the path.  The `gpu` variant at the bottom runs the same two-rank combine with the real kernels producing
the partials (both ranks on cuda:0, gloo for the exchange)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Chunk:
    def __init__(self, values, dtype, mask=None):
        self.values, self.dtype, self.offset, self.length = np.ascontiguousarray(values), dtype, 0, len(values)
        self.validity = None if mask is None else np.packbits(np.asarray(mask, bool), bitorder="little")
        self.null_count = 0 if mask is None else int((~np.asarray(mask, bool)).sum())

    def valid_mask(self):
        return np.ones(self.length, bool) if self.validity is None else np.unpackbits(self.validity, bitorder="little")[: self.length].astype(bool)


def make_chunks(dtype_name, seed, lens, null_frac, orc):
    dtype = getattr(orc, dtype_name)
    npdt = np.dtype(orc.NP_DTYPES[dtype])
    rng = np.random.default_rng(seed)
    out = []
    for n in lens:
        if npdt.kind == "f":
            v = rng.uniform(-1e3, 1e3, n).astype(npdt)
        else:
            info = np.iinfo(npdt)
            v = rng.integers(info.min, info.max, n, dtype=npdt, endpoint=True)
        out.append(Chunk(v, dtype, rng.random(n) >= null_frac if null_frac else None))
    return dtype, out


def oracle_partials(orc, dtype, chunks):
    """What one GPU would report for its shard (AggregateFunctions.all), computed by the oracle."""
    npdt = np.dtype(orc.NP_DTYPES[dtype])
    res = {"sum": npdt.type(0), "min": None, "max": None, "count": 0, "rows": sum(c.length for c in chunks), "would_panic": False}
    if not chunks:
        return res
    res["sum"] = orc.aggregate(orc.SUM, dtype, chunks)[1]
    res["count"] = int(orc.aggregate(orc.COUNT, dtype, chunks)[1])
    nonempty = [c for c in chunks if c.valid_mask().any()]
    res["would_panic"] = len(nonempty) != len(chunks)
    if npdt.kind != "f" and nonempty:
        res["min"] = orc.aggregate(orc.MIN, dtype, nonempty)[1]
        res["max"] = orc.aggregate(orc.MAX, dtype, nonempty)[1]
    return res


CASES = [("I64", 0.1), ("U64", 0.1), ("I8", 0.0), ("U16", 0.3), ("I32", 0.97), ("F64", 0.1), ("F32", 0.0)]
LENS = [1000, 1, 0, 5000, 33, 2048, 7]


def _worker(rank, world, port, use_gpu, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        import rust_dataframe_b200 as rdf
        from rust_dataframe_b200 import parallel

        results = {}
        for name, null_frac in CASES:
            dtype, chunks = make_chunks(name, 99, LENS, null_frac, orc)   # same column on every rank
            for balanced in (False, True):
                mine = parallel.shard(chunks, rank, world, balanced=balanced)
                if use_gpu:
                    arrs = [rdf.PrimitiveArray(dtype, c.values, c.validity, 0, c.length, c.null_count) for c in mine]
                    local = rdf.AggregateFunctions.all(arrs, dtype=dtype) if arrs else oracle_partials(orc, dtype, [])
                else:
                    local = oracle_partials(orc, dtype, mine)
                combined = parallel.combine_aggregates(local, dtype, device="cpu")
                results[(name, balanced)] = {k: (None if v is None else (float(v) if name.startswith("F") and k == "sum" else (bool(v) if k == "would_panic" else int(v))))
                                             for k, v in combined.items()}
        # elementwise operators need no collective: every rank's shard result equals the same chunks of the full result
        results["shards"] = parallel.shard_indices(len(LENS), rank, world), parallel.shard_indices(len(LENS), rank, world, LENS)
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def _run(world, use_gpu):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, use_gpu, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def _check(got, world):
    from oracle import pyoracle as orc

    for name, null_frac in CASES:
        dtype, chunks = make_chunks(name, 99, LENS, null_frac, orc)
        want = oracle_partials(orc, dtype, chunks)
        for balanced in (False, True):
            per_rank = [got[r][(name, balanced)] for r in range(world)]
            assert all(p == per_rank[0] for p in per_rank), f"{name}: ranks disagree"
            res = per_rank[0]
            assert res["count"] == want["count"] and res["rows"] == want["rows"]
            if name.startswith("F"):
                exact, sum_abs = orc.sum_exact(dtype, chunks)
                eps = 2.0 ** -53 if name == "F64" else 2.0 ** -24
                assert abs(np.longdouble(res["sum"]) - exact) <= 16 * np.log2(max(2, want["rows"])) * eps * sum_abs + 1e-300
                assert res["min"] is None and res["max"] is None
            else:  # identical to the single-process result for every world size
                assert res["sum"] == int(want["sum"])
                assert res["min"] == (None if want["min"] is None else int(want["min"]))
                assert res["max"] == (None if want["max"] is None else int(want["max"]))
    owners = [got[r]["shards"] for r in range(world)]
    for variant in (0, 1):  # round robin and balanced: a partition of the chunks
        allidx = sorted(i for o in owners for i in o[variant])
        assert allidx == list(range(len(LENS)))
    assert owners[0][0] == [0, 2, 4, 6][: len(owners[0][0])] or world != 2


def test_two_rank_combine_gloo_cpu():
    _check(_run(2, use_gpu=False), 2)


def test_three_rank_combine_gloo_cpu():
    _check(_run(3, use_gpu=False), 3)


def test_shard_map_properties():
    import sys
    sys.path.insert(0, ROOT)
    from rust_dataframe_b200 import parallel

    lens = [4_000_000] * 25
    for world in (1, 2, 4, 8):
        parts = [parallel.shard_indices(25, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(25))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        bal = [parallel.shard_indices(25, r, world, lens) for r in range(world)]
        assert sorted(i for p in bal for i in p) == list(range(25))
    skew = [10_000_000, 1, 1, 1, 5_000_000, 5_000_000]
    bal = [parallel.shard_indices(6, r, 2, skew) for r in range(2)]
    loads = [sum(skew[i] for i in p) for p in bal]
    assert abs(loads[0] - loads[1]) <= 2


@pytest.mark.gpu
def test_two_rank_combine_real_kernels():
    _check(_run(2, use_gpu=True), 2)


# ---- sharded from_arrow: every rank opens the same IPC file and takes its RecordBatches by the shard map ---------------

def _ipc_worker(rank, world, port, path, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        import rust_dataframe_b200 as rdf
        from rust_dataframe_b200 import parallel

        with rdf.IpcFile(path) as f:
            lens = [f.batch_rows(b) for b in range(f.num_batches)]
            mine = parallel.shard_indices(f.num_batches, rank, world, lens)
            out = {"batches": mine}
            for name, dtype, _ in f.schema:
                if dtype < 0 or dtype == 10:
                    continue
                views = [f.view(b, name) for b in mine]          # zero-copy host views into the mapping
                chunks = [Chunk(v.value_slice(), dtype, None if v.validity is None else v.valid_mask()) for v in views]
                local = oracle_partials(orc, dtype, chunks)      # stands in for the per-GPU reduction of bdf_ipc_read_batches' columns
                comb = parallel.combine_aggregates(local, dtype, device="cpu")
                out[name] = {k: (None if v is None else (float(v) if k == "sum" and dtype >= 8 else (bool(v) if k == "would_panic" else int(v)))) for k, v in comb.items()}
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_sharded_ipc_read_world_size_2(tmp_path):
    import pyarrow as pa
    import pyarrow.ipc
    from oracle import pyoracle as orc

    rng = np.random.default_rng(17)
    lens = [300, 0, 5000, 17, 2048, 900]
    batches = [pa.record_batch([pa.array(rng.integers(-2 ** 62, 2 ** 62, n), pa.int64(), mask=rng.random(n) < 0.2), pa.array(rng.normal(0, 100, n), pa.float64()),
                                pa.array([str(i) for i in range(n)], pa.string()), pa.array(rng.integers(0, 255, n).astype(np.uint8), pa.uint8())],
                               names=["k", "x", "s", "u"]) for n in lens]
    path = str(tmp_path / "sharded.arrow")
    with pa.ipc.new_file(path, batches[0].schema) as w:
        for b in batches:
            w.write_batch(b)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(got[0]["batches"] + got[1]["batches"]) == list(range(len(lens)))
    assert got[0]["k"] == got[1]["k"] and got[0]["x"] == got[1]["x"] and got[0]["u"] == got[1]["u"]
    k = pa.chunked_array([b.column("k") for b in batches])
    kv = k.drop_null().to_numpy()
    assert got[0]["k"]["count"] == len(kv) and got[0]["k"]["rows"] == sum(lens)
    assert got[0]["k"]["sum"] == int(np.sum(kv.astype(np.uint64), dtype=np.uint64).astype(np.int64))     # wrapping, order independent
    assert got[0]["k"]["min"] == int(kv.min()) and got[0]["k"]["max"] == int(kv.max())
    xv = np.concatenate([b.column("x").to_numpy() for b in batches])
    assert abs(got[0]["x"]["sum"] - float(np.sum(xv.astype(np.longdouble)))) <= 1e-6
    assert got[0]["u"]["sum"] == int(np.sum(np.concatenate([b.column("u").to_numpy() for b in batches]).astype(np.uint64)) % 256)
