#!/usr/bin/env python
"""N4 timing: an Arrow IPC file (2 x Float64, --rows rows, 25 RecordBatches) to device columns and back.

    file -> device   bdf_ipc_open + bdf_ipc_read (page cache -> pinned staging -> HBM), wall clock incl. the final sync
    device -> file   bdf_ipc_write (HBM -> pinned -> write(2))
    baseline         pyarrow: open_file().read_all() into host memory (zero-copy over a memory map), then the same
                     columns through Column.upload (what a caller without bdf_ipc_read would do)
The file lives in --dir (default /dev/shm when present, so the numbers measure the software path, not a disk)."""
import argparse
import os
import sys
import time

import numpy as np
import pyarrow as pa
import pyarrow.ipc

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_dataframe_b200 as rdf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
ctx = rdf.default_context()
nb = 25
n = args.rows // nb
path = os.path.join(args.dir, "bdf_ipc_bench.arrow")
out = os.path.join(args.dir, "bdf_ipc_bench_out.arrow")
rng = np.random.default_rng(0)
base = [rng.normal(0, 1, n), rng.normal(0, 1, n)]
with pa.ipc.new_file(path, pa.schema([("a", pa.float64()), ("b", pa.float64())])) as w:
    for k in range(nb):
        w.write_batch(pa.record_batch([pa.array(base[0] + k), pa.array(base[1] - k)], names=["a", "b"]))
gb = os.path.getsize(path) / 1e9


def best(fn):
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter(); r = fn(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
        del r
    return min(ts)


def lib_read():
    with rdf.IpcFile(path) as f:
        cols = f.read(["a", "b"])
        ctx.synchronize()
    return cols


def arrow_then_upload():
    with pa.memory_map(path) as src:
        t = pa.ipc.open_file(src).read_all()
        chunks = [[rdf.PrimitiveArray.from_arrow(c) for c in t.column(name).chunks] for name in ("a", "b")]
        cols = rdf.Column.upload_many(chunks)
        ctx.synchronize()
    return cols


t_lib = best(lib_read)
t_pa = best(arrow_then_upload)
cols = lib_read()
t_w = best(lambda: rdf.write_ipc(out, cols))
with pa.ipc.open_file(out) as r, pa.ipc.open_file(path) as r0:
    assert r.get_batch(nb - 1).equals(r0.get_batch(nb - 1))
print(f"file {gb:.2f} GB in {args.dir}: file->device {t_lib * 1e3:8.1f} ms ({gb / t_lib:5.1f} GB/s)   "
      f"pyarrow mmap + upload_many {t_pa * 1e3:8.1f} ms ({gb / t_pa:5.1f} GB/s)   device->file {t_w * 1e3:8.1f} ms ({gb / t_w:5.1f} GB/s)")
for p in (path, out):
    os.remove(p)
