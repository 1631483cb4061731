import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import rust_dataframe_b200 as rdf
ctx = rdf.Context(0)
CH, N = 4_000_000, 25
rng = np.random.default_rng(0)
host = [rdf.PrimitiveArray.from_numpy(rng.uniform(-1, 1, CH)) for _ in range(N)]
pinned = [ctx.pinned_array(rdf.F64, h.values) for h in host]
for name, arrs in (("pageable", host), ("pinned", pinned)):
    for _ in range(2):
        c = rdf.Column.upload(arrs, ctx=ctx); c.free()
    t0 = time.perf_counter()
    for _ in range(4):
        c = rdf.Column.upload(arrs, ctx=ctx); c.free()
    dt = (time.perf_counter() - t0) / 4
    print(f"threads={os.environ.get('BDF_COPY_THREADS','default')} upload {name}: {dt*1e3:.1f} ms {0.8/dt:.1f} GB/s", flush=True)
c = rdf.Column.upload(pinned, ctx=ctx)
outs = [np.empty(CH) for _ in range(N)]
for o in outs: o[:] = 0   # pre-faulted pageable destination
from rust_dataframe_b200 import native as Nn
import ctypes as C
def dl(pre):
    into = Nn.alloc_outputs(rdf.F64, [CH]*N, ctx)
    if pre:
        for (v, b, k) in into[1]: v[:] = 0
    t0 = time.perf_counter(); c.download(into=into); return time.perf_counter() - t0
for pre in (False, True):
    dl(pre); ts = [dl(pre) for _ in range(3)]
    print(f"download to pageable (pre-faulted={pre}): {min(ts)*1e3:.1f} ms {0.8/min(ts):.1f} GB/s", flush=True)
