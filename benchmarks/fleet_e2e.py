"""Drop-in e2e through ONE multi-GPU context (bdf_init_multi): what a single Rust process gets when its ScalarFunctions::add /
AggregateFunctions::sum bodies call libb200df -- the library shards the 25 chunks over the GPUs, each GPU uses its own PCIe link.
1e8 rows x 2 Float64 (pageable host buffers), c into freshly allocated / reused pageable buffers / pinned buffers.

    python benchmarks/fleet_e2e.py [max_gpus]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_dataframe_b200 as rdf  # noqa: E402
from rust_dataframe_b200 import native as N  # noqa: E402

ROWS, CHUNK = 100_000_000, 4_000_000


def main():
    import torch

    max_g = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    lens = [CHUNK] * (ROWS // CHUNK)
    rng = np.random.default_rng(0)
    a = [rdf.PrimitiveArray.from_numpy(rng.uniform(-1e3, 1e3, n)) for n in lens]
    b = [rdf.PrimitiveArray.from_numpy(rng.uniform(-1e3, 1e3, n)) for n in lens]
    want = float(sum(np.sum(x.value_slice() + y.value_slice(), dtype=np.longdouble) for x, y in zip(a, b)))
    out = []
    for g in [k for k in (1, 2, 4, 8) if k <= max_g]:
        ctx = rdf.Context.multi(g)
        reuse = N.alloc_outputs(rdf.F64, lens, ctx, pinned=False)
        for v, _, _ in reuse[1]:
            v[:] = 0.0
        pin_out = N.alloc_outputs(rdf.F64, lens, ctx, pinned=True)
        pa = [ctx.pinned_array(rdf.F64, x.value_slice()) for x in a]
        pb = [ctx.pinned_array(rdf.F64, x.value_slice()) for x in b]

        def dropin(x, y, into):
            outs, bufs = into if into is not None else N.alloc_outputs(rdf.F64, lens, ctx, pinned=False)
            for i in range(len(bufs)):
                outs[i].len = bufs[i][0].shape[0]
            N.raise_for_status(N.lib().bdf_binary(ctx.handle, N.ADD, rdf.F64, len(x), N.make_views(x), len(y), N.make_views(y), outs))
            c = N.collect_outputs(rdf.F64, outs, bufs)
            return float(rdf.AggregateFunctions.sum(c, dtype=rdf.F64, ctx=ctx))

        row = {"n_gpus": g}
        for name, fn in (("pageable_fresh", lambda: dropin(a, b, None)), ("pageable_reused", lambda: dropin(a, b, reuse)),
                         ("pinned", lambda: dropin(pa, pb, pin_out))):
            for _ in range(2):
                s = fn()
            assert abs(s - want) <= 1e-6 * abs(want) + 1e-3, (s, want)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                fn()
            dt = (time.perf_counter() - t0) / reps
            row[name] = {"ms_per_step": dt * 1e3, "rows_per_s": ROWS / dt, "pcie_GBs": 3.2e9 / dt / 1e9}
        out.append(row)
        print(json.dumps(row), flush=True)
        ctx.close()
    return out


if __name__ == "__main__":
    main()
