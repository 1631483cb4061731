"""Throughput of the host-in/host-out drop-in entry (bdf_binary through ScalarFunctions.add) with pageable vs pinned
Arrow buffers, 1e8 f64 rows in 25 chunks.  Complements bench.py's e2e (which uses the device-column chain)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rust_dataframe_b200 as rdf

ctx = rdf.default_context()
CH, N = 4_000_000, 25
rng = np.random.default_rng(0)
pa = [rng.uniform(-1e3, 1e3, CH) for _ in range(N)]
pb = [rng.uniform(-1e3, 1e3, CH) for _ in range(N)]
pageable_a = [rdf.PrimitiveArray.from_numpy(v) for v in pa]
pageable_b = [rdf.PrimitiveArray.from_numpy(v) for v in pb]
pinned_a = [ctx.pinned_array(rdf.F64, v) for v in pa]
pinned_b = [ctx.pinned_array(rdf.F64, v) for v in pb]
for name, a, b, pin_out in (("pageable in/out", pageable_a, pageable_b, False), ("pinned in, pageable out", pinned_a, pinned_b, False),
                            ("pinned in/out", pinned_a, pinned_b, True)):
    for _ in range(2):
        out = rdf.ScalarFunctions.add(a, b, pinned_out=pin_out)
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        out = rdf.ScalarFunctions.add(a, b, pinned_out=pin_out)
    dt = (time.perf_counter() - t0) / reps
    print(f"ScalarFunctions.add 1e8 f64 rows, {name:26s}: {dt * 1e3:8.2f} ms  {1e8 / dt:10.3e} rows/s  {2.4e9 / dt / 1e9:6.1f} GB/s over PCIe (incl. output allocation)", flush=True)
t0 = time.perf_counter()
for _ in range(2):
    s = [x + y for x, y in zip(pa, pb)]
print(f"numpy add on one core: {(time.perf_counter() - t0) / 2 * 1e3:.1f} ms")
