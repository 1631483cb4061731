"""Multi-GPU parity: the grouped ncclAllReduce that libb200df enqueues behind the per-GPU reductions.
Needs >= 2 GPUs on the box (skipped otherwise; the host-side shard/fold logic is covered on CPU by tests/test_parallel.py).
One process per GPU, launched exactly like the driver launches bench.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(world: int, script: str, *extra: str, timeout: int = 900, env=None) -> str:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, f"rc={r.returncode}\n{r.stdout[-4000:]}\n{r.stderr[-6000:]}"
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("combine", ["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_aggregates_match_the_oracle(world, combine):
    """Both transports of the combine: the grouped ncclAllReduce and the NVLink peer-memory mailboxes (k_p2p_combine)."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = _launch(world, "tests/comm_worker.py", env={"BDF_COMBINE": combine})
    assert "COMM-OK" in out, out[-2000:]
    assert ('"combine": "peer-memory"' in out) == (combine == "p2p"), out[-500:]


@pytest.mark.gpu
def test_bench_line_with_the_driver_command_line():
    """The exact launch the driver uses for the scaling run (--steps 20 --warmup 5), at the largest N the box has."""
    n = _gpu_count()
    world = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0
    if not world:
        pytest.skip("needs >= 2 GPUs")
    import json

    out = _launch(world, "bench.py", "--gpus", str(world), "--steps", "20", "--warmup", "5", "--skip-cpu", timeout=1500)
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == world and line["value"] > 0 and line["detail"]["collectives_in_timed_region"] >= 20
    assert line["check"]["rows"] == 100_000_000 * world and line["check"]["count"] == 100_000_000 * world
    assert line["other_series"]["scaling"] == "strong" and line["other_series"]["check"]["count"] == 100_000_000


@pytest.mark.gpu
def test_single_rank_communicator(rdf, oracle):
    """A communicator of ONE rank exercises the whole on-stream combine (pack -> grouped ncclAllReduce/ncclAllGather ->
    unpack into the pinned slot) on any box: the results must equal the plain single-GPU ones bit for bit."""
    import numpy as np

    ctx = rdf.Context(0)
    rng = np.random.default_rng(5)
    lens = [5000, 0, 1, 70001]
    ints = [rdf.PrimitiveArray.from_numpy(rng.integers(-2 ** 62, 2 ** 62, n), rng.random(n) > 0.1) for n in lens]
    flts = [rdf.PrimitiveArray.from_numpy(rng.uniform(-1e3, 1e3, n), rng.random(n) > 0.1) for n in lens]
    ci, cf = rdf.Column.upload(ints, ctx=ctx), rdf.Column.upload(flts, ctx=ctx)
    plain = [ci.aggregate_all(), cf.aggregate_all()]
    plain_many = rdf.Column.aggregate_all_many([ci, cf, ci])
    col, plain_fused = cf.binary_agg(rdf.native.ADD, cf)
    col.free()
    ctx.comm_attach(rdf.Context.comm_unique_id(), 0, 1)
    info = ctx.comm_info()
    assert info["world"] == 1 and info["nccl_version"] >= 20000
    assert [ci.aggregate_all(), cf.aggregate_all()] == plain
    assert rdf.Column.aggregate_all_many([ci, cf, ci]) == plain_many
    futs = [cf.binary_agg_async(rdf.native.ADD, cf) for _ in range(4)]
    for col, fut in futs:
        assert fut.result() == plain_fused
        col.free()
    assert ci.count() == sum(c.length - c.null_count for c in ints)
    assert ctx.comm_info()["collectives"] >= 8
    with pytest.raises(rdf.ReferencePanic):
        ci.max()        # the empty chunk: reference unwrap() panic, reported through the combined flags
    ctx.comm_barrier()
    assert float(ctx.comm_all_reduce([2.5], rdf.native.MAX)[0]) == 2.5
    ctx.comm_detach()
    assert ci.aggregate_all() == plain[0]
    ci.free(); cf.free()
    ctx.close()
