"""Multi-GPU layer: one process per GPU, chunks (RecordBatches) sharded across ranks, and ONE collective --
the all-reduce of the per-GPU partial aggregates -- where the path has a real exchange step.

The reference parallelises over the same axis with rayon (``par_iter`` over chunks in
``ScalarFunctions::add`` / ``par_multiply``, src/functions/scalar.rs:28-31, 99-102): chunks are independent,
so elementwise operators and casts need no communication at all; an aggregate is an associative fold, so
each rank reduces its own chunks on its GPU and the partials (a few 8-byte scalars per column) are combined
with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests).

Combine rules:
  sum    integers: wrapping 64-bit add (all_reduce SUM on the two's-complement bit pattern) -> bit-identical to
         the single-GPU result for every world size; floats: partial sums are all-gathered and folded in rank
         order (deterministic for a given world size; covered by the float-sum tolerance);
  min/max all_reduce MIN / MAX on an order-preserving int64 key (unsigned values get their top bit flipped);
  count  all_reduce SUM;  any_valid (min/max is None iff no rank saw a valid slot) all_reduce MAX.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from .arrays import NP_DTYPES, is_float

_SIGN = np.uint64(1 << 63)


def shard_indices(n_chunks: int, rank: int, world: int, lens: Optional[Sequence[int]] = None) -> List[int]:
    """Chunk -> rank map.  Round robin (chunk i -> rank i mod N) when lens is None, otherwise greedy
    balancing by row count (largest chunk first onto the lightest rank; ties -> lowest rank).  Every rank
    computes the same map, left and right inputs of a binary operator use the same map, so no data moves."""
    if world <= 1:
        return list(range(n_chunks))
    if lens is None:
        return [i for i in range(n_chunks) if i % world == rank]
    load = [0] * world
    owner = [0] * n_chunks
    for i in sorted(range(n_chunks), key=lambda i: (-int(lens[i]), i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += int(lens[i])
    return [i for i in range(n_chunks) if owner[i] == rank]


def shard(chunks: Sequence, rank: int, world: int, balanced: bool = False) -> List:
    lens = [c.length for c in chunks] if balanced else None
    return [chunks[i] for i in shard_indices(len(chunks), rank, world, lens)]


def _to_key(dtype: int, value) -> int:
    """Order-preserving map of a T::Native integer onto int64."""
    npdt = NP_DTYPES[dtype]
    if npdt.kind == "u":
        return int((np.uint64(value) ^ _SIGN).view(np.int64))
    return int(value)


def _from_key(dtype: int, key: int):
    npdt = NP_DTYPES[dtype]
    if npdt.kind == "u":
        return npdt.type((np.int64(key).view(np.uint64) ^ _SIGN))
    return npdt.type(key)


def combine_aggregates(local: Dict, dtype: int, group=None, device: Optional[str] = None) -> Dict:
    """All-reduce the partial aggregates of this rank's shard (the dict returned by
    ``AggregateFunctions.all`` / ``Column.aggregate_all``) into the aggregates of the whole column.
    Every rank gets the same result."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    world = dist.get_world_size(group)
    out = dict(local)
    any_valid = 1 if (local.get("count", 0) > 0) else 0
    counts = torch.tensor([int(local["count"]), int(local.get("rows", 0))], dtype=torch.int64, device=device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    out["count"], out["rows"] = int(counts[0]), int(counts[1])
    flags = torch.tensor([any_valid, 1 if local.get("would_panic") else 0], dtype=torch.int64, device=device)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    out["would_panic"] = bool(flags[1])
    if is_float(dtype):
        mine = torch.tensor([float(local["sum"])], dtype=torch.float64, device=device)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        total = 0.0
        for p in parts:  # fixed rank order -> deterministic
            total = total + float(p[0])
        out["sum"] = NP_DTYPES[dtype].type(total)
        out["min"] = out["max"] = None
        return out
    npdt = NP_DTYPES[dtype]
    bits = 8 * npdt.itemsize
    s = int(local["sum"]) & ((1 << 64) - 1)
    s = s - (1 << 64) if s >= (1 << 63) else s  # two's-complement bit pattern as int64
    t = torch.tensor([s], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)  # int64 add wraps: same bits as a 64-bit wrapping add
    total = int(t[0]) & ((1 << bits) - 1)
    if npdt.kind == "i" and total >= 1 << (bits - 1):
        total -= 1 << bits
    out["sum"] = npdt.type(total)
    i64 = np.iinfo(np.int64)
    kmin = _to_key(dtype, local["min"]) if any_valid and local.get("min") is not None else i64.max
    kmax = _to_key(dtype, local["max"]) if any_valid and local.get("max") is not None else i64.min
    tmin = torch.tensor([kmin], dtype=torch.int64, device=device)
    tmax = torch.tensor([kmax], dtype=torch.int64, device=device)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX, group=group)
    if int(flags[0]):
        out["min"], out["max"] = _from_key(dtype, int(tmin[0])), _from_key(dtype, int(tmax[0]))
    else:
        out["min"] = out["max"] = None
    return out


def sharded_aggregate_all(chunks_of_this_rank: Sequence, dtype: int, ctx=None, group=None) -> Dict:
    """sum/min/max/count of a column whose chunks are spread over the ranks: one fused reduction kernel on
    this rank's GPU, then the partial-aggregate all-reduce."""
    from .functions import AggregateFunctions

    local = AggregateFunctions.all(chunks_of_this_rank, dtype=dtype, ctx=ctx) if len(chunks_of_this_rank) else \
        {"sum": NP_DTYPES[dtype].type(0), "min": None, "max": None, "count": 0, "rows": 0, "would_panic": False}
    return combine_aggregates(local, dtype, group=group)
