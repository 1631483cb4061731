"""Multi-GPU layer: one process per GPU, chunks (RecordBatches) sharded across ranks, and ONE collective --
the all-reduce of the per-GPU partial aggregates -- where the path has a real exchange step.

The reference parallelises over the same axis with rayon (``par_iter`` over chunks in
``ScalarFunctions::add`` / ``par_multiply``, src/functions/scalar.rs:28-31, 99-102): chunks are independent,
so elementwise operators and casts need no communication at all; an aggregate is an associative fold
(src/functions/aggregate.rs:12-31,70-93), so each rank reduces its own chunks on its GPU and the partials (a few
8-byte scalars per column) are combined.

On GPUs the combine lives INSIDE libb200df.so (csrc/comm.cu): ``attach_communicator`` makes the context a rank of
an NCCL communicator and from then on every aggregate entry enqueues one grouped ``ncclAllReduce`` on the stream
that produced the partials and returns the aggregate of the whole column -- this module only ships the 128-byte
NCCL id between the ranks (through torch.distributed's rendezvous store: plumbing, no process group needed).

``combine_aggregates`` is the host-side statement of the same combine rules over ``torch.distributed`` (gloo in
the CPU tests), used where no GPU exists:
  sum    integers: wrapping 64-bit add of the two's-complement bit patterns -> bit-identical to the single-GPU result
         for every world size; floats: partial sums folded in rank order (deterministic for a given world size;
         covered by the float-sum tolerance);
  min/max on order-preserving keys;  count/rows added;  would_panic / "has chunks" OR-ed.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .arrays import NP_DTYPES, is_float

_SIGN = np.uint64(1 << 63)
_MASK64 = (1 << 64) - 1


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


def shard_row_ranges(lens: Sequence[int], rank: int, world: int, align: int = 64) -> List[Tuple[int, int, int]]:
    """Contiguous, row-balanced split of a chunked column ("contiguous blocks of chunks balanced by row
    count; a single huge chunk is split by row range on 64-row boundaries so bitmap words do not straddle GPUs").
    Returns this rank's pieces as (chunk index, first row in the chunk, rows); cuts are multiples of ``align``
    rows inside a chunk, so a piece is a zero-copy Arrow slice whose validity starts on a byte boundary."""
    total = int(sum(int(n) for n in lens))
    if world <= 1:
        return [(i, 0, int(n)) for i, n in enumerate(lens)]

    def cut(r: int) -> int:  # global row where rank r starts
        if r >= world:
            return total
        return total * r // world

    starts = []
    acc = 0
    for n in lens:
        starts.append(acc)
        acc += int(n)
    lo, hi = cut(rank), cut(rank + 1)

    def snap(g: int) -> int:  # move a global cut to an aligned row of the chunk it falls into (never past the chunk)
        if g <= 0 or g >= total:
            return max(0, min(g, total))
        for i, n in enumerate(lens):
            if starts[i] <= g < starts[i] + int(n):
                off = (g - starts[i]) // align * align
                return starts[i] + off
        return total

    lo, hi = snap(lo), snap(hi)
    out = []
    for i, n in enumerate(lens):
        b, e = max(lo, starts[i]), min(hi, starts[i] + int(n))
        if e > b:
            out.append((i, b - starts[i], e - b))
    return out


# ---------------------------------------------------------------------------------------------------------
# GPUs: the library owns the collective; only the NCCL id travels through here


def rendezvous_store():
    """The key-value store of the job's rendezvous (torchrun's agent store, or a TCPStore rank 0 hosts at
    MASTER_ADDR:MASTER_PORT).  Returns (store, rank, world)."""
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    store, rank, world = next(iter(dist.rendezvous("env://", rank=rank, world_size=world)))
    return store, rank, world


def attach_communicator(ctx, store=None, rank: Optional[int] = None, world: Optional[int] = None, key: str = "bdf/nccl_id") -> Tuple[int, int]:
    """Make ``ctx`` a rank of the job's NCCL communicator (bdf_comm_attach).  Rank 0 creates the id and publishes it
    in the store; every rank attaches.  Returns (rank, world).  With world == 1 nothing is attached."""
    if store is None:
        store, r, w = rendezvous_store()
        rank = r if rank is None else rank
        world = w if world is None else world
    if world <= 1:
        return 0, 1
    if rank == 0:
        store.set(key, type(ctx).comm_unique_id())
    uid = bytes(store.get(key))
    ctx.comm_attach(uid, rank, world)
    ctx._rendezvous_store = store   # the agent connection stays open for the life of the context
    return rank, world


# ---------------------------------------------------------------------------------------------------------
# host-side statement of the combine (CPU tests over gloo)


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


def _i64(x: int) -> int:
    x &= _MASK64
    return x - (1 << 64) if x >= (1 << 63) else x


def pack_partial(local: Dict, dtype: int) -> np.ndarray:
    """One rank's record, as 8 int64 words: [sum bits, count, rows, would_panic, has_chunks, min key, max key, any_valid]."""
    i64 = np.iinfo(np.int64)
    any_valid = 1 if (local.get("count", 0) > 0) else 0
    if is_float(dtype):
        sum_bits = int(np.array([float(local["sum"])], dtype=np.float64).view(np.int64)[0])
        kmin, kmax = i64.max, i64.min
    else:
        sum_bits = _i64(int(local["sum"]))
        kmin = _to_key(dtype, local["min"]) if any_valid and local.get("min") is not None else i64.max
        kmax = _to_key(dtype, local["max"]) if any_valid and local.get("max") is not None else i64.min
    has_chunks = int(local.get("n_chunks", 1 if local.get("rows", 0) or any_valid else 0) > 0)
    return np.array([sum_bits, int(local["count"]), int(local.get("rows", 0)), 1 if local.get("would_panic") else 0, has_chunks,
                     kmin, kmax, any_valid], dtype=np.int64)


def fold_partials(records: np.ndarray, dtype: int) -> Dict:
    """Fold the ranks' records (world x 8 int64, rank order) into the aggregates of the whole column."""
    npdt = NP_DTYPES[dtype]
    out: Dict = {"count": int(records[:, 1].sum()), "rows": int(records[:, 2].sum()), "would_panic": bool(records[:, 3].any()),
                 "n_chunks": int(records[:, 4].sum()), "min": None, "max": None}
    if is_float(dtype):
        total = 0.0
        for bits in records[:, 0]:  # fixed rank order -> deterministic
            total = total + float(np.array([bits], dtype=np.int64).view(np.float64)[0])
        out["sum"] = npdt.type(total)
        return out
    bits = 8 * npdt.itemsize
    s = 0
    for v in records[:, 0]:
        s = (s + int(v)) & _MASK64  # wrapping 64-bit add: associative, so any order gives the same bits
    s &= (1 << bits) - 1
    if npdt.kind == "i" and s >= 1 << (bits - 1):
        s -= 1 << bits
    out["sum"] = npdt.type(s)
    if records[:, 7].any():
        out["min"] = _from_key(dtype, int(records[:, 5].min()))
        out["max"] = _from_key(dtype, int(records[:, 6].max()))
    return out


def combine_aggregates(local: Dict, dtype: int, group=None, device: Optional[str] = None) -> Dict:
    """Combine the partial aggregates of this rank's shard (the dict returned by ``AggregateFunctions.all`` /
    ``Column.aggregate_all``) into the aggregates of the whole column with ONE collective (an all-gather of the
    8-word records, folded in rank order on every rank).  Every rank gets the same result."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    world = dist.get_world_size(group)
    mine = torch.from_numpy(pack_partial(local, dtype)).to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    records = np.stack([p.cpu().numpy() for p in parts])
    out = dict(local)
    out.update(fold_partials(records, dtype))
    return out


def sharded_aggregate_all(chunks_of_this_rank: Sequence, dtype: int, ctx=None, group=None) -> Dict:
    """sum/min/max/count of a column whose chunks are spread over the ranks.  On a context that is attached to a
    communicator the library call already returns the global aggregates; otherwise (CPU tests) the per-rank
    result goes through ``combine_aggregates``."""
    from .functions import AggregateFunctions

    attached = ctx is not None and ctx.comm_info()["world"] > 1
    if len(chunks_of_this_rank) or attached:
        local = AggregateFunctions.all(chunks_of_this_rank, dtype=dtype, ctx=ctx)
    else:
        local = {"sum": NP_DTYPES[dtype].type(0), "min": None, "max": None, "count": 0, "rows": 0, "would_panic": False, "n_chunks": 0}
    return local if attached else combine_aggregates(local, dtype, group=group)
