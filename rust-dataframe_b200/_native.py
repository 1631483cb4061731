"""ctypes binding of libb200df.so (include/b200df.h).  Loading fails loudly when the CUDA library has not
been built: there is no CPU or PyTorch fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .arrays import NP_DTYPES, PrimitiveArray, width_of

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200df.so")

# enums (must match include/b200df.h)
ADD, SUB, MUL, DIV, ATAN2, HYPOT, LOG = range(7)
(ABS, SIN, COS, TAN, ACOS, ASIN, ATAN, CBRT, CEIL, COSH, DEGREES, EXP, EXPM1, FLOOR, LOG10, LOG2, RADIANS, ROUND, SINH,
 SQRT, TANH) = range(21)
SUM, MIN, MAX, COUNT = range(4)
OK, LENGTH_MISMATCH, DIVIDE_BY_ZERO, UNSUPPORTED, CUDA, NCCL, OOM, WOULD_PANIC, INVALID = range(9)
ASYNC = 1
K_BINARY, K_UNARY, K_CAST, K_REDUCE, K_GENERATE, K_AVG, K_COMPARE, K_FILTER, K_EXPR, K_SORT, K_TAKE, K_GROUP = range(12)
KERNEL_NAMES = ["binary", "unary", "cast", "reduce", "generate", "avg", "compare", "filter", "expr", "sort", "take", "group"]
EXPR_UNARY = 100
GT, GE, EQ, NE, LT, LE = range(6)
AND, OR, NOT = range(3)
BOOL = 10


class View(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("len", C.c_int64), ("offset", C.c_int64),
                ("null_count", C.c_int64)]


class Out(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("len", C.c_int64), ("null_count", C.c_int64),
                ("has_validity", C.c_int32)]


class Agg4(C.Structure):
    _fields_ = [("sum", C.c_uint64), ("min", C.c_uint64), ("max", C.c_uint64), ("count", C.c_int64),
                ("rows", C.c_int64), ("any_valid", C.c_int32), ("would_panic", C.c_int32), ("n_chunks", C.c_int64)]


class ExprNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


class SortKeyC(C.Structure):
    _fields_ = [("column", C.c_void_p), ("descending", C.c_int32)]


class GroupOut(C.Structure):
    _fields_ = [("sum", C.c_void_p), ("count", C.c_void_p), ("min", C.c_void_p), ("max", C.c_void_p)]


class LaunchRecord(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("dtype", C.c_int32), ("rows", C.c_int64), ("bytes", C.c_int64),
                ("ms", C.c_float)]


class ArrowError(Exception):
    """arrow::error::ArrowError as surfaced by the reference API."""


class ComputeError(ArrowError):
    pass


class DivideByZero(ArrowError):
    def __init__(self, msg="Divide by zero error"):
        super().__init__(msg)


class UnsupportedType(ArrowError):
    """A trait bound the reference enforces at compile time (T::Native: Float / Signed / Ord ...)."""


class ReferencePanic(RuntimeError):
    """The reference panics on this input (e.g. max() .unwrap() on an all-null chunk, aggregate.rs:19)."""


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "rust-dataframe_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    sig = {
        "bdf_abi_version": ([], C.c_int),
        "bdf_last_error": ([], C.c_char_p),
        "bdf_init": ([C.c_int, P(vp)], C.c_int),
        "bdf_init_multi": ([C.c_int, P(C.c_int), P(vp)], C.c_int),
        "bdf_fleet_size": ([vp], C.c_int),
        "bdf_destroy": ([vp], None),
        "bdf_synchronize": ([vp], C.c_int),
        "bdf_device_info": ([vp, P(i32), P(i32), P(i32), P(i64)], C.c_int),
        "bdf_comm_unique_id": ([P(C.c_uint8)], C.c_int),
        "bdf_comm_attach": ([vp, P(C.c_uint8), C.c_int, C.c_int], C.c_int),
        "bdf_comm_detach": ([vp], C.c_int),
        "bdf_comm_info": ([vp, P(i32), P(i32), P(i32), P(i64)], C.c_int),
        "bdf_comm_collective": ([vp, C.c_int], C.c_int),
        "bdf_comm_barrier": ([vp], C.c_int),
        "bdf_comm_set_combine": ([vp, C.c_int], C.c_int),
        "bdf_comm_get_combine": ([vp], C.c_int),
        "bdf_comm_all_reduce_f64": ([vp, C.c_int, i64, P(C.c_double)], C.c_int),
        "bdf_aggregate_all_many_dev": ([vp, i32, P(vp), P(Agg4)], C.c_int),
        "bdf_aggregate_all_many_dev_async": ([vp, i32, P(vp), P(vp)], C.c_int),
        "bdf_future_count": ([vp], C.c_int),
        "bdf_host_alloc": ([vp, C.c_size_t, P(vp)], C.c_int),
        "bdf_host_free": ([vp, vp], C.c_int),
        "bdf_host_register": ([vp, vp, C.c_size_t], C.c_int),
        "bdf_host_unregister": ([vp, vp], C.c_int),
        "bdf_binary": ([vp, C.c_int, C.c_int, i64, P(View), i64, P(View), P(Out)], C.c_int),
        "bdf_unary": ([vp, C.c_int, C.c_int, i64, P(View), P(Out)], C.c_int),
        "bdf_cast": ([vp, C.c_int, C.c_int, i64, P(View), P(Out)], C.c_int),
        "bdf_aggregate": ([vp, C.c_int, C.c_int, i64, P(View), vp, P(i32)], C.c_int),
        "bdf_aggregate_all": ([vp, C.c_int, i64, P(View), P(Agg4)], C.c_int),
        "bdf_avg": ([vp, C.c_int, i64, P(View), P(C.c_double), P(i32)], C.c_int),
        "bdf_upload": ([vp, C.c_int, i64, P(View), C.c_int, P(vp)], C.c_int),
        "bdf_upload_many": ([vp, i64, P(i32), P(i64), P(P(View)), C.c_int, P(vp)], C.c_int),
        "bdf_col_wait": ([vp, vp], C.c_int),
        "bdf_col_describe": ([vp, P(i32), P(i64), P(i64)], C.c_int),
        "bdf_col_chunk_info": ([vp, vp, i64, P(i64), P(i64), P(i32)], C.c_int),
        "bdf_binary_dev": ([vp, C.c_int, vp, vp, P(vp)], C.c_int),
        "bdf_unary_dev": ([vp, C.c_int, vp, P(vp)], C.c_int),
        "bdf_cast_dev": ([vp, C.c_int, vp, P(vp)], C.c_int),
        "bdf_aggregate_dev": ([vp, C.c_int, vp, vp, P(i32)], C.c_int),
        "bdf_aggregate_all_dev": ([vp, vp, P(Agg4)], C.c_int),
        "bdf_avg_dev": ([vp, vp, P(C.c_double), P(i32)], C.c_int),
        "bdf_expr_check": ([i32, P(i32), i32, P(ExprNode), P(i32), P(i32)], C.c_int),
        "bdf_eval_expr_dev": ([vp, i32, P(vp), i32, P(ExprNode), P(vp)], C.c_int),
        "bdf_eval_expr_agg_dev": ([vp, i32, P(vp), i32, P(ExprNode), P(vp), P(Agg4)], C.c_int),
        "bdf_eval_expr_agg_dev_async": ([vp, i32, P(vp), i32, P(ExprNode), P(vp), P(vp)], C.c_int),
        "bdf_sort_indices_dev": ([vp, i32, P(SortKeyC), P(vp)], C.c_int),
        "bdf_take_dev": ([vp, vp, vp, P(vp)], C.c_int),
        "bdf_group_aggregate_dev": ([vp, vp, i32, P(vp), P(vp), P(GroupOut), P(i64)], C.c_int),
        "bdf_compare_dev": ([vp, C.c_int, vp, vp, C.c_double, P(vp)], C.c_int),
        "bdf_boolean_dev": ([vp, C.c_int, vp, vp, P(vp)], C.c_int),
        "bdf_filter_dev": ([vp, vp, vp, P(vp)], C.c_int),
        "bdf_download": ([vp, vp, P(Out)], C.c_int),
        "bdf_download_begin": ([vp, vp, P(Out)], C.c_int),
        "bdf_download_end": ([vp, vp, P(Out)], C.c_int),
        "bdf_binary_agg_dev": ([vp, C.c_int, vp, vp, P(vp), P(Agg4)], C.c_int),
        "bdf_binary_agg_dev_async": ([vp, C.c_int, vp, vp, P(vp), P(vp)], C.c_int),
        "bdf_aggregate_all_dev_async": ([vp, vp, P(vp)], C.c_int),
        "bdf_future_wait": ([vp, vp, P(Agg4)], C.c_int),
        "bdf_col_free": ([vp, vp], None),
        "bdf_profile_enable": ([vp, C.c_int], C.c_int),
        "bdf_profile_read": ([vp, P(LaunchRecord), i64, P(i64)], C.c_int),
        "bdf_launch_count": ([vp], i64),
        "bdf_timer_start": ([vp], C.c_int),
        "bdf_timer_stop": ([vp, P(C.c_float)], C.c_int),
        "bdf_flush_l2": ([vp, C.c_size_t], C.c_int),
        "bdf_ipc_open": ([C.c_char_p, P(vp)], C.c_int),
        "bdf_ipc_close": ([vp], None),
        "bdf_ipc_describe": ([vp, P(i32), P(i64), P(i64)], C.c_int),
        "bdf_ipc_column": ([vp, i32, P(C.c_char_p), P(i32), P(i32)], C.c_int),
        "bdf_ipc_batch_rows": ([vp, i64, P(i64)], C.c_int),
        "bdf_ipc_view": ([vp, i64, i32, P(View)], C.c_int),
        "bdf_ipc_read": ([vp, vp, i32, P(i32), C.c_int, P(vp)], C.c_int),
        "bdf_ipc_read_batches": ([vp, vp, i32, P(i32), i64, P(i64), C.c_int, P(vp)], C.c_int),
        "bdf_ipc_write_host": ([C.c_char_p, i32, P(C.c_char_p), P(i32), i64, P(P(View))], C.c_int),
        "bdf_ipc_write": ([vp, C.c_char_p, i32, P(C.c_char_p), P(vp)], C.c_int),
        "bdf_generate": ([vp, C.c_int, C.c_int, C.c_double, C.c_double, u64, u64, i64, P(i64), i64, C.c_uint32, P(vp)],
                         C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = args
        fn.restype = res
    if L.bdf_abi_version() != 2:
        raise RuntimeError("libb200df.so ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "bdf_abi_version", "bdf_last_error", "bdf_init", "bdf_init_multi", "bdf_fleet_size", "bdf_destroy", "bdf_synchronize", "bdf_device_info",
    "bdf_comm_unique_id", "bdf_comm_attach", "bdf_comm_detach", "bdf_comm_info", "bdf_comm_collective", "bdf_comm_barrier", "bdf_comm_set_combine", "bdf_comm_get_combine",
    "bdf_comm_all_reduce_f64", "bdf_aggregate_all_many_dev", "bdf_aggregate_all_many_dev_async", "bdf_future_count",
    "bdf_host_alloc", "bdf_host_free", "bdf_host_register", "bdf_host_unregister", "bdf_binary", "bdf_unary", "bdf_cast",
    "bdf_aggregate", "bdf_aggregate_all", "bdf_avg", "bdf_upload", "bdf_upload_many", "bdf_col_wait", "bdf_col_describe",
    "bdf_col_chunk_info", "bdf_binary_dev", "bdf_unary_dev", "bdf_cast_dev", "bdf_aggregate_dev",
    "bdf_aggregate_all_dev", "bdf_avg_dev", "bdf_expr_check", "bdf_eval_expr_dev", "bdf_eval_expr_agg_dev", "bdf_eval_expr_agg_dev_async", "bdf_sort_indices_dev", "bdf_take_dev", "bdf_group_aggregate_dev", "bdf_compare_dev", "bdf_boolean_dev", "bdf_filter_dev", "bdf_download", "bdf_download_begin", "bdf_download_end",
    "bdf_binary_agg_dev", "bdf_binary_agg_dev_async", "bdf_aggregate_all_dev_async", "bdf_future_wait", "bdf_col_free", "bdf_profile_enable", "bdf_profile_read",
    "bdf_launch_count", "bdf_timer_start", "bdf_timer_stop", "bdf_flush_l2", "bdf_generate",
    "bdf_ipc_open", "bdf_ipc_close", "bdf_ipc_describe", "bdf_ipc_column", "bdf_ipc_batch_rows", "bdf_ipc_view", "bdf_ipc_read", "bdf_ipc_read_batches",
    "bdf_ipc_write_host", "bdf_ipc_write",
]


def raise_for_status(st: int) -> None:
    if st == OK:
        return
    msg = (lib().bdf_last_error() or b"").decode("utf-8", "replace")
    if st == LENGTH_MISMATCH:
        raise ComputeError("Cannot perform math operation on arrays of different length")
    if st == DIVIDE_BY_ZERO:
        raise DivideByZero()
    if st == UNSUPPORTED:
        raise UnsupportedType(msg)
    if st == WOULD_PANIC:
        raise ReferencePanic(msg)
    if st == OOM:
        raise MemoryError(msg)
    raise ComputeError(f"[bdf status {st}] {msg}")


def make_views(chunks: Sequence[PrimitiveArray]):
    arr = (View * max(len(chunks), 1))()
    for i, c in enumerate(chunks):
        arr[i].values = c.values.ctypes.data if c.values.size else None
        arr[i].validity = c.validity.ctypes.data if c.validity is not None else None
        arr[i].len = c.length
        arr[i].offset = c.offset
        arr[i].null_count = c.null_count
    return arr


class PinnedBuffer:
    """cudaHostAlloc'ed bytes exposed as numpy arrays (freed with the object)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self._ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        raise_for_status(lib().bdf_host_alloc(ctx.handle, max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value
        self._raw = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)

    def array(self, dtype, count: int, byte_offset: int = 0) -> np.ndarray:
        a = np.frombuffer(self._raw, dtype=dtype, count=count, offset=byte_offset)
        return a

    def close(self):
        if self.ptr and self._ctx.handle:
            lib().bdf_host_free(self._ctx.handle, C.c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One bdf_ctx = one GPU.  One process drives one GPU (torchrun gives each rank its LOCAL_RANK)."""

    def __init__(self, device: Optional[int] = None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        raise_for_status(lib().bdf_init(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    @classmethod
    def multi(cls, n_gpus: int = 0, devices: Optional[Sequence[int]] = None) -> "Context":
        """ONE context over several GPUs of the box (bdf_init_multi; 0 = all visible): the library shards every call by
        row range, each GPU works over its own PCIe link, aggregates are combined by the grouped ncclAllReduce."""
        self = cls.__new__(cls)
        h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        raise_for_status(lib().bdf_init_multi(int(n_gpus if not devices else len(devices)), devs, C.byref(h)))
        self.handle = h
        self.device = int(devices[0]) if devices else 0
        return self

    @property
    def n_gpus(self) -> int:
        return int(lib().bdf_fleet_size(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            lib().bdf_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- misc --
    def synchronize(self):
        raise_for_status(lib().bdf_synchronize(self.handle))

    def device_info(self):
        sm, ma, mi, hbm = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        raise_for_status(lib().bdf_device_info(self.handle, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hbm)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hbm.value}

    def pinned(self, nbytes: int) -> PinnedBuffer:
        return PinnedBuffer(self, nbytes)

    # -- multi-GPU: this context as a rank of a communicator (include/b200df.h "multi-GPU") --
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        raise_for_status(lib().bdf_comm_unique_id(buf))
        return bytes(buf)

    def comm_attach(self, unique_id: bytes, rank: int, world: int):
        """ncclCommInitRank: afterwards every aggregate entry is collective and returns the aggregate of the whole
        (sharded) column on every rank."""
        buf = (C.c_uint8 * 128)(*unique_id)
        raise_for_status(lib().bdf_comm_attach(self.handle, buf, int(rank), int(world)))

    def comm_detach(self):
        raise_for_status(lib().bdf_comm_detach(self.handle))

    def comm_info(self) -> dict:
        r, w, v, n = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        raise_for_status(lib().bdf_comm_info(self.handle, C.byref(r), C.byref(w), C.byref(v), C.byref(n)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value, "collectives": n.value}

    def comm_collective(self, on: bool):
        raise_for_status(lib().bdf_comm_collective(self.handle, 1 if on else 0))

    def comm_barrier(self):
        raise_for_status(lib().bdf_comm_barrier(self.handle))

    def comm_set_combine(self, peer_memory: bool):
        """True: NVLink peer-memory mailboxes (one kernel of the library); False: the grouped ncclAllReduce."""
        raise_for_status(lib().bdf_comm_set_combine(self.handle, 1 if peer_memory else 0))

    def comm_get_combine(self) -> str:
        return "peer-memory" if lib().bdf_comm_get_combine(self.handle) else "nccl"

    def comm_all_reduce(self, values, op: int = SUM):
        """Blocking all-reduce of a few float64 values (timing max-over-ranks and the like)."""
        arr = np.ascontiguousarray(np.atleast_1d(np.asarray(values, dtype=np.float64))).copy()
        raise_for_status(lib().bdf_comm_all_reduce_f64(self.handle, op, arr.shape[0], arr.ctypes.data_as(C.POINTER(C.c_double))))
        return arr

    def pinned_array(self, dtype: int, values: np.ndarray, mask: Optional[np.ndarray] = None) -> PrimitiveArray:
        """Copy a numpy array (and optional bool mask, True = valid) into pinned host memory."""
        n = int(values.shape[0])
        w = width_of(dtype)
        vbytes = (n * w + 63) // 64 * 64
        bbytes = ((n + 7) // 8 + 63) // 64 * 64 if mask is not None else 0
        buf = self.pinned(vbytes + bbytes)
        v = buf.array(NP_DTYPES[dtype], n)
        v[:] = values
        validity, nulls = None, 0
        if mask is not None:
            validity = buf.array(np.uint8, (n + 7) // 8, vbytes)
            validity[:] = np.packbits(np.asarray(mask, dtype=bool), bitorder="little")
            nulls = int(n - np.count_nonzero(mask))
        return PrimitiveArray(dtype, v, validity, 0, n, nulls, keepalive=buf)

    def profile_enable(self, on: bool = True):
        raise_for_status(lib().bdf_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self, cap: int = 65536) -> List[dict]:
        buf = (LaunchRecord * cap)()
        n = C.c_int64(0)
        raise_for_status(lib().bdf_profile_read(self.handle, buf, cap, C.byref(n)))
        return [{"kernel": KERNEL_NAMES[buf[i].kernel], "dtype": buf[i].dtype, "rows": buf[i].rows,
                 "bytes": buf[i].bytes, "ms": buf[i].ms} for i in range(n.value)]

    def launch_count(self) -> int:
        return int(lib().bdf_launch_count(self.handle))

    def timer_start(self):
        raise_for_status(lib().bdf_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        raise_for_status(lib().bdf_timer_stop(self.handle, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self, nbytes: int = 256 << 20):
        raise_for_status(lib().bdf_flush_l2(self.handle, nbytes))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide lazy singleton (what the Rust shim does with a OnceCell)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def alloc_outputs(dtype: int, lens: Sequence[int], ctx: Optional[Context] = None, pinned: bool = False):
    """Caller-side output buffers (the Rust shim uses MutableBuffer::new): returns (Out[], [(values, bitmap)])."""
    outs = (Out * max(len(lens), 1))()
    bufs = []
    if dtype == BOOL:  # bit-packed values
        for i, n in enumerate(lens):
            v = np.zeros((n + 7) // 8, dtype=np.uint8)
            b = np.zeros((n + 7) // 8, dtype=np.uint8)
            outs[i].values = v.ctypes.data if v.size else None
            outs[i].validity = b.ctypes.data if b.size else None
            outs[i].len = n
            bufs.append((v, b, None))
        return outs, bufs
    for i, n in enumerate(lens):
        if pinned:
            w = width_of(dtype)
            vbytes = (n * w + 63) // 64 * 64
            pb = ctx.pinned(vbytes + ((n + 7) // 8 + 63) // 64 * 64)
            v = pb.array(NP_DTYPES[dtype], n)
            b = pb.array(np.uint8, (n + 7) // 8, vbytes)
            keep = pb
        else:
            v = np.empty(n, dtype=NP_DTYPES[dtype])
            b = np.zeros((n + 7) // 8, dtype=np.uint8)
            keep = None
        outs[i].values = v.ctypes.data if n else None
        outs[i].validity = b.ctypes.data if b.size else None
        outs[i].len = n
        bufs.append((v, b, keep))
    return outs, bufs


def collect_outputs(dtype: int, outs, bufs) -> List[PrimitiveArray]:
    res = []
    if dtype == BOOL:
        from .arrays import BooleanArray

        for i, (v, b, keep) in enumerate(bufs):
            has_v = bool(outs[i].has_validity)
            res.append(BooleanArray(v, b if has_v else None, 0, int(outs[i].len), int(outs[i].null_count) if has_v else 0))
        return res
    for i, (v, b, keep) in enumerate(bufs):
        has_v = bool(outs[i].has_validity)
        res.append(PrimitiveArray(dtype, v, b if has_v else None, 0, int(outs[i].len),
                                  int(outs[i].null_count) if has_v else 0, keepalive=keep))
    return res
