"""Python mirror of the reference's operator interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference (nevi-me/rust-dataframe @ a8310afd):

* ``ScalarFunctions``     src/functions/scalar.rs:14-497   (add, subtract, multiply, par_multiply, divide,
                          abs, sin, cos, tan, acos, ...; each takes the column as a list of chunks)
* ``AggregateFunctions``  src/functions/aggregate.rs:9-103 (sum, min, max, count, avg)
* ``cast``                the arrow::compute::cast call of Function::Cast, src/evaluation.rs:296-315

Every function makes exactly one call into libb200df.so (CUDA); nothing here computes on the CPU.
``Column`` is the device-resident handle used to chain operators without leaving HBM.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from .arrays import F64, NP_DTYPES, PrimitiveArray, is_float, width_of

Chunks = Sequence[PrimitiveArray]


def _dtype_of(chunks: Chunks, other: Optional[Chunks] = None) -> int:
    for c in list(chunks) + list(other or []):
        return c.dtype
    raise ValueError("empty column: the element type cannot be inferred; pass at least one chunk")


def _check_same_type(chunks: Chunks, dtype: int) -> None:
    for c in chunks:
        if c.dtype != dtype:
            # col_to_prim_arrays::<T> downcast .unwrap() panics on a type mismatch (src/table.rs:120)
            raise N.ReferencePanic("chunk type does not match the column type (downcast_ref::<PrimitiveArray<T>>().unwrap())")


def _binary(op: int, left: Chunks, right: Chunks, ctx: Optional[N.Context], pinned_out: bool) -> List[PrimitiveArray]:
    ctx = ctx or N.default_context()
    n = min(len(left), len(right))  # zip()
    if n == 0:
        return []
    dtype = _dtype_of(left, right)
    _check_same_type(left[:n], dtype)
    _check_same_type(right[:n], dtype)
    outs, bufs = N.alloc_outputs(dtype, [left[i].length for i in range(n)], ctx, pinned_out)
    st = N.lib().bdf_binary(ctx.handle, op, dtype, len(left), N.make_views(left), len(right), N.make_views(right), outs)
    N.raise_for_status(st)
    return N.collect_outputs(dtype, outs, bufs)


def _unary(op: int, array: Chunks, ctx: Optional[N.Context], pinned_out: bool) -> List[PrimitiveArray]:
    ctx = ctx or N.default_context()
    if len(array) == 0:
        return []
    dtype = _dtype_of(array)
    _check_same_type(array, dtype)
    outs, bufs = N.alloc_outputs(dtype, [c.length for c in array], ctx, pinned_out)
    N.raise_for_status(N.lib().bdf_unary(ctx.handle, op, dtype, len(array), N.make_views(array), outs))
    return N.collect_outputs(dtype, outs, bufs)


class ScalarFunctions:
    """src/functions/scalar.rs: static functions over ``Vec<&PrimitiveArray<T>>`` (one entry per chunk)."""

    # -- binary arithmetic: arrow::compute::{add,subtract,multiply,divide} per chunk (scalar.rs:16-103) --
    @staticmethod
    def add(left: Chunks, right: Chunks, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
        return _binary(N.ADD, left, right, ctx, pinned_out)

    @staticmethod
    def subtract(left: Chunks, right: Chunks, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
        return _binary(N.SUB, left, right, ctx, pinned_out)

    @staticmethod
    def multiply(left: Chunks, right: Chunks, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
        return _binary(N.MUL, left, right, ctx, pinned_out)

    @staticmethod
    def par_multiply(left: Chunks, right: Chunks, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
        # the reference's rayon variant (scalar.rs:87-103); on the GPU every chunk is already in one launch
        return _binary(N.MUL, left, right, ctx, pinned_out)

    @staticmethod
    def divide(left: Chunks, right: Chunks, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
        return _binary(N.DIV, left, right, ctx, pinned_out)

    # -- math_op binaries on single arrays (scalar.rs:148,274,291) --
    @staticmethod
    def atan2(a: PrimitiveArray, b: PrimitiveArray, ctx=None) -> PrimitiveArray:
        return _binary(N.ATAN2, [a], [b], ctx, False)[0]

    @staticmethod
    def hypot(a: PrimitiveArray, b: PrimitiveArray, ctx=None) -> PrimitiveArray:
        return _binary(N.HYPOT, [a], [b], ctx, False)[0]

    @staticmethod
    def log(a: PrimitiveArray, b: PrimitiveArray, ctx=None) -> PrimitiveArray:
        return _binary(N.LOG, [a], [b], ctx, False)[0]


def _add_unary(name: str, op: int) -> None:
    def fn(array: Chunks, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
        return _unary(op, array, ctx, pinned_out)

    fn.__name__ = name
    fn.__doc__ = f"ScalarFunctions::{name}: scalar_op over every chunk (src/functions/scalar.rs:525-540)."
    setattr(ScalarFunctions, name, staticmethod(fn))


for _name, _op in [("abs", N.ABS), ("sin", N.SIN), ("cos", N.COS), ("tan", N.TAN), ("acos", N.ACOS), ("asin", N.ASIN),
                   ("atan", N.ATAN), ("cbrt", N.CBRT), ("ceil", N.CEIL), ("cosh", N.COSH), ("degrees", N.DEGREES),
                   ("exp", N.EXP), ("expm1", N.EXPM1), ("floor", N.FLOOR), ("log10", N.LOG10), ("log2", N.LOG2),
                   ("radians", N.RADIANS), ("round", N.ROUND), ("sinh", N.SINH), ("sqrt", N.SQRT), ("tanh", N.TANH)]:
    _add_unary(_name, _op)


def cast(arrays: Chunks, to_dtype: int, ctx=None, pinned_out=False) -> List[PrimitiveArray]:
    """arrow::compute::cast over every chunk of a column, as Function::Cast does (src/evaluation.rs:296-315)."""
    ctx = ctx or N.default_context()
    if len(arrays) == 0:
        return []
    from_dtype = _dtype_of(arrays)
    _check_same_type(arrays, from_dtype)
    outs, bufs = N.alloc_outputs(to_dtype, [c.length for c in arrays], ctx, pinned_out)
    N.raise_for_status(N.lib().bdf_cast(ctx.handle, from_dtype, to_dtype, len(arrays), N.make_views(arrays), outs))
    return N.collect_outputs(to_dtype, outs, bufs)


def _scalar_from_bits(dtype: int, bits: int):
    return np.array([bits], dtype=np.uint64).view(NP_DTYPES[dtype])[0]


class AggregateFunctions:
    """src/functions/aggregate.rs.  ``None`` mirrors Rust's ``Option::None``.

    ``min`` returns the true minimum (what arrow's compute::min does).  The reference's ``min`` body is a
    copy of ``max`` (aggregate.rs:22-31); ``min_as_written`` reproduces that behaviour for completeness."""

    @staticmethod
    def _agg(op: int, arrays: Chunks, dtype: Optional[int], ctx):
        ctx = ctx or N.default_context()
        if dtype is None:
            dtype = _dtype_of(arrays) if len(arrays) else N_I64_FALLBACK
        _check_same_type(arrays, dtype)
        out = np.zeros(1, dtype=np.int64 if op == N.COUNT else NP_DTYPES[dtype])
        some = C.c_int32(0)
        st = N.lib().bdf_aggregate(ctx.handle, op, dtype, len(arrays), N.make_views(arrays), out.ctypes.data, C.byref(some))
        N.raise_for_status(st)
        return out[0] if some.value else None

    @staticmethod
    def sum(arrays: Chunks, dtype: Optional[int] = None, ctx=None):
        return AggregateFunctions._agg(N.SUM, arrays, dtype, ctx)

    @staticmethod
    def max(arrays: Chunks, dtype: Optional[int] = None, ctx=None):
        return AggregateFunctions._agg(N.MAX, arrays, dtype, ctx)

    @staticmethod
    def min(arrays: Chunks, dtype: Optional[int] = None, ctx=None):
        return AggregateFunctions._agg(N.MIN, arrays, dtype, ctx)

    @staticmethod
    def min_as_written(arrays: Chunks, dtype: Optional[int] = None, ctx=None):
        return AggregateFunctions._agg(N.MAX, arrays, dtype, ctx)

    @staticmethod
    def count(arrays: Chunks, dtype: Optional[int] = None, ctx=None):
        v = AggregateFunctions._agg(N.COUNT, arrays, dtype, ctx)
        return None if v is None else int(v)

    @staticmethod
    def avg(arrays: Chunks, dtype: Optional[int] = None, ctx=None):
        ctx = ctx or N.default_context()
        if dtype is None:
            dtype = _dtype_of(arrays) if len(arrays) else N_I64_FALLBACK
        out, some = C.c_double(0), C.c_int32(0)
        st = N.lib().bdf_avg(ctx.handle, dtype, len(arrays), N.make_views(arrays), C.byref(out), C.byref(some))
        N.raise_for_status(st)
        return out.value if some.value else None

    @staticmethod
    def all(arrays: Chunks, dtype: Optional[int] = None, ctx=None) -> dict:
        """sum, min, max and count from ONE pass over the data (fused 4-in-1 reduction)."""
        ctx = ctx or N.default_context()
        if dtype is None:
            dtype = _dtype_of(arrays)
        a = N.Agg4()
        N.raise_for_status(N.lib().bdf_aggregate_all(ctx.handle, dtype, len(arrays), N.make_views(arrays), C.byref(a)))
        return _agg4_to_dict(dtype, a)


N_I64_FALLBACK = 3  # Int64: only used to type an empty Vec (results do not depend on it)


def _agg4_to_dict(dtype: int, a: "N.Agg4") -> dict:
    res = {"sum": _scalar_from_bits(dtype, a.sum), "count": int(a.count), "rows": int(a.rows),
           "would_panic": bool(a.would_panic), "n_chunks": int(a.n_chunks), "min": None, "max": None}
    if not is_float(dtype) and a.any_valid:
        res["min"] = _scalar_from_bits(dtype, a.min)
        res["max"] = _scalar_from_bits(dtype, a.max)
    return res


# ---------------------------------------------------------------------------------------------------------
# device-resident columns


class Column:
    """A column living in HBM (list of Arrow chunks).  Operators return new Columns; nothing crosses PCIe
    until ``download``.  Mirrors how Evaluate::calculate (src/evaluation.rs:97-323) materialises every
    Calculation as a new column of the frame."""

    def __init__(self, ctx: N.Context, handle: C.c_void_p):
        self.ctx = ctx
        self.handle = handle

    # -- construction --
    @classmethod
    def upload(cls, arrays: Chunks, ctx: Optional[N.Context] = None, asynchronous: bool = False,
               dtype: Optional[int] = None) -> "Column":
        ctx = ctx or N.default_context()
        if dtype is None:
            dtype = _dtype_of(arrays)
        _check_same_type(arrays, dtype)
        h = C.c_void_p()
        st = N.lib().bdf_upload(ctx.handle, dtype, len(arrays), N.make_views(arrays), N.ASYNC if asynchronous else 0, C.byref(h))
        N.raise_for_status(st)
        col = cls(ctx, h)
        col._hold = arrays if asynchronous else None  # keep host buffers alive while copies are in flight
        return col

    @classmethod
    def upload_many(cls, columns: Sequence[Chunks], ctx: Optional[N.Context] = None, asynchronous: bool = False) -> List["Column"]:
        """Upload several columns with their chunks interleaved (a0,b0,a1,b1,...) -- one Vec<RecordBatch>."""
        ctx = ctx or N.default_context()
        k = len(columns)
        dtypes = (C.c_int32 * k)(*[_dtype_of(col) for col in columns])
        counts = (C.c_int64 * k)(*[len(col) for col in columns])
        views = [N.make_views(col) for col in columns]
        ptrs = (C.POINTER(N.View) * k)(*[C.cast(v, C.POINTER(N.View)) for v in views])
        outs = (C.c_void_p * k)()
        st = N.lib().bdf_upload_many(ctx.handle, k, dtypes, counts, ptrs, N.ASYNC if asynchronous else 0, outs)
        N.raise_for_status(st)
        cols = []
        for i in range(k):
            col = cls(ctx, C.c_void_p(outs[i]))
            col._hold = (columns[i], views) if asynchronous else None
            cols.append(col)
        return cols

    @classmethod
    def generate(cls, dtype: int, chunk_lens: Sequence[int], kind: int = 0, lo: float = 0.0, hi: float = 1.0,
                 seed: int = 20260924, col_id: int = 0, row0: int = 0, null_mod: int = 0,
                 ctx: Optional[N.Context] = None) -> "Column":
        ctx = ctx or N.default_context()
        lens = (C.c_int64 * max(len(chunk_lens), 1))(*chunk_lens)
        h = C.c_void_p()
        st = N.lib().bdf_generate(ctx.handle, dtype, kind, lo, hi, seed, col_id, len(chunk_lens), lens, row0, null_mod, C.byref(h))
        N.raise_for_status(st)
        return cls(ctx, h)

    # -- metadata --
    def describe(self):
        dt, n, total = C.c_int32(), C.c_int64(), C.c_int64()
        N.raise_for_status(N.lib().bdf_col_describe(self.handle, C.byref(dt), C.byref(n), C.byref(total)))
        return dt.value, n.value, total.value

    @property
    def dtype(self) -> int:
        return self.describe()[0]

    @property
    def n_chunks(self) -> int:
        return self.describe()[1]

    def __len__(self) -> int:
        return self.describe()[2]

    def chunk_info(self, i: int):
        ln, nc, hv = C.c_int64(), C.c_int64(), C.c_int32()
        N.raise_for_status(N.lib().bdf_col_chunk_info(self.ctx.handle, self.handle, i, C.byref(ln), C.byref(nc), C.byref(hv)))
        return {"len": ln.value, "null_count": nc.value, "has_validity": bool(hv.value)}

    def wait(self):
        N.raise_for_status(N.lib().bdf_col_wait(self.ctx.handle, self.handle))

    # -- operators --
    def _bin(self, op: int, other: "Column") -> "Column":
        h = C.c_void_p()
        N.raise_for_status(N.lib().bdf_binary_dev(self.ctx.handle, op, self.handle, other.handle, C.byref(h)))
        return Column(self.ctx, h)

    def add(self, o): return self._bin(N.ADD, o)
    def subtract(self, o): return self._bin(N.SUB, o)
    def multiply(self, o): return self._bin(N.MUL, o)
    def divide(self, o): return self._bin(N.DIV, o)
    def atan2(self, o): return self._bin(N.ATAN2, o)
    def hypot(self, o): return self._bin(N.HYPOT, o)
    def log(self, o): return self._bin(N.LOG, o)

    # -- fused operator + aggregate of the result (one pass) --
    def binary_agg(self, op: int, other: "Column"):
        """(self op other) materialised as a Column AND sum/min/max/count of it, from one kernel pass."""
        h, a = C.c_void_p(), N.Agg4()
        N.raise_for_status(N.lib().bdf_binary_agg_dev(self.ctx.handle, op, self.handle, other.handle, C.byref(h), C.byref(a)))
        return Column(self.ctx, h), _agg4_to_dict(self.dtype, a)

    def add_agg(self, o): return self.binary_agg(N.ADD, o)

    def binary_agg_async(self, op: int, other: "Column"):
        """Like binary_agg but returns (Column, AggFuture) without waiting for the GPU."""
        h, f = C.c_void_p(), C.c_void_p()
        N.raise_for_status(N.lib().bdf_binary_agg_dev_async(self.ctx.handle, op, self.handle, other.handle, C.byref(h), C.byref(f)))
        return Column(self.ctx, h), AggFuture(self.ctx, f, self.dtype)

    def take(self, indices: "Column") -> "Column":
        """Column::take (src/table.rs:218-241): values at the UInt32 row numbers, as ONE chunk; null index -> null slot."""
        h = C.c_void_p()
        N.raise_for_status(N.lib().bdf_take_dev(self.ctx.handle, self.handle, indices.handle, C.byref(h)))
        return Column(self.ctx, h)

    def aggregate_all_async(self) -> "AggFuture":
        f = C.c_void_p()
        N.raise_for_status(N.lib().bdf_aggregate_all_dev_async(self.ctx.handle, self.handle, C.byref(f)))
        return AggFuture(self.ctx, f, self.dtype)

    # -- N2: BooleanFilter comparisons, boolean kernels, filter --
    def compare(self, op: int, other) -> "Column":
        """BooleanFilter::{Gt,Ge,Eq,Ne,Lt,Le}: both sides cast to Float64, compared; `other` is a Column or a scalar
        (BooleanInput::Scalar).  Returns a boolean mask column."""
        h = C.c_void_p()
        if isinstance(other, Column):
            st = N.lib().bdf_compare_dev(self.ctx.handle, op, self.handle, other.handle, 0.0, C.byref(h))
        else:
            st = N.lib().bdf_compare_dev(self.ctx.handle, op, self.handle, None, float(other), C.byref(h))
        N.raise_for_status(st)
        return Column(self.ctx, h)

    def gt(self, o): return self.compare(N.GT, o)
    def ge(self, o): return self.compare(N.GE, o)
    def eq(self, o): return self.compare(N.EQ, o)
    def ne(self, o): return self.compare(N.NE, o)
    def lt(self, o): return self.compare(N.LT, o)
    def le(self, o): return self.compare(N.LE, o)

    def _boolean(self, op: int, other: Optional["Column"]) -> "Column":
        h = C.c_void_p()
        N.raise_for_status(N.lib().bdf_boolean_dev(self.ctx.handle, op, self.handle, other.handle if other is not None else None, C.byref(h)))
        return Column(self.ctx, h)

    def logical_and(self, o): return self._boolean(N.AND, o)
    def logical_or(self, o): return self._boolean(N.OR, o)
    def logical_not(self): return self._boolean(N.NOT, None)

    def filter(self, mask: "Column") -> "Column":
        """ChunkedArray::filter: keep the slots where the boolean mask column is valid and true."""
        h = C.c_void_p()
        N.raise_for_status(N.lib().bdf_filter_dev(self.ctx.handle, self.handle, mask.handle, C.byref(h)))
        return Column(self.ctx, h)

    def unary(self, op: int) -> "Column":
        h = C.c_void_p()
        N.raise_for_status(N.lib().bdf_unary_dev(self.ctx.handle, op, self.handle, C.byref(h)))
        return Column(self.ctx, h)

    def sin(self): return self.unary(N.SIN)
    def cos(self): return self.unary(N.COS)
    def tan(self): return self.unary(N.TAN)
    def abs(self): return self.unary(N.ABS)

    def cast(self, to_dtype: int) -> "Column":
        h = C.c_void_p()
        N.raise_for_status(N.lib().bdf_cast_dev(self.ctx.handle, to_dtype, self.handle, C.byref(h)))
        return Column(self.ctx, h)

    def aggregate(self, op: int):
        dtype = self.dtype
        out = np.zeros(1, dtype=np.int64 if (op == N.COUNT or dtype == N.BOOL) else NP_DTYPES[dtype])
        some = C.c_int32(0)
        N.raise_for_status(N.lib().bdf_aggregate_dev(self.ctx.handle, op, self.handle, out.ctypes.data, C.byref(some)))
        return out[0] if some.value else None

    def sum(self): return self.aggregate(N.SUM)
    def min(self): return self.aggregate(N.MIN)
    def max(self): return self.aggregate(N.MAX)
    def count(self): return int(self.aggregate(N.COUNT))

    def aggregate_all(self) -> dict:
        a = N.Agg4()
        N.raise_for_status(N.lib().bdf_aggregate_all_dev(self.ctx.handle, self.handle, C.byref(a)))
        return _agg4_to_dict(self.dtype, a)

    def avg(self):
        out, some = C.c_double(0), C.c_int32(0)
        N.raise_for_status(N.lib().bdf_avg_dev(self.ctx.handle, self.handle, C.byref(out), C.byref(some)))
        return out.value if some.value else None

    @staticmethod
    def aggregate_all_many(columns: Sequence["Column"], asynchronous: bool = False):
        """sum/min/max/count of several columns in ONE call: one reduction per column, one host wait and (on a rank of
        a communicator) one grouped collective for all of them.  Returns a list of dicts (or an AggFuture of one)."""
        ctx = columns[0].ctx
        k = len(columns)
        cols = (C.c_void_p * k)(*[c.handle for c in columns])
        dtypes = [c.dtype for c in columns]
        if asynchronous:
            f = C.c_void_p()
            N.raise_for_status(N.lib().bdf_aggregate_all_many_dev_async(ctx.handle, k, cols, C.byref(f)))
            return AggFuture(ctx, f, dtypes)
        out = (N.Agg4 * k)()
        N.raise_for_status(N.lib().bdf_aggregate_all_many_dev(ctx.handle, k, cols, out))
        return [_agg4_to_dict(dtypes[i], out[i]) for i in range(k)]

    # -- back to the host --
    def download(self, pinned: bool = False, into=None) -> List[PrimitiveArray]:
        dtype, n, _ = self.describe()
        if into is not None:
            outs, bufs = into
        else:
            lens = [self.chunk_len(i) for i in range(n)]
            outs, bufs = N.alloc_outputs(dtype, lens, self.ctx, pinned)
        if dtype != N.BOOL:
            for i in range(n):
                outs[i].len = bufs[i][0].shape[0]
        N.raise_for_status(N.lib().bdf_download(self.ctx.handle, self.handle, outs))
        return N.collect_outputs(dtype, outs, bufs)

    def download_begin(self, into):
        """Enqueue the device->host copies into `into` = alloc_outputs(...); pair with download_end(into)."""
        outs, bufs = into
        for i in range(len(bufs)):
            outs[i].len = bufs[i][0].shape[0]
        N.raise_for_status(N.lib().bdf_download_begin(self.ctx.handle, self.handle, outs))

    def download_end(self, into) -> List[PrimitiveArray]:
        outs, bufs = into
        N.raise_for_status(N.lib().bdf_download_end(self.ctx.handle, self.handle, outs))
        return N.collect_outputs(self.dtype, outs, bufs)

    def chunk_len(self, i: int) -> int:
        ln = C.c_int64()
        N.raise_for_status(N.lib().bdf_col_chunk_info(self.ctx.handle, self.handle, i, C.byref(ln), None, None))
        return ln.value

    def free(self):
        if self.handle is not None and self.ctx.handle:
            N.lib().bdf_col_free(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def sort_indices(criteria: Sequence[tuple]) -> "Column":
    """compute::lexsort_to_indices as DataFrame::sort calls it (src/dataframe.rs:194-213).  criteria: [(Column, descending)];
    stable, nulls last.  Returns a one-chunk UInt32 Column of row numbers."""
    if not criteria:
        raise N.ComputeError("Sort criteria cannot be empty")
    ctx = criteria[0][0].ctx
    arr = (N.SortKeyC * len(criteria))()
    for i, (col, desc) in enumerate(criteria):
        arr[i].column, arr[i].descending = col.handle.value if isinstance(col.handle, C.c_void_p) else col.handle, 1 if desc else 0
    h = C.c_void_p()
    N.raise_for_status(N.lib().bdf_sort_indices_dev(ctx.handle, len(criteria), arr, C.byref(h)))
    return Column(ctx, h)


def sort_columns(criteria: Sequence[tuple], columns: Sequence["Column"]) -> List["Column"]:
    """DataFrame::sort: the index sort, then DataFrame::sort_by_indices (Column::take of every column, one chunk each)."""
    idx = sort_indices(criteria)
    return [c.take(idx) for c in columns]


def group_aggregate(key: "Column", values: Sequence["Column"]):
    """Group-by aggregate (`GroupAggregate`, unimplemented in the reference: src/evaluation.rs:73): returns
    ``(keys Column, [ {"sum": Column, "count": Column, "min": Column | None, "max": Column | None} per value column ])``.
    Groups in ascending key order, the null key last; see include/b200df.h bdf_group_aggregate_dev."""
    ctx = key.ctx
    k = len(values)
    cols = (C.c_void_p * max(k, 1))(*[v.handle for v in values])
    outs = (N.GroupOut * max(k, 1))()
    hk, ng = C.c_void_p(), C.c_int64(0)
    N.raise_for_status(N.lib().bdf_group_aggregate_dev(ctx.handle, key.handle, k, cols, C.byref(hk), outs, C.byref(ng)))
    wrap = lambda p: Column(ctx, C.c_void_p(p)) if p else None
    return Column(ctx, hk), [{"sum": wrap(outs[j].sum), "count": wrap(outs[j].count), "min": wrap(outs[j].min), "max": wrap(outs[j].max)} for j in range(k)]


def _expr_nodes(nodes: Sequence[tuple]):
    arr = (N.ExprNode * len(nodes))()
    for k, nd in enumerate(nodes):
        if len(nd) == 2:
            op = nd[0]
            if isinstance(op, str):
                op = N.EXPR_UNARY + getattr(N, op.upper())
            arr[k].op, arr[k].a, arr[k].b = op, nd[1], nd[1]
        else:
            arr[k].op, arr[k].a, arr[k].b = nd
    return arr


def eval_expr(inputs: Sequence["Column"], nodes: Sequence[tuple]) -> "Column":
    """Evaluate a chain of Calculations in ONE pass over Float64 columns.

    ``nodes`` is a straight-line program: slots ``0..len(inputs)-1`` are the input columns, node ``k`` writes slot
    ``len(inputs)+k``.  A node is ``(binop, a, b)`` with ``binop`` in native.ADD..LOG, or ``("sin", a)`` / ``(native.EXPR_UNARY + unop, a)``
    for a unary function.  The last node is the result column; intermediates never touch HBM."""
    ctx = inputs[0].ctx
    cols = (C.c_void_p * len(inputs))(*[c.handle for c in inputs])
    h = C.c_void_p()
    N.raise_for_status(N.lib().bdf_eval_expr_dev(ctx.handle, len(inputs), cols, len(nodes), _expr_nodes(nodes), C.byref(h)))
    return Column(ctx, h)


def eval_expr_agg(inputs: Sequence["Column"], nodes: Sequence[tuple], materialise: bool = True, asynchronous: bool = False):
    """eval_expr with sum/count of the result folded into the same pass.  Returns ``(Column or None, aggregate)`` where the
    aggregate is a dict (or an AggFuture when ``asynchronous``); ``materialise=False`` never writes the result column."""
    ctx = inputs[0].ctx
    cols = (C.c_void_p * len(inputs))(*[c.handle for c in inputs])
    h = C.c_void_p()
    hp = C.byref(h) if materialise else None
    if asynchronous:
        f = C.c_void_p()
        N.raise_for_status(N.lib().bdf_eval_expr_agg_dev_async(ctx.handle, len(inputs), cols, len(nodes), _expr_nodes(nodes), hp, C.byref(f)))
        agg = AggFuture(ctx, f, F64)
    else:
        a = N.Agg4()
        N.raise_for_status(N.lib().bdf_eval_expr_agg_dev(ctx.handle, len(inputs), cols, len(nodes), _expr_nodes(nodes), hp, C.byref(a)))
        agg = _agg4_to_dict(F64, a)
    return (Column(ctx, h) if materialise else None), agg


class AggFuture:
    """An aggregate whose kernels are enqueued; result() blocks until the value is on the host."""

    def __init__(self, ctx: N.Context, handle: C.c_void_p, dtype):
        """dtype: an int (one aggregate -> result() is a dict) or a list (multi-column call -> a list of dicts)."""
        self.ctx, self.handle, self.dtype = ctx, handle, dtype
        self._value = None

    def result(self):
        if self.handle is not None:
            many = isinstance(self.dtype, (list, tuple))
            k = len(self.dtype) if many else 1
            a = (N.Agg4 * k)()
            N.raise_for_status(N.lib().bdf_future_wait(self.ctx.handle, self.handle, a))
            self.handle = None
            self._value = [_agg4_to_dict(self.dtype[i], a[i]) for i in range(k)] if many else _agg4_to_dict(self.dtype, a[0])
        return self._value

    def __del__(self):
        try:
            if self.handle is not None and self.ctx.handle:
                N.lib().bdf_future_wait(self.ctx.handle, self.handle, None)
                self.handle = None
        except Exception:
            pass
