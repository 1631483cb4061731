"""ctypes front-end for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package.  Chunks are duck-typed: any object with
``dtype, values (np.ndarray), validity (np.ndarray[uint8] | None), offset, length, null_count``.
Results come back as plain ``OracleArray`` tuples of numpy buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

I8, I16, I32, I64, U8, U16, U32, U64, F32, F64 = range(10)
NP_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]
ADD, SUB, MUL, DIV, ATAN2, HYPOT, LOG = range(7)
(ABS, SIN, COS, TAN, ACOS, ASIN, ATAN, CBRT, CEIL, COSH, DEGREES, EXP, EXPM1, FLOOR, LOG10, LOG2, RADIANS, ROUND, SINH,
 SQRT, TANH) = range(21)
SUM, MIN, MAX, COUNT, MIN_AS_WRITTEN = range(5)
OK, LENGTH_MISMATCH, DIVIDE_BY_ZERO, UNSUPPORTED, PANIC = 0, 1, 2, 3, 7


class View(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("len", C.c_int64), ("offset", C.c_int64),
                ("null_count", C.c_int64)]


class Out(C.Structure):
    _fields_ = [("values", C.c_void_p), ("validity", C.c_void_p), ("len", C.c_int64), ("null_count", C.c_int64),
                ("has_validity", C.c_int32)]


@dataclass
class OracleArray:
    dtype: int
    values: np.ndarray            # length == len, offset 0
    validity: Optional[np.ndarray]  # uint8 bitmap (offset 0) or None when arrow-rs attaches none
    null_count: int

    @property
    def length(self) -> int:
        return int(self.values.shape[0])

    offset = 0

    def valid_mask(self) -> np.ndarray:
        if self.validity is None:
            return np.ones(self.length, dtype=bool)
        return np.unpackbits(self.validity, bitorder="little")[: self.length].astype(bool)


def build(force: bool = False) -> str:
    """Compile liboracle.so from oracle.c (gcc; see Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        PV, PO = C.POINTER(View), C.POINTER(Out)
        L.orc_binary.argtypes = [C.c_int, C.c_int, PV, PV, PO]
        L.orc_unary.argtypes = [C.c_int, C.c_int, PV, PO]
        L.orc_cast.argtypes = [C.c_int, C.c_int, PV, PO]
        L.orc_col_binary.argtypes = [C.c_int, C.c_int, C.c_int64, PV, C.c_int64, PV, PO, C.c_int]
        L.orc_col_unary.argtypes = [C.c_int, C.c_int, C.c_int64, PV, PO, C.c_int]
        L.orc_col_cast.argtypes = [C.c_int, C.c_int, C.c_int64, PV, PO, C.c_int]
        L.orc_aggregate.argtypes = [C.c_int, C.c_int, C.c_int64, PV, C.c_void_p, C.POINTER(C.c_int32)]
        L.orc_avg.argtypes = [C.c_int, C.c_int64, PV, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.orc_sum_exact.argtypes = [C.c_int, C.c_int64, PV, C.POINTER(C.c_longdouble), C.POINTER(C.c_longdouble)]
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_generate.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_int64,
                                   C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_generate.restype = None
        L.orc_null_count.argtypes = [PV]
        L.orc_null_count.restype = C.c_int64
        _lib = L
    return _lib


def _views(chunks: Sequence) -> "C.Array[View]":
    arr = (View * max(len(chunks), 1))()
    for i, c in enumerate(chunks):
        arr[i].values = c.values.ctypes.data if c.values.size else 0
        arr[i].validity = c.validity.ctypes.data if c.validity is not None else None
        arr[i].len = c.length
        arr[i].offset = c.offset
        arr[i].null_count = getattr(c, "null_count", -1)
    return arr


def _alloc_outs(dtype: int, lens: Sequence[int]):
    outs = (Out * max(len(lens), 1))()
    bufs = []
    for i, n in enumerate(lens):
        v = np.zeros(n, dtype=NP_DTYPES[dtype])
        b = np.zeros((n + 7) // 8, dtype=np.uint8)
        outs[i].values = v.ctypes.data if n else 0
        outs[i].validity = b.ctypes.data if b.size else 0
        bufs.append((v, b))
    return outs, bufs


def _collect(dtype, outs, bufs):
    res = []
    for i, (v, b) in enumerate(bufs):
        res.append(OracleArray(dtype, v, b if outs[i].has_validity else None, int(outs[i].null_count)))
    return res


def col_binary(op: int, dtype: int, left: Sequence, right: Sequence, threads: int = 1):
    """ScalarFunctions::{add,subtract,multiply,divide}.  Returns (status, [OracleArray])."""
    n = min(len(left), len(right))
    outs, bufs = _alloc_outs(dtype, [left[i].length for i in range(n)])
    st = lib().orc_col_binary(op, dtype, len(left), _views(left), len(right), _views(right), outs, threads)
    return st, (_collect(dtype, outs, bufs) if st == OK else None)


def col_unary(op: int, dtype: int, chunks: Sequence, threads: int = 1):
    outs, bufs = _alloc_outs(dtype, [c.length for c in chunks])
    st = lib().orc_col_unary(op, dtype, len(chunks), _views(chunks), outs, threads)
    return st, (_collect(dtype, outs, bufs) if st == OK else None)


def col_cast(from_t: int, to_t: int, chunks: Sequence, threads: int = 1):
    outs, bufs = _alloc_outs(to_t, [c.length for c in chunks])
    st = lib().orc_col_cast(from_t, to_t, len(chunks), _views(chunks), outs, threads)
    return st, (_collect(to_t, outs, bufs) if st == OK else None)


def aggregate(op: int, dtype: int, chunks: Sequence):
    """AggregateFunctions::{sum,min,max,count}.  Returns (status, value | None)."""
    out_dtype = np.int64 if op == COUNT else NP_DTYPES[dtype]
    out = np.zeros(1, dtype=out_dtype)
    some = C.c_int32(0)
    st = lib().orc_aggregate(op, dtype, len(chunks), _views(chunks), out.ctypes.data, C.byref(some))
    if st != OK:
        return st, None
    return st, (out[0] if some.value else None)


def avg(dtype: int, chunks: Sequence):
    out = C.c_double(0)
    some = C.c_int32(0)
    st = lib().orc_avg(dtype, len(chunks), _views(chunks), C.byref(out), C.byref(some))
    return st, (out.value if (st == OK and some.value) else None)


def sum_exact(dtype: int, chunks: Sequence):
    """(compensated long-double sum, sum of |x|) over valid slots, as Python floats (np.longdouble)."""
    s, sa = C.c_longdouble(0), C.c_longdouble(0)
    st = lib().orc_sum_exact(dtype, len(chunks), _views(chunks), C.byref(s), C.byref(sa))
    assert st == OK
    return np.longdouble(s.value), np.longdouble(sa.value)


GT, GE, EQ, NE, LT, LE = range(6)
AND, OR, NOT = range(3)
BOOL = 10


def _bool_out(n: int):
    out = (Out * 1)()
    v = np.zeros((n + 7) // 8, dtype=np.uint8)
    b = np.zeros((n + 7) // 8, dtype=np.uint8)
    out[0].values = v.ctypes.data if v.size else 0
    out[0].validity = b.ctypes.data if b.size else 0
    return out, v, b


@dataclass
class OracleBoolArray:
    """Arrow BooleanArray: bit-packed values + optional validity."""
    values: np.ndarray
    validity: Optional[np.ndarray]
    length: int
    null_count: int
    dtype = BOOL
    offset = 0

    def value_bits(self) -> np.ndarray:
        return np.unpackbits(self.values, bitorder="little")[: self.length].astype(bool)

    def valid_mask(self) -> np.ndarray:
        if self.validity is None:
            return np.ones(self.length, dtype=bool)
        return np.unpackbits(self.validity, bitorder="little")[: self.length].astype(bool)


def compare(op: int, left, right=None, scalar: Optional[float] = None):
    """BooleanFilter::{Gt,..}: one chunk.  right=None + scalar broadcasts a BooleanInput::Scalar."""
    L = lib()
    L.orc_compare.argtypes = [C.c_int, C.c_int, C.POINTER(View), C.c_int, C.POINTER(View), C.c_int, C.c_double, C.POINTER(Out)]
    out, v, b = _bool_out(left.length)
    use_scalar = right is None
    st = L.orc_compare(op, left.dtype, _views([left]), 0 if use_scalar else right.dtype, None if use_scalar else _views([right]),
                       1 if use_scalar else 0, float(scalar or 0.0), out)
    if st != OK:
        return st, None
    return st, OracleBoolArray(v, b if out[0].has_validity else None, int(out[0].len), int(out[0].null_count))


def boolean(op: int, a, b=None):
    L = lib()
    L.orc_bool.argtypes = [C.c_int, C.POINTER(View), C.POINTER(View), C.POINTER(Out)]
    out, v, bb = _bool_out(a.length)
    st = L.orc_bool(op, _views([a]), None if b is None else _views([b]), out)
    if st != OK:
        return st, None
    return st, OracleBoolArray(v, bb if out[0].has_validity else None, int(out[0].len), int(out[0].null_count))


def filter_chunk(values, mask):
    """arrow::compute::filter for one chunk; values may be a primitive chunk or an OracleBoolArray."""
    L = lib()
    L.orc_filter.argtypes = [C.c_int, C.POINTER(View), C.POINTER(View), C.POINTER(Out)]
    n = values.length
    out = (Out * 1)()
    is_bool = values.dtype == BOOL
    v = np.zeros((n + 7) // 8, dtype=np.uint8) if is_bool else np.zeros(n, dtype=NP_DTYPES[values.dtype])
    b = np.zeros((n + 7) // 8, dtype=np.uint8)
    out[0].values = v.ctypes.data if v.size else 0
    out[0].validity = b.ctypes.data if b.size else 0
    st = L.orc_filter(values.dtype, _views([values]), _views([mask]), out)
    if st != OK:
        return st, None
    k = int(out[0].len)
    if is_bool:
        return st, OracleBoolArray(v[: (k + 7) // 8], b[: (k + 7) // 8] if out[0].has_validity else None, k, int(out[0].null_count))
    return st, OracleArray(values.dtype, v[:k], b[: (k + 7) // 8] if out[0].has_validity else None, int(out[0].null_count))


class SortKey(C.Structure):
    _fields_ = [("dtype", C.c_int), ("n_chunks", C.c_int64), ("chunks", C.POINTER(View)), ("descending", C.c_int)]


def lexsort_indices(keys: Sequence[tuple]):
    """DataFrame::sort's index step.  keys: [(chunks, descending)], chunks = host arrays of one numeric column.
    Returns (status, np.uint32 row numbers)."""
    L = lib()
    L.orc_lexsort_indices.argtypes = [C.c_int, C.POINTER(SortKey), C.POINTER(C.c_uint32)]
    arr = (SortKey * max(len(keys), 1))()
    keep = []
    n = 0
    for i, (chunks, desc) in enumerate(keys):
        v = _views(chunks)
        keep.append(v)
        arr[i].dtype = chunks[0].dtype if chunks else F64
        arr[i].n_chunks = len(chunks)
        arr[i].chunks = C.cast(v, C.POINTER(View))
        arr[i].descending = 1 if desc else 0
        n = sum(c.length for c in chunks)
    out = np.zeros(max(n, 1), dtype=np.uint32)
    st = L.orc_lexsort_indices(len(keys), arr, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return st, (out[:n] if st == OK else None)


def take(chunks: Sequence, indices: np.ndarray, index_valid: Optional[np.ndarray] = None):
    """Column::take: one output chunk.  indices: uint32 row numbers over the concatenated chunks; index_valid: bool mask."""
    L = lib()
    L.orc_take.argtypes = [C.c_int, C.c_int64, C.POINTER(View), C.c_int64, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(Out)]
    dtype = chunks[0].dtype
    idx = np.ascontiguousarray(indices, dtype=np.uint32)
    n = idx.shape[0]
    iv = None if index_valid is None else np.packbits(np.asarray(index_valid, bool), bitorder="little")
    is_bool = dtype == BOOL
    v = np.zeros((n + 7) // 8, dtype=np.uint8) if is_bool else np.zeros(n, dtype=NP_DTYPES[dtype])
    b = np.zeros((n + 7) // 8, dtype=np.uint8)
    out = (Out * 1)()
    out[0].values = v.ctypes.data if v.size else 0
    out[0].validity = b.ctypes.data if b.size else 0
    st = L.orc_take(dtype, len(chunks), _views(chunks), n, idx.ctypes.data_as(C.POINTER(C.c_uint32)) if n else None,
                    iv.ctypes.data if iv is not None and iv.size else None, out)
    if st != OK:
        return st, None
    if is_bool:
        return st, OracleBoolArray(v, b if out[0].has_validity else None, n, int(out[0].null_count))
    return st, OracleArray(dtype, v, b if out[0].has_validity else None, int(out[0].null_count))


def generate(dtype: int, kind: int, lo: float, hi: float, seed: int, col: int, row0: int, length: int,
             null_mod: int = 0) -> OracleArray:
    v = np.zeros(length, dtype=NP_DTYPES[dtype])
    b = np.zeros((length + 7) // 8, dtype=np.uint8) if null_mod else None
    nc = C.c_int64(0)
    lib().orc_generate(dtype, kind, lo, hi, seed, col, row0, length, null_mod, v.ctypes.data if length else 0,
                       b.ctypes.data if (b is not None and b.size) else None, C.byref(nc))
    return OracleArray(dtype, v, b, int(nc.value))


# ---------------------------------------------------------------------------------------------------------
# group-by aggregate (numpy restatement; the reference's GroupAggregate is a panic!, src/evaluation.rs:73 -- the operator is
# DEFINED by include/b200df.h bdf_group_aggregate_dev: one numeric key, groups in ascending key order with the null key last
# (DataFrame::sort's order), NaN keys one group after every number, -0.0 == 0.0; per group the aggregates of
# AggregateFunctions (aggregate.rs:12-93): wrapping integer sum, valid count, min / max over valid slots (None if there is none)).
# Parity unpinned: no reference test exists for this operator; cross-checked against pyarrow's group_by in tests/test_group_gpu.py.

def group_aggregate(key_chunks, value_chunks):
    """key_chunks / value_chunks: lists of chunks (same lengths).  Returns
    (keys ndarray, key_valid bool ndarray, {"sum": ndarray, "count": int64 ndarray, "min"/"max": (ndarray, valid) or None,
    "exact": longdouble sums and sums of magnitudes for float columns})."""
    kv = np.concatenate([np.asarray(c.values)[c.offset:c.offset + c.length] for c in key_chunks]) if key_chunks else np.zeros(0)
    km = np.concatenate([c.valid_mask() for c in key_chunks]) if key_chunks else np.zeros(0, bool)
    vv = np.concatenate([np.asarray(c.values)[c.offset:c.offset + c.length] for c in value_chunks])
    vm = np.concatenate([c.valid_mask() for c in value_chunks])
    n = kv.shape[0]
    if kv.dtype.kind == "f":
        isnan = np.isnan(kv)
        ck = np.where(isnan, 0.0, kv) + 0.0          # -0.0 + 0.0 = +0.0: one group for both zeros
    else:
        isnan = np.zeros(n, bool)
        ck = kv
    ck = np.where(km, ck, 0)                         # payload under a null key does not matter
    isnan = isnan & km
    order = np.lexsort((ck, isnan, ~km))             # stable: nulls last, NaN after every number, then the value
    sk, sn, sm = ck[order], isnan[order], km[order]
    head = np.ones(n, bool)
    if n > 1:
        same = (sm[1:] == sm[:-1]) & (~sm[1:] | ((sn[1:] == sn[:-1]) & (sn[1:] | (sk[1:] == sk[:-1]))))
        head[1:] = ~same
    starts = np.nonzero(head)[0]
    keys = kv[order][starts] if n else kv[:0]
    key_valid = km[order][starts] if n else km[:0]
    sv, svm = vv[order], vm[order]
    counts = np.add.reduceat(svm.astype(np.int64), starts) if n else np.zeros(0, np.int64)
    out = {"count": counts, "min": None, "max": None, "exact": None}
    if vv.dtype.kind == "f":
        ld = np.where(svm, sv, 0).astype(np.longdouble)
        out["sum"] = np.add.reduceat(np.where(svm, sv, 0).astype(np.float64), starts).astype(vv.dtype) if n else vv[:0]
        out["exact"] = (np.add.reduceat(ld, starts), np.add.reduceat(np.abs(ld), starts)) if n else (ld[:0], ld[:0])
    else:
        u = np.dtype(f"u{vv.dtype.itemsize}")
        # dtype=u: numpy would otherwise accumulate small integers in the platform word; the reference wraps in T's width
        out["sum"] = np.add.reduceat(np.where(svm, sv, 0).astype(vv.dtype).view(u), starts, dtype=u).view(vv.dtype) if n else vv[:0]
        info = np.iinfo(vv.dtype)
        out["min"] = (np.minimum.reduceat(np.where(svm, sv, info.max), starts), counts > 0) if n else (vv[:0], counts > 0)
        out["max"] = (np.maximum.reduceat(np.where(svm, sv, info.min), starts), counts > 0) if n else (vv[:0], counts > 0)
    return keys, key_valid, out
