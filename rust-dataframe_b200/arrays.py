"""Host-side Arrow ``PrimitiveArray<T>`` stand-in used by the Python mirror of the reference API.

The reference passes ``Vec<&PrimitiveArray<T>>`` (one array per RecordBatch chunk, src/table.rs:114-123)
into ``ScalarFunctions`` / ``AggregateFunctions``.  This class keeps the same memory contract the C ABI
consumes (include/b200df.h ``bdf_view``): a values buffer, an optional LSB-first validity bitmap
(1 = valid), a logical ``offset`` that applies to both, ``len`` and a cached ``null_count``.
pyarrow arrays convert in both directions without copying the buffers.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np

I8, I16, I32, I64, U8, U16, U32, U64, F32, F64 = range(10)
NP_DTYPES = [np.dtype(t) for t in (np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64,
                                   np.float32, np.float64)]
DTYPE_NAMES = ["Int8", "Int16", "Int32", "Int64", "UInt8", "UInt16", "UInt32", "UInt64", "Float32", "Float64"]
_NP_TO_DTYPE = {dt: i for i, dt in enumerate(NP_DTYPES)}


def dtype_of(np_dtype) -> int:
    return _NP_TO_DTYPE[np.dtype(np_dtype)]


def width_of(dtype: int) -> int:
    return NP_DTYPES[dtype].itemsize


def is_float(dtype: int) -> bool:
    return dtype in (F32, F64)


def pack_validity(mask: np.ndarray) -> np.ndarray:
    """bool mask (True = valid) -> Arrow bitmap bytes (LSB first), padded to whole bytes."""
    return np.packbits(np.asarray(mask, dtype=bool), bitorder="little")


class PrimitiveArray:
    """One chunk of a column: Arrow layout over numpy buffers."""

    __slots__ = ("dtype", "values", "validity", "offset", "length", "null_count", "_keepalive")

    def __init__(self, dtype: int, values: np.ndarray, validity: Optional[np.ndarray] = None, offset: int = 0,
                 length: Optional[int] = None, null_count: int = -1, keepalive=None):
        values = np.asarray(values)
        if values.dtype != NP_DTYPES[dtype]:
            raise TypeError(f"values dtype {values.dtype} does not match {DTYPE_NAMES[dtype]}")
        if not values.flags.c_contiguous:
            values = np.ascontiguousarray(values)
        self.dtype = dtype
        self.values = values
        self.validity = None if validity is None else np.ascontiguousarray(validity, dtype=np.uint8)
        self.offset = int(offset)
        self.length = int(values.shape[0] - offset if length is None else length)
        if self.offset < 0 or self.length < 0 or self.offset + self.length > values.shape[0]:
            raise ValueError("offset/length outside the values buffer")
        if self.validity is not None and self.validity.shape[0] * 8 < self.offset + self.length:
            raise ValueError("validity bitmap too short")
        self.null_count = 0 if self.validity is None else int(null_count)
        self._keepalive = keepalive

    # ---- constructors ------------------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, values: np.ndarray, mask: Optional[np.ndarray] = None) -> "PrimitiveArray":
        """mask: True = valid.  No bitmap is attached when mask is None (like a Vec<T> -> array)."""
        values = np.ascontiguousarray(values)
        dtype = dtype_of(values.dtype)
        if mask is None:
            return cls(dtype, values)
        mask = np.asarray(mask, dtype=bool)
        return cls(dtype, values, pack_validity(mask), 0, values.shape[0], int((~mask).sum()))

    @classmethod
    def from_pylist(cls, dtype: int, items: Iterable) -> "PrimitiveArray":
        """None entries become nulls (payload 0), like Arrow's From<Vec<Option<T>>>."""
        items = list(items)
        mask = np.array([x is not None for x in items], dtype=bool)
        vals = np.array([0 if x is None else x for x in items], dtype=NP_DTYPES[dtype])
        if mask.all():
            return cls(dtype, vals)
        return cls(dtype, vals, pack_validity(mask), 0, len(items), int((~mask).sum()))

    @classmethod
    def from_arrow(cls, arr) -> "PrimitiveArray":
        """Zero-copy view of a pyarrow primitive array (offset and validity preserved)."""
        import pyarrow as pa

        np_dt = np.dtype(arr.type.to_pandas_dtype())
        dtype = dtype_of(np_dt)
        vbuf, dbuf = arr.buffers()
        n_total = arr.offset + len(arr)
        values = (np.frombuffer(dbuf, dtype=np_dt, count=n_total) if dbuf is not None and n_total
                  else np.zeros(n_total, dtype=np_dt))
        validity = np.frombuffer(vbuf, dtype=np.uint8) if vbuf is not None else None
        return cls(dtype, values, validity, arr.offset, len(arr), arr.null_count if vbuf is not None else 0,
                   keepalive=arr)

    def to_arrow(self):
        import pyarrow as pa

        typ = pa.from_numpy_dtype(NP_DTYPES[self.dtype])
        vbuf = pa.py_buffer(self.validity) if self.validity is not None else None
        return pa.Array.from_buffers(typ, self.length, [vbuf, pa.py_buffer(self.values)],
                                     null_count=self.null_count if self.validity is not None else 0,
                                     offset=self.offset)

    # ---- accessors (names follow arrow-rs) -----------------------------------------------------------
    def __len__(self) -> int:
        return self.length

    def len(self) -> int:
        return self.length

    def valid_mask(self) -> np.ndarray:
        if self.validity is None:
            return np.ones(self.length, dtype=bool)
        bits = np.unpackbits(self.validity, bitorder="little")
        return bits[self.offset:self.offset + self.length].astype(bool)

    def is_valid(self, i: int) -> bool:
        if self.validity is None:
            return True
        j = self.offset + i
        return bool((self.validity[j >> 3] >> (j & 7)) & 1)

    def is_null(self, i: int) -> bool:
        return not self.is_valid(i)

    def value(self, i: int):
        return self.values[self.offset + i]

    def value_slice(self) -> np.ndarray:
        return self.values[self.offset:self.offset + self.length]

    def compute_null_count(self) -> int:
        if self.validity is None:
            return 0
        if self.null_count < 0:
            self.null_count = int(self.length - self.valid_mask().sum())
        return self.null_count

    def slice(self, offset: int, length: int) -> "PrimitiveArray":
        """Zero-copy slice (ArrayData::slice): shares buffers, moves the offset."""
        if offset < 0 or length < 0 or offset + length > self.length:
            raise ValueError("slice out of bounds")
        return PrimitiveArray(self.dtype, self.values, self.validity, self.offset + offset, length, -1,
                              keepalive=self._keepalive)

    def to_pylist(self) -> List:
        m = self.valid_mask()
        v = self.value_slice()
        return [v[i].item() if m[i] else None for i in range(self.length)]

    def __repr__(self) -> str:
        return f"PrimitiveArray<{DTYPE_NAMES[self.dtype]}>(len={self.length}, offset={self.offset}, nulls={self.null_count})"


class BooleanArray:
    """Arrow BooleanArray chunk: bit-packed values (LSB first) + optional validity; `offset` counts bits."""

    dtype = 10  # BDF_BOOL
    __slots__ = ("values", "validity", "offset", "length", "null_count", "_keepalive")

    def __init__(self, values: np.ndarray, validity: Optional[np.ndarray] = None, offset: int = 0, length: int = 0,
                 null_count: int = -1, keepalive=None):
        self._keepalive = keepalive    # e.g. the mapped IPC file the buffers point into
        self.values = np.ascontiguousarray(values, dtype=np.uint8)
        self.validity = None if validity is None else np.ascontiguousarray(validity, dtype=np.uint8)
        self.offset, self.length = int(offset), int(length)
        self.null_count = 0 if self.validity is None else int(null_count)

    @classmethod
    def from_numpy(cls, bits: np.ndarray, mask: Optional[np.ndarray] = None) -> "BooleanArray":
        bits = np.asarray(bits, dtype=bool)
        validity = None if mask is None else pack_validity(mask)
        return cls(pack_validity(bits), validity, 0, bits.shape[0], 0 if mask is None else int((~np.asarray(mask, bool)).sum()))

    def slice(self, offset: int, length: int) -> "BooleanArray":
        return BooleanArray(self.values, self.validity, self.offset + offset, length, -1, keepalive=self._keepalive)

    def value_bits(self) -> np.ndarray:
        return np.unpackbits(self.values, bitorder="little")[self.offset:self.offset + self.length].astype(bool)

    def valid_mask(self) -> np.ndarray:
        if self.validity is None:
            return np.ones(self.length, dtype=bool)
        return np.unpackbits(self.validity, bitorder="little")[self.offset:self.offset + self.length].astype(bool)

    def to_pylist(self) -> List:
        v, m = self.value_bits(), self.valid_mask()
        return [bool(v[i]) if m[i] else None for i in range(self.length)]

    def __len__(self) -> int:
        return self.length


def chunk_lengths(chunks: Sequence[PrimitiveArray]) -> List[int]:
    return [c.length for c in chunks]
