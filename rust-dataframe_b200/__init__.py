"""rust-dataframe_b200: the Blackwell (sm_100a) execution path for rust-dataframe's per-RecordBatch compute
(elementwise add/sub/mul/div, trig, numeric cast, sum/min/max/count with null-bitmap propagation).

The product is ``libb200df.so`` -- hand-written CUDA kernels behind the C ABI in ``include/b200df.h``.
This package is the thin host-side mirror of the reference's operator interface over that ABI
(``ScalarFunctions``, ``AggregateFunctions``, ``cast``) plus the device-resident ``Column`` handle.
Import as ``rust_dataframe_b200`` (alias module at the repository root).
"""
from . import _native as native
from ._native import (ArrowError, ComputeError, Context, DivideByZero, ReferencePanic, UnsupportedType,
                      default_context)
from .arrays import (BooleanArray, DTYPE_NAMES, F32, F64, I8, I16, I32, I64, NP_DTYPES, U8, U16, U32, U64, PrimitiveArray, dtype_of,
                     width_of)
from .ipc import IpcFile, write_ipc, write_ipc_host
from . import frame
from .frame import DeviceFrame, plan_fusion
from .functions import AggFuture, AggregateFunctions, Column, ScalarFunctions, cast, eval_expr, eval_expr_agg, group_aggregate, sort_columns, sort_indices

__all__ = [
    "native", "ArrowError", "ComputeError", "Context", "DivideByZero", "ReferencePanic", "UnsupportedType",
    "default_context", "PrimitiveArray", "BooleanArray", "ScalarFunctions", "AggregateFunctions", "Column", "AggFuture", "cast", "eval_expr", "eval_expr_agg", "group_aggregate", "sort_indices", "sort_columns", "IpcFile", "write_ipc", "write_ipc_host", "frame", "DeviceFrame", "plan_fusion",
    "I8", "I16", "I32", "I64", "U8", "U16", "U32", "U64", "F32", "F64", "NP_DTYPES", "DTYPE_NAMES", "dtype_of", "width_of",
]
