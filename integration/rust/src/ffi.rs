//! FFI surface of libb200df.so (mirrors include/b200df.h).  UNCOMPILED in this repository's image -- see ../README.md.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

use arrow::array::{Array, PrimitiveArray};
use arrow::datatypes::{ArrowPrimitiveType, DataType};
use arrow::error::ArrowError;

#[repr(C)]
pub struct BdfView {
    pub values: *const c_void,   // base of the values buffer, NOT offset-adjusted
    pub validity: *const u8,     // null when the array has no null buffer
    pub len: i64,
    pub offset: i64,             // elements; applies to values and validity bits
    pub null_count: i64,         // -1 = unknown
}

#[repr(C)]
pub struct BdfOut {
    pub values: *mut c_void,
    pub validity: *mut u8,
    pub len: i64,                // in: capacity (= result length); out: length
    pub null_count: i64,
    pub has_validity: i32,
}

pub enum BdfCtx {}
pub enum BdfCol {}   // a column resident in HBM (one chunk per RecordBatch)
pub enum BdfIpc {}   // a mapped Arrow IPC file

#[repr(C)]
pub struct BdfExprNode { pub op: i32, pub a: i32, pub b: i32 }   // op: bdf_binop, or BDF_EXPR_UNARY + bdf_unop
pub const BDF_EXPR_UNARY: i32 = 100;
#[repr(C)]
pub struct BdfSortKey { pub column: *const BdfCol, pub descending: i32 }
pub const BDF_ASYNC: c_int = 1;

pub const BDF_OK: c_int = 0;
pub const BDF_LENGTH_MISMATCH: c_int = 1;
pub const BDF_DIVIDE_BY_ZERO: c_int = 2;
pub const BDF_WOULD_PANIC: c_int = 7;

extern "C" {
    pub fn bdf_init(device: c_int, out: *mut *mut BdfCtx) -> c_int;
    pub fn bdf_last_error() -> *const c_char;
    pub fn bdf_host_register(ctx: *mut BdfCtx, p: *mut c_void, bytes: usize) -> c_int;
    pub fn bdf_host_unregister(ctx: *mut BdfCtx, p: *mut c_void) -> c_int;
    pub fn bdf_binary(ctx: *mut BdfCtx, op: c_int, dtype: c_int, n_left: i64, left: *const BdfView, n_right: i64,
                      right: *const BdfView, out: *mut BdfOut) -> c_int;
    pub fn bdf_unary(ctx: *mut BdfCtx, op: c_int, dtype: c_int, n: i64, input: *const BdfView, out: *mut BdfOut) -> c_int;
    pub fn bdf_cast(ctx: *mut BdfCtx, from: c_int, to: c_int, n: i64, input: *const BdfView, out: *mut BdfOut) -> c_int;
    pub fn bdf_aggregate(ctx: *mut BdfCtx, op: c_int, dtype: c_int, n: i64, input: *const BdfView, out_scalar: *mut c_void,
                         is_some: *mut i32) -> c_int;
    pub fn bdf_avg(ctx: *mut BdfCtx, dtype: c_int, n: i64, input: *const BdfView, out: *mut f64, is_some: *mut i32) -> c_int;
    // device-resident columns (the lazy evaluator keeps a frame's numeric columns in HBM between Calculations)
    pub fn bdf_upload_many(ctx: *mut BdfCtx, n_cols: i64, dtypes: *const i32, n_chunks: *const i64, input: *const *const BdfView, flags: c_int,
                           out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_binary_dev(ctx: *mut BdfCtx, op: c_int, left: *const BdfCol, right: *const BdfCol, out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_unary_dev(ctx: *mut BdfCtx, op: c_int, input: *const BdfCol, out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_cast_dev(ctx: *mut BdfCtx, to: c_int, input: *const BdfCol, out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_eval_expr_dev(ctx: *mut BdfCtx, n_inputs: i32, inputs: *const *const BdfCol, n_nodes: i32, nodes: *const BdfExprNode,
                             out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_download(ctx: *mut BdfCtx, col: *const BdfCol, out: *mut BdfOut) -> c_int;
    pub fn bdf_col_free(ctx: *mut BdfCtx, col: *mut BdfCol);
    // DataFrame::sort
    pub fn bdf_sort_indices_dev(ctx: *mut BdfCtx, n_keys: i32, keys: *const BdfSortKey, indices: *mut *mut BdfCol) -> c_int;
    pub fn bdf_take_dev(ctx: *mut BdfCtx, values: *const BdfCol, indices: *const BdfCol, out: *mut *mut BdfCol) -> c_int;
    // Arrow IPC files (DataFrame::from_arrow / to_arrow)
    pub fn bdf_ipc_open(path: *const c_char, out: *mut *mut BdfIpc) -> c_int;
    pub fn bdf_ipc_close(file: *mut BdfIpc);
    pub fn bdf_ipc_describe(file: *const BdfIpc, n_columns: *mut i32, n_batches: *mut i64, n_rows: *mut i64) -> c_int;
    pub fn bdf_ipc_column(file: *const BdfIpc, col: i32, name: *mut *const c_char, dtype: *mut i32, nullable: *mut i32) -> c_int;
    pub fn bdf_ipc_read(ctx: *mut BdfCtx, file: *const BdfIpc, n_cols: i32, cols: *const i32, flags: c_int, out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_ipc_read_batches(ctx: *mut BdfCtx, file: *const BdfIpc, n_cols: i32, cols: *const i32, n_batches: i64, batches: *const i64,
                                flags: c_int, out: *mut *mut BdfCol) -> c_int;
    pub fn bdf_ipc_write(ctx: *mut BdfCtx, path: *const c_char, n_cols: i32, names: *const *const c_char, cols: *const *const BdfCol) -> c_int;
}

/// Process-wide context: one GPU per process, LOCAL_RANK selects the device under a multi-process launcher.
pub fn ctx() -> *mut BdfCtx {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut CTX: *mut BdfCtx = std::ptr::null_mut();
    unsafe {
        INIT.call_once(|| {
            let dev = std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut c = std::ptr::null_mut();
            let st = bdf_init(dev, &mut c);
            assert_eq!(st, BDF_OK, "bdf_init failed: {}", last_error());
            CTX = c;
        });
        CTX
    }
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(bdf_last_error()) }.to_string_lossy().into_owned()
}

/// bdf_dtype: Arrow DataType order Int8..UInt64, Float32, Float64.
pub fn dtype_id(t: &DataType) -> c_int {
    match t {
        DataType::Int8 => 0, DataType::Int16 => 1, DataType::Int32 => 2, DataType::Int64 => 3,
        DataType::UInt8 => 4, DataType::UInt16 => 5, DataType::UInt32 => 6, DataType::UInt64 => 7,
        DataType::Float32 => 8, DataType::Float64 => 9,
        other => panic!("{:?} is not a primitive numeric type", other),
    }
}

pub fn view<T: ArrowPrimitiveType>(a: &PrimitiveArray<T>) -> BdfView {
    let d = a.data();
    BdfView {
        values: d.buffers()[0].raw_data() as *const c_void,
        validity: d.null_buffer().map_or(std::ptr::null(), |b| b.raw_data()),
        len: a.len() as i64,
        offset: a.offset() as i64,
        null_count: a.null_count() as i64,
    }
}

pub fn to_arrow_error(st: c_int) -> ArrowError {
    match st {
        BDF_LENGTH_MISMATCH => ArrowError::ComputeError("Cannot perform math operation on arrays of different length".to_string()),
        BDF_DIVIDE_BY_ZERO => ArrowError::DivideByZero,
        _ => ArrowError::ComputeError(last_error()),
    }
}
