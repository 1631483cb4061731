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
/// bdf_group_out: per value column; `min` / `max` are NULL for Float columns (T::Native: Ord).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct BdfGroupOut {
    pub sum: *mut BdfCol,
    pub count: *mut BdfCol,
    pub min: *mut BdfCol,
    pub max: *mut BdfCol,
}

#[repr(C)]
pub struct BdfSortKey { pub column: *const BdfCol, pub descending: i32 }
pub const BDF_ASYNC: c_int = 1;

/// All four aggregates of one pass (include/b200df.h `bdf_agg4`, ABI version 2: `n_chunks` added).
#[repr(C)]
pub struct BdfAgg4 { pub sum: u64, pub min: u64, pub max: u64, pub count: i64, pub rows: i64, pub any_valid: i32, pub would_panic: i32, pub n_chunks: i64 }
pub enum BdfFuture {}

pub const BDF_OK: c_int = 0;
pub const BDF_LENGTH_MISMATCH: c_int = 1;
pub const BDF_DIVIDE_BY_ZERO: c_int = 2;
pub const BDF_WOULD_PANIC: c_int = 7;

extern "C" {
    pub fn bdf_init(device: c_int, out: *mut *mut BdfCtx) -> c_int;
    /// ONE context over `n_gpus` GPUs of the box (0 = all visible): the library shards every call by row range, each GPU
    /// moves its pieces over its own PCIe link, aggregates are combined over NVLink (grouped ncclAllReduce / peer-memory
    /// mailboxes).  What a single Rust process binds; every entry below accepts it.
    pub fn bdf_init_multi(n_gpus: c_int, devices: *const c_int, out: *mut *mut BdfCtx) -> c_int;
    pub fn bdf_fleet_size(ctx: *mut BdfCtx) -> c_int;
    /// Process per GPU (MPI-style launchers): rank 0 makes the id, ships the 128 bytes, every rank attaches its one-GPU context.
    pub fn bdf_comm_unique_id(id: *mut u8) -> c_int;
    pub fn bdf_comm_attach(ctx: *mut BdfCtx, id: *const u8, rank: c_int, world: c_int) -> c_int;
    pub fn bdf_aggregate_all(ctx: *mut BdfCtx, dtype: c_int, n: i64, input: *const BdfView, out: *mut BdfAgg4) -> c_int;
    pub fn bdf_aggregate_all_many_dev(ctx: *mut BdfCtx, n_cols: i32, cols: *const *const BdfCol, out: *mut BdfAgg4) -> c_int;
    pub fn bdf_binary_agg_dev_async(ctx: *mut BdfCtx, op: c_int, left: *const BdfCol, right: *const BdfCol, out: *mut *mut BdfCol,
                                    fut: *mut *mut BdfFuture) -> c_int;
    pub fn bdf_future_wait(ctx: *mut BdfCtx, fut: *mut BdfFuture, out: *mut BdfAgg4) -> c_int;
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
    // group-by aggregate (Transformation::GroupAggregate, src/evaluation.rs:73; planned shape src/expression.rs:114-221)
    pub fn bdf_group_aggregate_dev(ctx: *mut BdfCtx, key: *const BdfCol, n_values: i32, values: *const *const BdfCol, out_keys: *mut *mut BdfCol,
                                   out: *mut BdfGroupOut, n_groups: *mut i64) -> c_int;
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

/// Process-wide context.  A plain process gets ONE context over every GPU of the box (`bdf_init_multi(0, ..)`: the library
/// shards each call over the GPUs -- the role rayon's par_iter plays in the reference, src/functions/scalar.rs:28-31).  Under a
/// process-per-GPU launcher (RANK / WORLD_SIZE / LOCAL_RANK set) each process takes its LOCAL_RANK's GPU and the launcher's
/// side channel carries the communicator id: see `attach` below.  Set BDF_ONE_GPU=1 to pin a plain process to GPU 0.
pub fn ctx() -> *mut BdfCtx {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut CTX: *mut BdfCtx = std::ptr::null_mut();
    unsafe {
        INIT.call_once(|| {
            let mut c = std::ptr::null_mut();
            let launcher = std::env::var("WORLD_SIZE").ok().and_then(|s| s.parse::<i32>().ok()).map_or(false, |w| w > 1);
            let st = if launcher || std::env::var("BDF_ONE_GPU").is_ok() {
                let dev = std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
                bdf_init(dev, &mut c)
            } else {
                bdf_init_multi(0, std::ptr::null(), &mut c)
            };
            assert_eq!(st, BDF_OK, "libb200df initialisation failed: {}", last_error());
            CTX = c;
        });
        CTX
    }
}

/// A one-GPU context for the entries a multi-GPU context does not offer (sort / take / filter move rows between chunks; the IPC
/// readers map one file): GPU LOCAL_RANK (0 in a plain process).  The same object as `ctx()` under a process-per-GPU launcher.
pub fn ctx_one_gpu() -> *mut BdfCtx {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut CTX: *mut BdfCtx = std::ptr::null_mut();
    unsafe {
        INIT.call_once(|| {
            let all = ctx();
            if bdf_fleet_size(all) <= 1 { CTX = all; return; }
            let mut c = std::ptr::null_mut();
            let st = bdf_init(0, &mut c);
            assert_eq!(st, BDF_OK, "bdf_init failed: {}", last_error());
            CTX = c;
        });
        CTX
    }
}

/// Process per GPU: make this process's context a rank of the job's communicator.  `id` = the 128 bytes rank 0 obtained from
/// `bdf_comm_unique_id` and the launcher distributed (an MPI broadcast, a file, a TCP store ...).  Afterwards every aggregate
/// entry returns the aggregate of the whole sharded column on every rank.
pub fn attach(id: &[u8; 128], rank: i32, world: i32) -> Result<(), ArrowError> {
    let st = unsafe { bdf_comm_attach(ctx(), id.as_ptr(), rank, world) };
    if st == BDF_OK { Ok(()) } else { Err(to_arrow_error(st)) }
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
