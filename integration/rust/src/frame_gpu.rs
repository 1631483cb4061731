//! Device-resident frames: how `DataFrame::{from_arrow, to_arrow, sort}`, the group-by aggregate and a fused run of Calculations map onto
//! libb200df.so.  UNCOMPILED here (see ../README.md); the same flow is exercised from Python in
//! rust-dataframe_b200/frame.py (tests/test_frame.py, tests/test_ipc.py, tests/test_sort_gpu.py).
use std::ffi::{CStr, CString};
use std::os::raw::c_char;

use crate::ffi::*;

/// Columns of a frame that live in HBM: one `BdfCol` per numeric/boolean column, one chunk per RecordBatch.
pub struct DeviceColumns {
    pub names: Vec<String>,
    pub cols: Vec<*mut BdfCol>,
}

impl Drop for DeviceColumns {
    fn drop(&mut self) {
        for c in &self.cols {
            if !c.is_null() { unsafe { bdf_col_free(dev(), *c) }; }
        }
    }
}

/// The frames of this module live on ONE GPU: sort, take and the IPC readers are one-GPU entries (include/b200df.h).
fn dev() -> *mut BdfCtx { ctx_one_gpu() }

/// DataFrame::from_arrow (src/dataframe.rs:391-407) for the columns on the path: the file is mapped and its body buffers
/// go straight to the device.  Columns reported with dtype -1 (Utf8, List, ...) keep coming from arrow's FileReader.
pub fn from_arrow_device(path: &str) -> Result<DeviceColumns, String> {
    let cpath = CString::new(path).unwrap();
    let mut f: *mut BdfIpc = std::ptr::null_mut();
    check(unsafe { bdf_ipc_open(cpath.as_ptr(), &mut f) })?;
    // the file handle is closed on every path out of this function
    struct Close(*mut BdfIpc);
    impl Drop for Close { fn drop(&mut self) { unsafe { bdf_ipc_close(self.0) } } }
    let _close = Close(f);
    let (mut n_cols, mut n_batches, mut n_rows) = (0i32, 0i64, 0i64);
    check(unsafe { bdf_ipc_describe(f, &mut n_cols, &mut n_batches, &mut n_rows) })?;
    let mut wanted = vec![];
    let mut names = vec![];
    for c in 0..n_cols {
        let (mut name, mut dtype, mut nullable): (*const c_char, i32, i32) = (std::ptr::null(), -1, 0);
        check(unsafe { bdf_ipc_column(f, c, &mut name, &mut dtype, &mut nullable) })?;
        if dtype >= 0 {
            wanted.push(c);
            names.push(unsafe { CStr::from_ptr(name) }.to_string_lossy().into_owned());
        }
    }
    if wanted.is_empty() {
        return Ok(DeviceColumns { names, cols: vec![] });   // no numeric/boolean column: nothing for the device (bdf_ipc_read rejects n_cols == 0)
    }
    let mut cols = vec![std::ptr::null_mut(); wanted.len()];
    check(unsafe { bdf_ipc_read(dev(), f, wanted.len() as i32, wanted.as_ptr(), 0, cols.as_mut_ptr()) })?;   // synchronous: the copies are done
    Ok(DeviceColumns { names, cols })
}

/// DataFrame::to_arrow (src/dataframe.rs:515-525): chunk b of every column is RecordBatch b.
pub fn to_arrow_device(frame: &DeviceColumns, path: &str) -> Result<(), String> {
    let cpath = CString::new(path).unwrap();
    let cnames: Vec<CString> = frame.names.iter().map(|s| CString::new(s.as_str()).unwrap()).collect();
    let name_ptrs: Vec<*const c_char> = cnames.iter().map(|s| s.as_ptr()).collect();
    let col_ptrs: Vec<*const BdfCol> = frame.cols.iter().map(|c| *c as *const BdfCol).collect();
    check(unsafe { bdf_ipc_write(dev(), cpath.as_ptr(), col_ptrs.len() as i32, name_ptrs.as_ptr(), col_ptrs.as_ptr()) })
}

/// DataFrame::sort (src/dataframe.rs:194-222): criteria = (column index, descending); nulls last like the reference.
pub fn sort_device(frame: &DeviceColumns, criteria: &[(usize, bool)]) -> Result<DeviceColumns, String> {
    if criteria.is_empty() {
        return Err("Sort criteria cannot be empty".to_string());
    }
    let keys: Vec<BdfSortKey> = criteria.iter().map(|(c, d)| BdfSortKey { column: frame.cols[*c], descending: *d as i32 }).collect();
    let mut indices: *mut BdfCol = std::ptr::null_mut();
    check(unsafe { bdf_sort_indices_dev(dev(), keys.len() as i32, keys.as_ptr(), &mut indices) })?;   // lexsort_to_indices
    // both the index column and the columns taken so far are owned by guards: an error half-way frees them (raw pointers have no Drop)
    let idx = DeviceColumns { names: vec![], cols: vec![indices] };
    let mut out = DeviceColumns { names: frame.names.clone(), cols: Vec::with_capacity(frame.cols.len()) };
    for c in &frame.cols {
        let mut taken: *mut BdfCol = std::ptr::null_mut();
        check(unsafe { bdf_take_dev(dev(), *c, idx.cols[0], &mut taken) })?;                            // sort_by_indices -> Column::take
        out.cols.push(taken);
    }
    Ok(out)
}

/// One aggregate of the plan `Dataset::try_aggregate` makes (src/expression.rs:114-221): the ones on this path.
#[derive(Clone, Copy, PartialEq)]
pub enum GroupFn { Sum, Count, Min, Max }

/// `Transformation::GroupAggregate` (a `panic!("aggregations not supported")` in src/evaluation.rs:73): group by ONE key column,
/// output = the key column, then `sum(x)` / `min(x)` / `max(x)` (type of x) or `count(x)` (UInt32) per requested aggregate -- the
/// names and types of the planned schema.  Groups in ascending key order, the null key last.
pub fn group_aggregate_device(frame: &DeviceColumns, key: usize, aggr: &[(usize, GroupFn)]) -> Result<DeviceColumns, String> {
    const BDF_U32: i32 = 6;   // include/b200df.h bdf_dtype
    let mut value_cols: Vec<usize> = vec![];
    for (c, _) in aggr { if !value_cols.contains(c) { value_cols.push(*c); } }
    let vals: Vec<*const BdfCol> = value_cols.iter().map(|c| frame.cols[*c] as *const BdfCol).collect();
    let mut keys: *mut BdfCol = std::ptr::null_mut();
    let mut n_groups = 0i64;
    let null = std::ptr::null_mut();
    let mut outs = vec![BdfGroupOut { sum: null, count: null, min: null, max: null }; vals.len()];
    check(unsafe { bdf_group_aggregate_dev(dev(), frame.cols[key], vals.len() as i32, vals.as_ptr(), &mut keys, outs.as_mut_ptr(), &mut n_groups) })?;
    // every column the library returned is owned by a guard from here on: the ones the plan did not ask for are dropped with it
    let mut spare = DeviceColumns { names: vec![], cols: outs.iter().flat_map(|o| vec![o.sum, o.count, o.min, o.max]).collect() };
    let mut out = DeviceColumns { names: vec![frame.names[key].clone()], cols: vec![keys] };
    for (c, f) in aggr {
        let j = value_cols.iter().position(|v| v == c).unwrap();
        let slot = 4 * j + match f { GroupFn::Sum => 0, GroupFn::Count => 1, GroupFn::Min => 2, GroupFn::Max => 3 };
        let col = spare.cols[slot];
        if col.is_null() {
            return Err("min/max need T::Native: Ord (integer columns), or the aggregate was requested twice".to_string());
        }
        let x = &frame.names[*c];
        match f {
            GroupFn::Count => {   // count(x) is planned as UInt32; group sizes are below 2^32 (row numbers are UInt32), the cast cannot fail
                let mut c32: *mut BdfCol = std::ptr::null_mut();
                check(unsafe { bdf_cast_dev(dev(), BDF_U32, col, &mut c32) })?;
                out.names.push(format!("count({})", x));
                out.cols.push(c32);
            }
            _ => {
                spare.cols[slot] = null;   // moves to the result
                out.names.push(format!("{}({})", match f { GroupFn::Sum => "sum", GroupFn::Min => "min", _ => "max" }, x));
                out.cols.push(col);
            }
        }
    }
    Ok(out)
}

/// A run of Float64 Calculations whose intermediates are not kept (config 2: h = sin(((a+b)*c)/d)) as one pass.
/// `inputs` are column indices; node k writes slot inputs.len() + k (see include/b200df.h, bdf_eval_expr_dev).
pub fn fused_chain(frame: &DeviceColumns, inputs: &[usize], nodes: &[BdfExprNode]) -> Result<*mut BdfCol, String> {
    let cols: Vec<*const BdfCol> = inputs.iter().map(|i| frame.cols[*i] as *const BdfCol).collect();
    let mut out: *mut BdfCol = std::ptr::null_mut();
    check(unsafe { bdf_eval_expr_dev(dev(), cols.len() as i32, cols.as_ptr(), nodes.len() as i32, nodes.as_ptr(), &mut out) })?;
    Ok(out)
}

fn check(st: i32) -> Result<(), String> {
    if st == BDF_OK { Ok(()) } else { Err(last_error()) }
}
