//! Replacement bodies for src/functions/aggregate.rs (signatures unchanged).  UNCOMPILED here -- see ../README.md.
use std::os::raw::{c_int, c_void};

use arrow::array::PrimitiveArray;
use arrow::datatypes::ArrowNumericType;

use crate::ffi::*;

/// op: 0 sum, 1 min, 2 max, 3 count (bdf_aggop).  R = T::Native for sum/min/max, i64 for count.
pub fn gpu_agg<T: ArrowNumericType, R: Default>(op: c_int, arrays: &[&PrimitiveArray<T>]) -> Option<R> {
    let v: Vec<BdfView> = arrays.iter().map(|a| view(*a)).collect();
    let (mut out, mut some) = (R::default(), 0i32);
    let st = unsafe {
        bdf_aggregate(ctx(), op, dtype_id(&T::get_data_type()), v.len() as i64, v.as_ptr(), &mut out as *mut R as *mut c_void, &mut some)
    };
    match st {
        BDF_OK => if some != 0 { Some(out) } else { None },
        // the reference calls compute::max(array).unwrap() per chunk (aggregate.rs:19,29): None panics there too
        BDF_WOULD_PANIC => panic!("called `Option::unwrap()` on a `None` value"),
        st => panic!("{}", to_arrow_error(st)),
    }
}

pub fn gpu_avg<T: ArrowNumericType>(arrays: &[&PrimitiveArray<T>]) -> Option<f64> {
    let v: Vec<BdfView> = arrays.iter().map(|a| view(*a)).collect();
    let (mut out, mut some) = (0f64, 0i32);
    let st = unsafe { bdf_avg(ctx(), dtype_id(&T::get_data_type()), v.len() as i64, v.as_ptr(), &mut out, &mut some) };
    assert_eq!(st, BDF_OK, "{}", last_error());
    if some != 0 { Some(out) } else { None }
}

// In src/functions/aggregate.rs:
//   pub fn sum<T>(arrays)   -> gpu_agg::<T, T::Native>(0, &arrays)
//   pub fn min<T>(arrays)   -> gpu_agg::<T, T::Native>(1, &arrays)   // the intended minimum; the current body is a copy of max (aggregate.rs:22-31)
//   pub fn max<T>(arrays)   -> gpu_agg::<T, T::Native>(2, &arrays)
//   pub fn count<T>(arrays) -> gpu_agg::<T, i64>(3, &arrays)         // metadata only when the null counts are known
//   pub fn avg<T>(arrays)   -> gpu_avg(&arrays)
