//! Replacement bodies for src/functions/scalar.rs (signatures unchanged).  UNCOMPILED here -- see ../README.md.
use std::os::raw::{c_int, c_void};
use std::sync::Arc;

use arrow::array::{Array, ArrayData, PrimitiveArray};
use arrow::buffer::MutableBuffer;
use arrow::datatypes::ArrowNumericType;
use arrow::error::ArrowError;

use crate::ffi::*;

/// op: 0 add, 1 subtract, 2 multiply, 3 divide, 4 atan2, 5 hypot, 6 log (bdf_binop)
pub fn gpu_binary<T: ArrowNumericType>(op: c_int, left: Vec<&PrimitiveArray<T>>, right: Vec<&PrimitiveArray<T>>)
    -> Result<Vec<PrimitiveArray<T>>, ArrowError>
{
    let n = left.len().min(right.len()); // zip()
    let lv: Vec<BdfView> = left.iter().map(|a| view(*a)).collect();
    let rv: Vec<BdfView> = right.iter().map(|a| view(*a)).collect();
    let (vals, bits, mut outs) = alloc_outputs::<T>(&left[..n].iter().map(|a| a.len()).collect::<Vec<_>>());
    let st = unsafe {
        bdf_binary(ctx(), op, dtype_id(&T::get_data_type()), lv.len() as i64, lv.as_ptr(), rv.len() as i64, rv.as_ptr(), outs.as_mut_ptr())
    };
    if st != BDF_OK { return Err(to_arrow_error(st)); }
    Ok(finish::<T>(vals, bits, &outs))
}

/// op: 0 abs, 1 sin, 2 cos, 3 tan, 4 acos, 5 asin, 6 atan, 7 cbrt, 8 ceil, 9 cosh, 10 degrees, 11 exp, 12 expm1,
/// 13 floor, 14 log10, 15 log2, 16 radians, 17 round, 18 sinh, 19 sqrt, 20 tanh (bdf_unop)
pub fn gpu_unary<T: ArrowNumericType>(op: c_int, array: Vec<&PrimitiveArray<T>>) -> Result<Vec<PrimitiveArray<T>>, ArrowError> {
    let v: Vec<BdfView> = array.iter().map(|a| view(*a)).collect();
    let (vals, bits, mut outs) = alloc_outputs::<T>(&array.iter().map(|a| a.len()).collect::<Vec<_>>());
    let st = unsafe { bdf_unary(ctx(), op, dtype_id(&T::get_data_type()), v.len() as i64, v.as_ptr(), outs.as_mut_ptr()) };
    if st != BDF_OK { return Err(to_arrow_error(st)); }
    Ok(finish::<T>(vals, bits, &outs))
}

fn alloc_outputs<T: ArrowNumericType>(lens: &[usize]) -> (Vec<MutableBuffer>, Vec<MutableBuffer>, Vec<BdfOut>) {
    let width = std::mem::size_of::<T::Native>();
    // Both buffers get their FINAL length before the call (zero-filled once, as arrow's own kernels do with `with_bitset`): the
    // library then writes into initialised memory of the right size and nothing has to be resized afterwards -- on arrow
    // versions where `MutableBuffer::resize` zero-fills the grown range, resizing AFTER the call would wipe the results.
    let mut vals: Vec<MutableBuffer> = lens.iter().map(|&n| MutableBuffer::new(n * width).with_bitset(n * width, false)).collect();
    let mut bits: Vec<MutableBuffer> = lens.iter().map(|&n| MutableBuffer::new((n + 7) / 8).with_bitset((n + 7) / 8, false)).collect();
    let outs = (0..lens.len()).map(|i| BdfOut {
        values: vals[i].raw_data_mut() as *mut c_void, validity: bits[i].raw_data_mut(), len: lens[i] as i64, null_count: 0, has_validity: 0,
    }).collect();
    (vals, bits, outs)
}

fn finish<T: ArrowNumericType>(vals: Vec<MutableBuffer>, bits: Vec<MutableBuffer>, outs: &[BdfOut]) -> Vec<PrimitiveArray<T>> {
    outs.iter().zip(vals.into_iter().zip(bits.into_iter())).map(|(o, (v, b))| {
        let len = o.len as usize;   // == the length the buffers were created with (bdf_out.len is checked by the library)
        let nulls = if o.has_validity != 0 { Some(b.freeze()) } else { None };
        let data = ArrayData::new(T::get_data_type(), len, Some(o.null_count as usize), nulls, 0, vec![v.freeze()], vec![]);
        PrimitiveArray::<T>::from(Arc::new(data))
    }).collect()
}

// In src/functions/scalar.rs the bodies become:
//   pub fn add<T>(left, right)          -> gpu_binary(0, left, right)     // was: left.par_iter().zip(..).map(compute::add)   (scalar.rs:16-32)
//   pub fn subtract<T>(left, right)     -> gpu_binary(1, left, right)     // (scalar.rs:34-50)
//   pub fn divide<T>(left, right)       -> gpu_binary(3, left, right)     // Err(ArrowError::DivideByZero) preserved (scalar.rs:51-68)
//   pub fn multiply / par_multiply<T>   -> gpu_binary(2, left, right)     // (scalar.rs:69-103)
//   pub fn abs/sin/cos/tan/acos/...<T>  -> gpu_unary(op, array)           // was: scalar_op(a, |a| Ok(num::Float::sin(a)))      (scalar.rs:106-457)
