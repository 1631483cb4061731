/*
 * oracle.h -- CPU restatement ("oracle") of rust-dataframe's per-RecordBatch compute path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rust-dataframe_b200/ (the product) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * What it restates (reference = nevi-me/rust-dataframe @ a8310afd, paths relative to /root/reference):
 *   src/functions/scalar.rs:16-103     ScalarFunctions::{add,subtract,multiply,par_multiply,divide}
 *   src/functions/scalar.rs:106-457    float unaries built on scalar_op (abs, sin, cos, tan, acos, ...)
 *   src/functions/scalar.rs:499-540    math_op / scalar_op (null handling, length-mismatch error text)
 *   src/functions/aggregate.rs:12-93   AggregateFunctions::{max,min,avg,count,sum}
 *   src/evaluation.rs:296-315          Function::Cast -> arrow::compute::cast per chunk
 *
 * The arithmetic itself (add/subtract/multiply/divide/sum/min/max/cast) lives in the third-party
 * `arrow` crate (Cargo.toml:9: git branch `rust-parquet-arrow-writer` of apache/arrow, features
 * ["prettyprint","simd"], ~2.0.0-SNAPSHOT; no Cargo.lock, not vendored) and in `num-traits 0.2`
 * (NumCast).  Neither source tree is on disk; their PUBLISHED semantics are restated here:
 *   - arithmetic.rs  simd_math_op / simd_divide: computes every slot (also under nulls), validity =
 *     AND of the inputs' bitmaps (absent if neither has one), integers wrap, divide returns
 *     Err(DivideByZero) if any VALID slot has a zero divisor (ints and floats), null slots divide by 1;
 *   - aggregate.rs   sum = None if all null/empty else left fold from 0 over valid slots,
 *     min/max = None if all null/empty else scan over valid slots;
 *   - cast.rs        numeric_cast: null -> null, else num::cast::cast(v): Some -> value, None -> NULL;
 *   - num-traits     ToPrimitive: int->int in-range check; float->int truncation inside the
 *     exclusive (MIN-1, MAX+1) window; anything->float always Some (`as`).
 *
 * PARITY PINNING.  Pinned by the reference's own tests (see tests/test_oracle_golden.py):
 *   abs i32/f64 (scalar.rs:565-584), acos/cos goldens (scalar.rs:587-602), count (aggregate.rs:123-127),
 *   avg two-chunk (aggregate.rs:130-146), add on the CSV fixture row 0 (dataframe.rs:803-808),
 *   par_multiply bench input (scalar.rs:623).
 * PARITY UNPINNED (no reference test pins them; the Rust reference cannot be built here -- no
 *   cargo/rustc, nightly-only crate, un-pinned git deps, no network): subtract/multiply/divide values,
 *   DivideByZero rule, sum/min/max, numeric cast values, sin/tan values, integer wrapping, sliced
 *   arrays, null-slot payload bytes, and whether compute::sum at the pinned branch folds
 *   sequentially or lane-wise (float sums are therefore compared under a tolerance, not bit-exact).
 */
#ifndef RDF_ORACLE_H
#define RDF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Primitive types in Arrow DataType order (Int8..UInt64, Float32, Float64). */
enum { ORC_I8 = 0, ORC_I16, ORC_I32, ORC_I64, ORC_U8, ORC_U16, ORC_U32, ORC_U64, ORC_F32, ORC_F64, ORC_NTYPES };

enum { ORC_ADD = 0, ORC_SUB, ORC_MUL, ORC_DIV, /* math_op binaries (floats only): */ ORC_ATAN2, ORC_HYPOT, ORC_LOG };

enum {
    ORC_ABS = 0, ORC_SIN, ORC_COS, ORC_TAN,
    ORC_ACOS, ORC_ASIN, ORC_ATAN, ORC_CBRT, ORC_CEIL, ORC_COSH, ORC_DEGREES, ORC_EXP, ORC_EXPM1,
    ORC_FLOOR, ORC_LOG10, ORC_LOG2, ORC_RADIANS, ORC_ROUND, ORC_SINH, ORC_SQRT, ORC_TANH,
    ORC_NUNARY
};

enum { ORC_SUM = 0, ORC_MIN, ORC_MAX, ORC_COUNT, ORC_MIN_AS_WRITTEN /* = max: aggregate.rs:22-31 */ };

enum { ORC_OK = 0, ORC_LENGTH_MISMATCH = 1, ORC_DIVIDE_BY_ZERO = 2, ORC_UNSUPPORTED = 3, ORC_PANIC = 7 };

/* One chunk (= one PrimitiveArray<T>) in Arrow memory layout. */
typedef struct {
    const void*    values;     /* base of the values buffer (NOT offset-adjusted)            */
    const uint8_t* validity;   /* LSB-first bitmap, 1 = valid; NULL = no nulls               */
    int64_t        len;        /* logical length                                             */
    int64_t        offset;     /* element offset applied to values and to validity bits      */
    int64_t        null_count; /* -1 = unknown (computed on demand)                          */
} orc_view;

typedef struct {
    void*    values;       /* caller-allocated, >= len * width bytes                            */
    uint8_t* validity;     /* caller-allocated, >= ceil(len/8) bytes; written with bit offset 0 */
    int64_t  len;          /* OUT */
    int64_t  null_count;   /* OUT */
    int32_t  has_validity; /* OUT: 1 if arrow-rs would attach a bitmap                          */
} orc_out;

int  orc_width(int dtype);
int64_t orc_null_count(const orc_view* v);

/* Per-chunk kernels (what arrow::compute::{add,...} / scalar_op / cast do for ONE array). */
int orc_binary(int op, int dtype, const orc_view* a, const orc_view* b, orc_out* out);
int orc_unary(int op, int dtype, const orc_view* in, orc_out* out);
int orc_cast(int from, int to, const orc_view* in, orc_out* out);

/* Column-level drivers (what ScalarFunctions::* do over Vec<&PrimitiveArray<T>>).
 * threads <= 1: chunks in order on the calling thread (subtract/multiply/divide/sin/...);
 * threads  > 1: one OpenMP task per chunk, mirroring rayon par_iter in add/par_multiply
 *               (scalar.rs:28-31, 99-102).  n = min(n_left, n_right) as zip() does. */
int orc_col_binary(int op, int dtype, int64_t n_left, const orc_view* l, int64_t n_right, const orc_view* r,
                   orc_out* out, int threads);
int orc_col_unary(int op, int dtype, int64_t n, const orc_view* in, orc_out* out, int threads);
int orc_col_cast(int from, int to, int64_t n, const orc_view* in, orc_out* out, int threads);

/* AggregateFunctions::{sum,min,max,count}.  *out holds T::Native (count: int64).  *is_some = 0 <=> None.
 * Returns ORC_PANIC where the reference would panic (max/min .unwrap() on an all-null/empty chunk). */
int orc_aggregate(int op, int dtype, int64_t n, const orc_view* chunks, void* out, int32_t* is_some);
/* AggregateFunctions::avg (aggregate.rs:32-65).  Types with f64: From<T::Native> only. */
int orc_avg(int dtype, int64_t n, const orc_view* chunks, double* out, int32_t* is_some);
/* Float-sum bounds: long-double Neumaier-compensated sum and sum of |x| over valid slots. */
int orc_sum_exact(int dtype, int64_t n, const orc_view* chunks, long double* sum, long double* sum_abs);

/* ---- the step right after the hot path -- BooleanFilter::eval_to_array + ChunkedArray::filter ----
 * Boolean arrays are bit-packed (Arrow BooleanArray): `values` is an LSB-first bitmap, offset in bits.
 *   orc_compare   BooleanFilter::{Gt,Ge,Eq,Ne,Lt,Le} (src/expression.rs:820-852): both sides are cast to Float64
 *                 (numeric_cast, `as f64`), then arrow compute::{gt,gt_eq,eq,neq,lt,lt_eq} on Float64Arrays: IEEE
 *                 comparison per slot (computed under nulls), validity = AND, absent if neither side has one.
 *                 A NULL right view with `use_scalar` broadcasts BooleanInput::Scalar (vec![v; len], :783-802).
 *   orc_bool      arrow compute::{and,or,not} on BooleanArrays: values op values, validity AND.  (The reference's
 *                 And/Or arms reinterpret Float64 buffers as bitmaps -- expression.rs:808-819 -- which is not
 *                 restated; Not casts to Boolean first.)
 *   orc_filter    arrow compute::filter(array, mask) per chunk (src/table.rs:97-107): keeps slot i iff the mask is
 *                 valid and true there; values and validity of kept slots are compacted in order.
 * No reference test pins any of this ("parity unpinned"); pyarrow cross-checks it in tests/test_oracle_golden.py. */
enum { ORC_GT = 0, ORC_GE, ORC_EQ, ORC_NE, ORC_LT, ORC_LE };
enum { ORC_AND = 0, ORC_OR, ORC_NOT };
#define ORC_BOOL 10
int orc_compare(int op, int ltype, const orc_view* l, int rtype, const orc_view* r, int use_scalar, double scalar, orc_out* out);
int orc_bool(int op, const orc_view* a, const orc_view* b, orc_out* out);
int orc_filter(int dtype, const orc_view* values, const orc_view* mask, orc_out* out);

/* ---- DataFrame::sort (src/dataframe.rs:194-222) = arrow compute::lexsort_to_indices + Column::take (src/table.rs:218-241) ----
 *   orc_lexsort_indices  every criterion column is concatenated (Column::to_array), SortOptions { descending, nulls_first:
 *                 false } -- the reference ignores SortCriteria::nulls_first (:205-208).  arrow-rs builds a
 *                 LexicographicalComparator and runs a STABLE sort (slice::sort_by) of the row numbers with it: per criterion,
 *                 (valid, valid) -> value order, reversed when descending; (null, valid) -> Greater; (valid, null) -> Less;
 *                 (null, null) -> next criterion.  Floats use cmp_nans_last: NaN == NaN, NaN greater than any number,
 *                 otherwise partial_cmp (so -0.0 == 0.0).  Restated here as a bottom-up merge sort with that comparator
 *                 (nothing shared with the product's radix sort).
 *   orc_take      arrow compute::take(values, indices): out[i] = values[indices[i]]; a null index or a null value gives a
 *                 null slot; the reference's repartitioning loop (table.rs:223-236) always yields ONE chunk.  ORC_PANIC for
 *                 an index past the end (arrow returns an error).
 * Pinned by the reference's own test_sort vector (src/dataframe.rs:963-1002) in tests/test_oracle_golden.py; pyarrow
 * cross-checks the NaN-free cases.  The arrow-rs source is not vendored: the comparator above is restated from recall. */
typedef struct { int dtype; int64_t n_chunks; const orc_view* chunks; int descending; } orc_sort_key;
int orc_lexsort_indices(int n_keys, const orc_sort_key* keys, uint32_t* out_indices);
int orc_take(int dtype, int64_t n_chunks, const orc_view* chunks, int64_t n_idx, const uint32_t* idx, const uint8_t* idx_validity,
             orc_out* out);

/* Counter-based synthetic data; the CUDA generator in the product reproduces it bit-for-bit.
 * kind 0: real uniform [lo,hi)   1: real +-[1,2)   2: integer, full range of the type
 * kind 3: integer uniform [-2^40, 2^40) (truncated to the type)
 * null_mod: 0 = no nulls, else slot is null when splitmix64(h) % null_mod == 0. */
uint64_t orc_splitmix64(uint64_t x);
void orc_generate(int dtype, int kind, double lo, double hi, uint64_t seed, uint64_t col, int64_t row0, int64_t len,
                  uint32_t null_mod, void* values, uint8_t* validity, int64_t* null_count);

#ifdef __cplusplus
}
#endif
#endif
