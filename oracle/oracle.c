/*
 * oracle.c -- CPU restatement of rust-dataframe's per-RecordBatch compute path.  See oracle.h for
 * scope, provenance and the pinned / unpinned parity statement.  TEST INFRASTRUCTURE ONLY.
 *
 * Build: gcc -O3 -march=x86-64-v3 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC oracle.c -o liboracle.so -lm
 * (-ffp-contract=off: IEEE add/sub/mul/div must round once per operation, as Rust does.)
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* Arrow bitmap helpers (LSB-first, 1 = valid).                                                      */

static inline int bit_get(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
static inline void bit_set(uint8_t* b, int64_t i) { b[i >> 3] |= (uint8_t)(1u << (i & 7)); }

static inline int view_valid(const orc_view* v, int64_t i) {
    return v->validity == NULL || bit_get(v->validity, v->offset + i);
}

int orc_width(int dtype) {
    static const int w[ORC_NTYPES] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    return (dtype >= 0 && dtype < ORC_NTYPES) ? w[dtype] : 0;
}

int64_t orc_null_count(const orc_view* v) {
    if (v->validity == NULL) return 0;
    if (v->null_count >= 0) return v->null_count;
    int64_t n = 0;
    for (int64_t i = 0; i < v->len; i++) n += !bit_get(v->validity, v->offset + i);
    return n;
}

static void out_begin(orc_out* out, int64_t len, int has_validity) {
    out->len = len;
    out->null_count = 0;
    out->has_validity = has_validity;
    if (out->validity) memset(out->validity, 0, (size_t)((len + 7) / 8));
}

/* Type list: X(ENUM, ctype, unsigned-ctype-of-same-width, is_float, is_signed) */
#define ORC_INT_TYPES(X)            \
    X(ORC_I8, int8_t, uint8_t)      \
    X(ORC_I16, int16_t, uint16_t)   \
    X(ORC_I32, int32_t, uint32_t)   \
    X(ORC_I64, int64_t, uint64_t)   \
    X(ORC_U8, uint8_t, uint8_t)     \
    X(ORC_U16, uint16_t, uint16_t)  \
    X(ORC_U32, uint32_t, uint32_t)  \
    X(ORC_U64, uint64_t, uint64_t)

static inline int dtype_is_float(int t) { return t == ORC_F32 || t == ORC_F64; }
static inline int dtype_is_signed_int(int t) { return t >= ORC_I8 && t <= ORC_I64; }

/* ------------------------------------------------------------------------------------------------ */
/* Binary arithmetic: arrow-rs compute::{add,subtract,multiply,divide} as called per chunk from       */
/* ScalarFunctions::{add,subtract,multiply,par_multiply,divide} (src/functions/scalar.rs:16-103).     */
/*                                                                                                    */
/*  - length check first: Err(ComputeError("Cannot perform math operation on arrays of different     */
/*    length")) -- same text as the reference's own math_op (scalar.rs:508-511);                      */
/*  - validity = a AND b, absent when neither input has a bitmap (combine_option_bitmap);             */
/*  - add/sub/mul compute EVERY slot, also under nulls; integers wrap (SIMD path, release build);     */
/*  - divide: any VALID slot with divisor == 0 (ints and floats, +-0.0) => Err(DivideByZero) for the  */
/*    whole array; null slots use divisor 1 (their payload is a/1 = a).  iN::MIN / -1 is unspecified  */
/*    in the reference (overflow); this restatement wraps to iN::MIN.                                 */

/* validity = a AND b written at bit offset 0; returns the null count.  Byte-wise when both inputs are
 * byte-aligned (what combine_option_bitmap's buffer AND does), bit-wise for sliced arrays. */
static int64_t and_validity(const orc_view* a, const orc_view* b, uint8_t* out, int64_t n) {
    int64_t valid = 0;
    int aligned = (a->validity == NULL || (a->offset & 7) == 0) && (b->validity == NULL || (b->offset & 7) == 0);
    if (aligned) {
        const uint8_t* pa = a->validity ? a->validity + (a->offset >> 3) : NULL;
        const uint8_t* pb = b->validity ? b->validity + (b->offset >> 3) : NULL;
        int64_t nb = n >> 3;
        for (int64_t k = 0; k < nb; k++) {
            uint8_t m = (uint8_t)((pa ? pa[k] : 0xFF) & (pb ? pb[k] : 0xFF));
            out[k] = m;
            valid += __builtin_popcount(m);
        }
        for (int64_t i = nb << 3; i < n; i++)
            if (view_valid(a, i) && view_valid(b, i)) { bit_set(out, i); valid++; }
    } else {
        for (int64_t i = 0; i < n; i++)
            if (view_valid(a, i) && view_valid(b, i)) { bit_set(out, i); valid++; }
    }
    return n - valid;
}

#define DEF_INT_BINARY(ENUM, T, U)                                                                         \
    static int binary_##ENUM(int op, const orc_view* a, const orc_view* b, orc_out* out, int has_v) {    \
        const T* restrict x = (const T*)a->values + a->offset;                                             \
        const T* restrict y = (const T*)b->values + b->offset;                                             \
        T* restrict z = (T*)out->values;                                                                   \
        int64_t n = a->len;                                                                                \
        if (op == ORC_DIV) {                                                                               \
            for (int64_t i = 0; i < n; i++)                                                                \
                if (view_valid(a, i) && view_valid(b, i) && y[i] == 0) return ORC_DIVIDE_BY_ZERO;          \
        }                                                                                                  \
        if (has_v) out->null_count = and_validity(a, b, out->validity, n);                                 \
        switch (op) { /* tight loops: the reference's arrow kernels are SIMD (feature "simd") */           \
            case ORC_ADD: for (int64_t i = 0; i < n; i++) z[i] = (T)(U)((uint64_t)(U)x[i] + (uint64_t)(U)y[i]); break; \
            case ORC_SUB: for (int64_t i = 0; i < n; i++) z[i] = (T)(U)((uint64_t)(U)x[i] - (uint64_t)(U)y[i]); break; \
            case ORC_MUL: for (int64_t i = 0; i < n; i++) z[i] = (T)(U)((uint64_t)(U)x[i] * (uint64_t)(U)y[i]); break; \
            default:                                                                                       \
                for (int64_t i = 0; i < n; i++) {                                                          \
                    T d = (view_valid(a, i) && view_valid(b, i)) ? y[i] : (T)1;                            \
                    if ((T)-1 < 0 && d == (T)-1) z[i] = (T)((U)0 - (U)x[i]); /* wraps at MIN */            \
                    else z[i] = (T)(x[i] / d);                                                             \
                }                                                                                          \
        }                                                                                                  \
        return ORC_OK;                                                                                     \
    }
ORC_INT_TYPES(DEF_INT_BINARY)

#define DEF_FLT_BINARY(ENUM, T, ATAN2, HYPOT, LOGF)                                                        \
    static int binary_##ENUM(int op, const orc_view* a, const orc_view* b, orc_out* out, int has_v) {    \
        const T* restrict x = (const T*)a->values + a->offset;                                             \
        const T* restrict y = (const T*)b->values + b->offset;                                             \
        T* restrict z = (T*)out->values;                                                                   \
        int64_t n = a->len;                                                                                \
        if (op == ORC_DIV) {                                                                               \
            for (int64_t i = 0; i < n; i++)                                                                \
                if (view_valid(a, i) && view_valid(b, i) && y[i] == (T)0) return ORC_DIVIDE_BY_ZERO;       \
        }                                                                                                  \
        if (has_v) out->null_count = and_validity(a, b, out->validity, n);                                 \
        switch (op) {                                                                                      \
            case ORC_ADD: for (int64_t i = 0; i < n; i++) z[i] = x[i] + y[i]; break;                       \
            case ORC_SUB: for (int64_t i = 0; i < n; i++) z[i] = x[i] - y[i]; break;                       \
            case ORC_MUL: for (int64_t i = 0; i < n; i++) z[i] = x[i] * y[i]; break;                       \
            case ORC_DIV:                                                                                  \
                for (int64_t i = 0; i < n; i++)                                                            \
                    z[i] = x[i] / ((view_valid(a, i) && view_valid(b, i)) ? y[i] : (T)1);                  \
                break;                                                                                     \
            default: /* math_op (scalar.rs:499-523): builder loop, null -> append_null (payload 0) */      \
                for (int64_t i = 0; i < n; i++) {                                                          \
                    int valid = view_valid(a, i) && view_valid(b, i);                                      \
                    if (op == ORC_ATAN2) z[i] = valid ? ATAN2(x[i], y[i]) : (T)0;                          \
                    else if (op == ORC_HYPOT) z[i] = valid ? HYPOT(x[i], y[i]) : (T)0;                     \
                    else z[i] = valid ? (T)(LOGF(x[i]) / LOGF(y[i])) : (T)0; /* ln(x)/ln(base) */          \
                }                                                                                          \
        }                                                                                                  \
        return ORC_OK;                                                                                     \
    }
DEF_FLT_BINARY(ORC_F32, float, atan2f, hypotf, logf)
DEF_FLT_BINARY(ORC_F64, double, atan2, hypot, log)

int orc_binary(int op, int dtype, const orc_view* a, const orc_view* b, orc_out* out) {
    if (a->len != b->len) return ORC_LENGTH_MISMATCH;
    if (op < ORC_ADD || op > ORC_LOG) return ORC_UNSUPPORTED;
    if (op > ORC_DIV && !dtype_is_float(dtype)) return ORC_UNSUPPORTED; /* T::Native: Float */
    /* math_op always goes through a builder => bitmap always present; arrow kernels: only if an input has one */
    int has_v = (op > ORC_DIV) ? 1 : (a->validity != NULL || b->validity != NULL);
    out_begin(out, a->len, has_v);
    int st;
    switch (dtype) {
#define CASE_BIN(ENUM, T, U) case ENUM: st = binary_##ENUM(op, a, b, out, has_v); break;
        ORC_INT_TYPES(CASE_BIN)
        case ORC_F32: st = binary_ORC_F32(op, a, b, out, has_v); break;
        case ORC_F64: st = binary_ORC_F64(op, a, b, out, has_v); break;
        default: st = ORC_UNSUPPORTED;
    }
    return st;
}

/* ------------------------------------------------------------------------------------------------ */
/* Unary: scalar_op (src/functions/scalar.rs:525-540): for i in 0..len: null -> append_null           */
/* (payload 0), else append_value(op(x)).  The op is num::Float::* = Rust std = platform libm.        */
/* abs = num::abs (Signed): wraps at iN::MIN in release.  Output always carries a bitmap (builder).   */

static const double PI_F64 = 3.14159265358979323846264338327950288;

static double unary_f64(int op, double x) {
    switch (op) {
        case ORC_ABS: return fabs(x);
        case ORC_SIN: return sin(x);
        case ORC_COS: return cos(x);
        case ORC_TAN: return tan(x);
        case ORC_ACOS: return acos(x);
        case ORC_ASIN: return asin(x);
        case ORC_ATAN: return atan(x);
        case ORC_CBRT: return cbrt(x);
        case ORC_CEIL: return ceil(x);
        case ORC_COSH: return cosh(x);
        case ORC_DEGREES: return x * (180.0 / PI_F64);           /* f64::to_degrees */
        case ORC_EXP: return exp(x);
        case ORC_EXPM1: return expm1(x);
        case ORC_FLOOR: return floor(x);
        case ORC_LOG10: return log10(x);
        case ORC_LOG2: return log2(x);
        case ORC_RADIANS: return x * (PI_F64 / 180.0);           /* f64::to_radians */
        case ORC_ROUND: return round(x);                         /* half away from zero, like f64::round */
        case ORC_SINH: return sinh(x);
        case ORC_SQRT: return sqrt(x);
        default: return tanh(x);
    }
}

static float unary_f32(int op, float x) {
    switch (op) {
        case ORC_ABS: return fabsf(x);
        case ORC_SIN: return sinf(x);
        case ORC_COS: return cosf(x);
        case ORC_TAN: return tanf(x);
        case ORC_ACOS: return acosf(x);
        case ORC_ASIN: return asinf(x);
        case ORC_ATAN: return atanf(x);
        case ORC_CBRT: return cbrtf(x);
        case ORC_CEIL: return ceilf(x);
        case ORC_COSH: return coshf(x);
        case ORC_DEGREES: return x * 57.2957795130823208767981548141051703f; /* f32::to_degrees constant */
        case ORC_EXP: return expf(x);
        case ORC_EXPM1: return expm1f(x);
        case ORC_FLOOR: return floorf(x);
        case ORC_LOG10: return log10f(x);
        case ORC_LOG2: return log2f(x);
        case ORC_RADIANS: return x * (3.14159265358979323846264338327950288f / 180.0f);
        case ORC_ROUND: return roundf(x);
        case ORC_SINH: return sinhf(x);
        case ORC_SQRT: return sqrtf(x);
        default: return tanhf(x);
    }
}

int orc_unary(int op, int dtype, const orc_view* in, orc_out* out) {
    if (op < 0 || op >= ORC_NUNARY) return ORC_UNSUPPORTED;
    if (op == ORC_ABS) {
        if (!dtype_is_float(dtype) && !dtype_is_signed_int(dtype)) return ORC_UNSUPPORTED; /* T::Native: Signed */
    } else if (!dtype_is_float(dtype)) {
        return ORC_UNSUPPORTED; /* T::Native: Float */
    }
    int64_t n = in->len;
    out_begin(out, n, 1);
    for (int64_t i = 0; i < n; i++) {
        int valid = view_valid(in, i);
        if (valid) bit_set(out->validity, i); else out->null_count++;
        int64_t j = in->offset + i;
        switch (dtype) {
            case ORC_F64: ((double*)out->values)[i] = valid ? unary_f64(op, ((const double*)in->values)[j]) : 0.0; break;
            case ORC_F32: ((float*)out->values)[i] = valid ? unary_f32(op, ((const float*)in->values)[j]) : 0.0f; break;
#define CASE_ABS(ENUM, T, U)                                                                   \
            case ENUM: {                                                                       \
                T v = ((const T*)in->values)[j];                                               \
                ((T*)out->values)[i] = valid ? (T)(v < 0 ? (T)((U)0 - (U)v) : v) : (T)0;       \
            } break;
            CASE_ABS(ORC_I8, int8_t, uint8_t)
            CASE_ABS(ORC_I16, int16_t, uint16_t)
            CASE_ABS(ORC_I32, int32_t, uint32_t)
            CASE_ABS(ORC_I64, int64_t, uint64_t)
            default: return ORC_UNSUPPORTED;
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* Cast: Function::Cast (src/evaluation.rs:296-315) -> arrow::compute::cast -> numeric_cast:          */
/* per element: null -> null; else num::cast::cast::<From,To>(v): Some(v') -> v', None -> NULL.       */
/* num-traits 0.2 ToPrimitive rules:                                                                  */
/*   int -> int   : Some iff the value is representable in the destination;                           */
/*   int -> float : always Some(v as f) (round-to-nearest-even);                                      */
/*   float -> float: always Some(v as f) (f64->f32 rounds to nearest, overflow -> +-inf);             */
/*   float -> signed iN : if size_of(f) > size_of(iN): Some iff MIN-1 < v < MAX+1 (both exact);       */
/*                        else Some iff (iN::MIN as f) <= v < (iN::MAX as f) [= 2^(bits-1)];          */
/*   float -> unsigned uN: Some iff -1 < v < MAX+1 (or < (uN::MAX as f) = 2^bits when f is not wider) */
/*   NaN/+-inf -> None.  In-window values truncate toward zero.                                       */

typedef struct { int ok; int64_t i; uint64_t u; double f; } num_any;  /* value carried as i (signed), u (unsigned) or f */

static inline int int_bits(int t) { return orc_width(t) * 8; }

static int float_to_int_ok(int from, int to, double v) {
    int fbytes = orc_width(from), ibytes = orc_width(to), bits = ibytes * 8;
    if (v != v) return 0;
    if (dtype_is_signed_int(to)) {
        double max_p1 = ldexp(1.0, bits - 1);  /* 2^(bits-1), exact in f32 and f64 */
        if (fbytes > ibytes) return v > -max_p1 - 1.0 && v < max_p1; /* MIN-1 exact for these pairs */
        return v >= -max_p1 && v < max_p1;
    } else {
        double max_p1 = ldexp(1.0, bits);
        return v > -1.0 && v < max_p1;
    }
}

#define LOAD_AS(dst_i, dst_u, dst_f, ENUMV, p, j)                 \
    switch (ENUMV) {                                              \
        case ORC_I8: dst_i = ((const int8_t*)p)[j]; break;        \
        case ORC_I16: dst_i = ((const int16_t*)p)[j]; break;      \
        case ORC_I32: dst_i = ((const int32_t*)p)[j]; break;      \
        case ORC_I64: dst_i = ((const int64_t*)p)[j]; break;      \
        case ORC_U8: dst_u = ((const uint8_t*)p)[j]; break;       \
        case ORC_U16: dst_u = ((const uint16_t*)p)[j]; break;     \
        case ORC_U32: dst_u = ((const uint32_t*)p)[j]; break;     \
        case ORC_U64: dst_u = ((const uint64_t*)p)[j]; break;     \
        case ORC_F32: dst_f = ((const float*)p)[j]; break;        \
        default: dst_f = ((const double*)p)[j]; break;            \
    }

static const int64_t INT_MIN_OF[4] = {INT8_MIN, INT16_MIN, INT32_MIN, INT64_MIN};
static const int64_t INT_MAX_OF[4] = {INT8_MAX, INT16_MAX, INT32_MAX, INT64_MAX};
static const uint64_t UINT_MAX_OF[4] = {UINT8_MAX, UINT16_MAX, UINT32_MAX, UINT64_MAX};

int orc_cast(int from, int to, const orc_view* in, orc_out* out) {
    if (from < 0 || from >= ORC_NTYPES || to < 0 || to >= ORC_NTYPES) return ORC_UNSUPPORTED;
    int64_t n = in->len;
    int wt = orc_width(to);
    if (from == to) { /* (from == to) => array.clone(): same logical content */
        out_begin(out, n, in->validity != NULL);
        memcpy(out->values, (const char*)in->values + in->offset * wt, (size_t)(n * wt));
        if (in->validity)
            for (int64_t i = 0; i < n; i++) { if (view_valid(in, i)) bit_set(out->validity, i); else out->null_count++; }
        return ORC_OK;
    }
    out_begin(out, n, 1); /* builder => bitmap always present */
    int from_signed = dtype_is_signed_int(from), from_float = dtype_is_float(from);
    int to_signed = dtype_is_signed_int(to), to_float = dtype_is_float(to);
    for (int64_t i = 0; i < n; i++) {
        int64_t j = in->offset + i;
        int ok = view_valid(in, i);
        int64_t si = 0; uint64_t ui = 0; double fv = 0.0;
        LOAD_AS(si, ui, fv, from, in->values, j)
        int64_t ri = 0; uint64_t ru = 0;
        if (ok && !to_float) {
            if (from_float) {
                ok = float_to_int_ok(from, to, fv);
                if (ok) { if (to_signed) ri = (int64_t)fv; else ru = (uint64_t)fv; }
            } else if (from_signed) {
                if (to_signed) { ok = si >= INT_MIN_OF[to - ORC_I8] && si <= INT_MAX_OF[to - ORC_I8]; ri = si; }
                else { ok = si >= 0 && (uint64_t)si <= UINT_MAX_OF[to - ORC_U8]; ru = (uint64_t)si; }
            } else {
                if (to_signed) { ok = ui <= (uint64_t)INT_MAX_OF[to - ORC_I8]; ri = (int64_t)ui; }
                else { ok = ui <= UINT_MAX_OF[to - ORC_U8]; ru = ui; }
            }
        }
        if (ok) bit_set(out->validity, i); else out->null_count++;
        switch (to) {
            case ORC_I8: ((int8_t*)out->values)[i] = ok ? (int8_t)ri : 0; break;
            case ORC_I16: ((int16_t*)out->values)[i] = ok ? (int16_t)ri : 0; break;
            case ORC_I32: ((int32_t*)out->values)[i] = ok ? (int32_t)ri : 0; break;
            case ORC_I64: ((int64_t*)out->values)[i] = ok ? ri : 0; break;
            case ORC_U8: ((uint8_t*)out->values)[i] = ok ? (uint8_t)ru : 0; break;
            case ORC_U16: ((uint16_t*)out->values)[i] = ok ? (uint16_t)ru : 0; break;
            case ORC_U32: ((uint32_t*)out->values)[i] = ok ? (uint32_t)ru : 0; break;
            case ORC_U64: ((uint64_t*)out->values)[i] = ok ? ru : 0; break;
            case ORC_F32: {
                float r = 0.0f;
                if (ok) r = from_float ? (float)fv : (from_signed ? (float)si : (float)ui);
                ((float*)out->values)[i] = r;
            } break;
            default: {
                double r = 0.0;
                if (ok) r = from_float ? fv : (from_signed ? (double)si : (double)ui);
                ((double*)out->values)[i] = r;
            }
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* Column-level drivers: zip over chunks (scalar.rs:28-31 etc.).  First Err wins (collect into Result). */

int orc_col_binary(int op, int dtype, int64_t n_left, const orc_view* l, int64_t n_right, const orc_view* r,
                   orc_out* out, int threads) {
    int64_t n = n_left < n_right ? n_left : n_right; /* zip truncates to the shorter Vec */
    int status = ORC_OK;
    if (threads > 1) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int64_t c = 0; c < n; c++) {
            int st = orc_binary(op, dtype, &l[c], &r[c], &out[c]);
            if (st != ORC_OK) {
#pragma omp critical
                if (status == ORC_OK) status = st;
            }
        }
    } else {
        for (int64_t c = 0; c < n; c++) {
            int st = orc_binary(op, dtype, &l[c], &r[c], &out[c]);
            if (st != ORC_OK) return st;
        }
    }
    return status;
}

int orc_col_unary(int op, int dtype, int64_t n, const orc_view* in, orc_out* out, int threads) {
    int status = ORC_OK;
    if (threads > 1) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int64_t c = 0; c < n; c++) {
            int st = orc_unary(op, dtype, &in[c], &out[c]);
            if (st != ORC_OK) {
#pragma omp critical
                if (status == ORC_OK) status = st;
            }
        }
    } else {
        for (int64_t c = 0; c < n; c++) {
            int st = orc_unary(op, dtype, &in[c], &out[c]);
            if (st != ORC_OK) return st;
        }
    }
    return status;
}

int orc_col_cast(int from, int to, int64_t n, const orc_view* in, orc_out* out, int threads) {
    int status = ORC_OK;
    if (threads > 1) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int64_t c = 0; c < n; c++) {
            int st = orc_cast(from, to, &in[c], &out[c]);
            if (st != ORC_OK) {
#pragma omp critical
                if (status == ORC_OK) status = st;
            }
        }
    } else {
        for (int64_t c = 0; c < n; c++) {
            int st = orc_cast(from, to, &in[c], &out[c]);
            if (st != ORC_OK) return st;
        }
    }
    return status;
}

/* ------------------------------------------------------------------------------------------------ */
/* Aggregates (src/functions/aggregate.rs).                                                           */
/*  sum  :82-93  sum = T::default(); for chunk in order: sum = sum + compute::sum(chunk).unwrap_or(0) */
/*               compute::sum: None if null_count == len (incl. empty), else left fold from 0 over    */
/*               valid slots in index order.  Integer + wraps (release).  Result always Some.         */
/*  max  :12-21  per chunk compute::max(chunk).unwrap() (PANICS on all-null/empty chunk), then        */
/*               Iterator::max over chunks; None only for an empty Vec.  T::Native: Ord => ints only. */
/*  min  :22-31  AS WRITTEN identical to max (bug).  ORC_MIN = the intended min (what                 */
/*               arrow compute::min does); ORC_MIN_AS_WRITTEN = max.                                  */
/*  count:70-80  sum over chunks of len - null_count, as i64; always Some.                            */

#define DEF_INT_AGG(ENUM, T, U)                                                                     \
    static int agg_##ENUM(int op, int64_t n, const orc_view* ch, void* out, int32_t* is_some) {    \
        if (op == ORC_SUM) {                                                                        \
            U total = 0;                                                                            \
            for (int64_t c = 0; c < n; c++) {                                                       \
                const T* x = (const T*)ch[c].values + ch[c].offset;                                 \
                U s = 0;                                                                            \
                for (int64_t i = 0; i < ch[c].len; i++) if (view_valid(&ch[c], i)) s = (U)(s + (U)x[i]); \
                total = (U)(total + s);                                                             \
            }                                                                                       \
            *(T*)out = (T)total; *is_some = 1; return ORC_OK;                                       \
        }                                                                                           \
        int want_max = (op != ORC_MIN);                                                             \
        int have = 0; T best = 0;                                                                   \
        for (int64_t c = 0; c < n; c++) {                                                           \
            const T* x = (const T*)ch[c].values + ch[c].offset;                                     \
            int chave = 0; T cbest = 0;                                                             \
            for (int64_t i = 0; i < ch[c].len; i++) {                                               \
                if (!view_valid(&ch[c], i)) continue;                                               \
                if (!chave || (want_max ? x[i] > cbest : x[i] < cbest)) { cbest = x[i]; chave = 1; } \
            }                                                                                       \
            if (!chave) return ORC_PANIC; /* compute::max(..).unwrap() on None */                   \
            if (!have || (want_max ? cbest > best : cbest < best)) { best = cbest; have = 1; }      \
        }                                                                                           \
        *is_some = have; if (have) *(T*)out = best;                                                 \
        return ORC_OK;                                                                              \
    }
ORC_INT_TYPES(DEF_INT_AGG)

#define DEF_FLT_SUM(ENUM, T)                                                                        \
    static int agg_##ENUM(int op, int64_t n, const orc_view* ch, void* out, int32_t* is_some) {    \
        if (op != ORC_SUM) return ORC_UNSUPPORTED; /* T::Native: Ord does not hold for floats */    \
        T total = (T)0;                                                                             \
        for (int64_t c = 0; c < n; c++) {                                                           \
            const T* x = (const T*)ch[c].values + ch[c].offset;                                     \
            T s = (T)0; int any = 0;                                                                \
            for (int64_t i = 0; i < ch[c].len; i++) if (view_valid(&ch[c], i)) { s = s + x[i]; any = 1; } \
            total = total + (any ? s : (T)0);                                                       \
        }                                                                                           \
        *(T*)out = total; *is_some = 1; return ORC_OK;                                              \
    }
DEF_FLT_SUM(ORC_F32, float)
DEF_FLT_SUM(ORC_F64, double)

int orc_aggregate(int op, int dtype, int64_t n, const orc_view* chunks, void* out, int32_t* is_some) {
    if (op == ORC_COUNT) {
        int64_t total = 0;
        for (int64_t c = 0; c < n; c++) total += chunks[c].len - orc_null_count(&chunks[c]);
        *(int64_t*)out = total; *is_some = 1;
        return ORC_OK;
    }
    if (op != ORC_SUM && op != ORC_MIN && op != ORC_MAX && op != ORC_MIN_AS_WRITTEN) return ORC_UNSUPPORTED;
    switch (dtype) {
#define CASE_AGG(ENUM, T, U) case ENUM: return agg_##ENUM(op, n, chunks, out, is_some);
        ORC_INT_TYPES(CASE_AGG)
        case ORC_F32: return agg_ORC_F32(op, n, chunks, out, is_some);
        case ORC_F64: return agg_ORC_F64(op, n, chunks, out, is_some);
        default: return ORC_UNSUPPORTED;
    }
}

static double load_as_f64(int dtype, const void* p, int64_t j) {
    int64_t si = 0; uint64_t ui = 0; double f = 0;
    LOAD_AS(si, ui, f, dtype, p, j)
    if (dtype_is_float(dtype)) return f;
    return dtype_is_signed_int(dtype) ? (double)si : (double)ui;
}

/* avg (aggregate.rs:32-65): per chunk running mean m += (x - m)/(i + 1 - nulls), then weighted merge
 * mean += (m - mean)*len/count in chunk order.  None iff no valid value.  f64: From<T::Native> excludes
 * Int64/UInt64. */
int orc_avg(int dtype, int64_t n, const orc_view* chunks, double* out, int32_t* is_some) {
    if (dtype == ORC_I64 || dtype == ORC_U64 || dtype < 0 || dtype >= ORC_NTYPES) return ORC_UNSUPPORTED;
    double mean = 0.0; int64_t count = 0;
    for (int64_t c = 0; c < n; c++) {
        double m = 0.0; int64_t nulls = 0;
        for (int64_t i = 0; i < chunks[c].len; i++) {
            if (view_valid(&chunks[c], i)) {
                double x = load_as_f64(dtype, chunks[c].values, chunks[c].offset + i);
                m = m + (x - m) / (double)(i + 1 - nulls);
            } else nulls++;
        }
        int64_t len = chunks[c].len - nulls;
        count += len;
        mean = mean + ((m - mean) * (double)len) / (double)count; /* 0/0 = NaN if count == 0, as in Rust */
    }
    *is_some = count != 0;
    *out = mean;
    return ORC_OK;
}

int orc_sum_exact(int dtype, int64_t n, const orc_view* chunks, long double* sum, long double* sum_abs) {
    if (!dtype_is_float(dtype)) return ORC_UNSUPPORTED;
    long double s = 0, comp = 0, sa = 0;
    for (int64_t c = 0; c < n; c++)
        for (int64_t i = 0; i < chunks[c].len; i++) {
            if (!view_valid(&chunks[c], i)) continue;
            long double x = (long double)load_as_f64(dtype, chunks[c].values, chunks[c].offset + i);
            long double t = s + x;
            if (fabsl(s) >= fabsl(x)) comp += (s - t) + x; else comp += (x - t) + s;
            s = t; sa += fabsl(x);
        }
    *sum = s + comp; *sum_abs = sa;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* N2: BooleanFilter comparisons, boolean kernels, filter (see oracle.h).                             */

int orc_compare(int op, int ltype, const orc_view* l, int rtype, const orc_view* r, int use_scalar, double scalar, orc_out* out) {
    if (op < ORC_GT || op > ORC_LE || ltype < 0 || ltype >= ORC_NTYPES) return ORC_UNSUPPORTED;
    if (!use_scalar && (r == NULL || rtype < 0 || rtype >= ORC_NTYPES)) return ORC_UNSUPPORTED;
    if (!use_scalar && l->len != r->len) return ORC_LENGTH_MISMATCH;
    const int64_t n = l->len;
    const int has_v = l->validity != NULL || (!use_scalar && r->validity != NULL);
    out_begin(out, n, has_v);
    memset(out->values, 0, (size_t)((n + 7) / 8));
    uint8_t* bits = (uint8_t*)out->values;
    for (int64_t i = 0; i < n; i++) {
        /* cast(.., Float64): numeric_cast appends NULL (payload 0) for null slots; a Float64 input is cloned */
        const double a = (ltype != ORC_F64 && !view_valid(l, i)) ? 0.0 : load_as_f64(ltype, l->values, l->offset + i);
        const double b = use_scalar ? scalar
                         : (rtype != ORC_F64 && !view_valid(r, i)) ? 0.0 : load_as_f64(rtype, r->values, r->offset + i);
        int t;
        switch (op) {
            case ORC_GT: t = a > b; break;
            case ORC_GE: t = a >= b; break;
            case ORC_EQ: t = a == b; break;
            case ORC_NE: t = a != b; break;
            case ORC_LT: t = a < b; break;
            default: t = a <= b; break;
        }
        if (t) bit_set(bits, i);
        if (has_v) {
            const int valid = view_valid(l, i) && (use_scalar || view_valid(r, i));
            if (valid) bit_set(out->validity, i); else out->null_count++;
        }
    }
    return ORC_OK;
}

int orc_bool(int op, const orc_view* a, const orc_view* b, orc_out* out) {
    if (op < ORC_AND || op > ORC_NOT) return ORC_UNSUPPORTED;
    if (op != ORC_NOT && a->len != b->len) return ORC_LENGTH_MISMATCH;
    const int64_t n = a->len;
    const int has_v = a->validity != NULL || (op != ORC_NOT && b->validity != NULL);
    out_begin(out, n, has_v);
    memset(out->values, 0, (size_t)((n + 7) / 8));
    for (int64_t i = 0; i < n; i++) {
        const int x = bit_get((const uint8_t*)a->values, a->offset + i);
        const int y = op == ORC_NOT ? 0 : bit_get((const uint8_t*)b->values, b->offset + i);
        const int t = op == ORC_AND ? (x & y) : op == ORC_OR ? (x | y) : !x;
        if (t) bit_set((uint8_t*)out->values, i);
        if (has_v) {
            const int valid = view_valid(a, i) && (op == ORC_NOT || view_valid(b, i));
            if (valid) bit_set(out->validity, i); else out->null_count++;
        }
    }
    return ORC_OK;
}

int orc_filter(int dtype, const orc_view* values, const orc_view* mask, orc_out* out) {
    if (values->len != mask->len) return ORC_LENGTH_MISMATCH;
    const int w = dtype == ORC_BOOL ? 0 : orc_width(dtype);
    if (dtype != ORC_BOOL && w == 0) return ORC_UNSUPPORTED;
    const int64_t n = values->len;
    int64_t k = 0;
    out->null_count = 0;
    out->has_validity = values->validity != NULL;
    if (out->validity) memset(out->validity, 0, (size_t)((n + 7) / 8));
    if (dtype == ORC_BOOL) memset(out->values, 0, (size_t)((n + 7) / 8));
    for (int64_t i = 0; i < n; i++) {
        if (!view_valid(mask, i) || !bit_get((const uint8_t*)mask->values, mask->offset + i)) continue;  /* null mask = false */
        if (dtype == ORC_BOOL) {
            if (bit_get((const uint8_t*)values->values, values->offset + i)) bit_set((uint8_t*)out->values, k);
        } else {
            memcpy((char*)out->values + k * w, (const char*)values->values + (values->offset + i) * w, (size_t)w);
        }
        if (values->validity) {
            if (view_valid(values, i)) bit_set(out->validity, k); else out->null_count++;
        }
        k++;
    }
    out->len = k;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* Counter-based generator.                                                             */

uint64_t orc_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_generate(int dtype, int kind, double lo, double hi, uint64_t seed, uint64_t col, int64_t row0, int64_t len,
                  uint32_t null_mod, void* values, uint8_t* validity, int64_t* null_count) {
    int64_t nulls = 0;
    if (validity) memset(validity, 0, (size_t)((len + 7) / 8));
    for (int64_t i = 0; i < len; i++) {
        uint64_t h = orc_splitmix64(seed ^ (col << 56) ^ (uint64_t)(row0 + i));
        double f = 0.0; int64_t iv = 0; int is_real = kind <= 1;
        if (kind == 0) {
            double u = (double)(h >> 11) * 0x1.0p-53;
            double scaled = (hi - lo) * u;
            f = lo + scaled;
        } else if (kind == 1) {
            double u = (double)(h >> 11) * 0x1.0p-53;
            f = 1.0 + u;
            if (h & 1) f = -f;
        } else if (kind == 2) {
            iv = (int64_t)h;
        } else {
            iv = (int64_t)(h >> 23) - ((int64_t)1 << 40);
        }
        switch (dtype) {
            case ORC_F64: ((double*)values)[i] = is_real ? f : (double)iv; break;
            case ORC_F32: ((float*)values)[i] = is_real ? (float)f : (float)iv; break;
            default: {
                int64_t v = is_real ? (int64_t)f : iv;
                switch (orc_width(dtype)) {
                    case 1: ((uint8_t*)values)[i] = (uint8_t)v; break;
                    case 2: ((uint16_t*)values)[i] = (uint16_t)v; break;
                    case 4: ((uint32_t*)values)[i] = (uint32_t)v; break;
                    default: ((uint64_t*)values)[i] = (uint64_t)v; break;
                }
            }
        }
        if (validity) {
            int valid = null_mod == 0 || (orc_splitmix64(h) % null_mod) != 0;
            if (valid) bit_set(validity, i); else nulls++;
        }
    }
    if (null_count) *null_count = nulls;
}

/* ---- DataFrame::sort: lexsort_to_indices + take (src/dataframe.rs:194-222, src/table.rs:218-241) ------------------ */
typedef struct {
    int dtype, descending;
    const void* flat;       /* concatenated values (Column::to_array) */
    const uint8_t* valid;   /* one byte per row */
} sort_col;

static int cmp_values(const sort_col* k, uint32_t a, uint32_t b) {
    switch (k->dtype) {
#define ORC_CMP_INT(ID, T) case ID: { T x = ((const T*)k->flat)[a], y = ((const T*)k->flat)[b]; return x < y ? -1 : (x > y ? 1 : 0); }
        ORC_CMP_INT(ORC_I8, int8_t) ORC_CMP_INT(ORC_I16, int16_t) ORC_CMP_INT(ORC_I32, int32_t) ORC_CMP_INT(ORC_I64, int64_t)
        ORC_CMP_INT(ORC_U8, uint8_t) ORC_CMP_INT(ORC_U16, uint16_t) ORC_CMP_INT(ORC_U32, uint32_t) ORC_CMP_INT(ORC_U64, uint64_t)
#undef ORC_CMP_INT
        case ORC_F32: {   /* cmp_nans_last */
            float x = ((const float*)k->flat)[a], y = ((const float*)k->flat)[b];
            if (x != x) return y != y ? 0 : 1;
            if (y != y) return -1;
            return x < y ? -1 : (x > y ? 1 : 0);
        }
        default: {
            double x = ((const double*)k->flat)[a], y = ((const double*)k->flat)[b];
            if (x != x) return y != y ? 0 : 1;
            if (y != y) return -1;
            return x < y ? -1 : (x > y ? 1 : 0);
        }
    }
}

/* LexicographicalComparator::compare with nulls_first == false */
static int cmp_rows(int n_keys, const sort_col* k, uint32_t a, uint32_t b) {
    for (int i = 0; i < n_keys; i++) {
        const int va = k[i].valid[a], vb = k[i].valid[b];
        if (va && vb) {
            int c = cmp_values(&k[i], a, b);
            if (c) return k[i].descending ? -c : c;
        } else if (va != vb) {
            return va ? -1 : 1;
        }
    }
    return 0;
}

int orc_lexsort_indices(int n_keys, const orc_sort_key* keys, uint32_t* out) {
    if (n_keys < 1) return ORC_UNSUPPORTED;   /* DataFrameError::ComputeError("Sort criteria cannot be empty") */
    int64_t n = -1;
    for (int i = 0; i < n_keys; i++) {
        int64_t t = 0;
        for (int64_t c = 0; c < keys[i].n_chunks; c++) t += keys[i].chunks[c].len;
        if (n < 0) n = t;
        if (t != n) return ORC_LENGTH_MISMATCH;
        if (keys[i].dtype < 0 || keys[i].dtype > ORC_F64) return ORC_UNSUPPORTED;
    }
    sort_col* k = (sort_col*)calloc((size_t)n_keys, sizeof(sort_col));
    for (int i = 0; i < n_keys; i++) {
        const int w = orc_width(keys[i].dtype);
        char* flat = (char*)malloc((size_t)(n > 0 ? n : 1) * (size_t)w);
        uint8_t* valid = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
        int64_t at = 0;
        for (int64_t c = 0; c < keys[i].n_chunks; c++) {
            const orc_view* v = &keys[i].chunks[c];
            if (v->len) memcpy(flat + at * w, (const char*)v->values + v->offset * w, (size_t)v->len * (size_t)w);
            for (int64_t j = 0; j < v->len; j++) valid[at + j] = v->validity ? (uint8_t)bit_get(v->validity, v->offset + j) : 1;
            at += v->len;
        }
        k[i].dtype = keys[i].dtype; k[i].descending = keys[i].descending; k[i].flat = flat; k[i].valid = valid;
    }
    uint32_t* tmp = (uint32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(uint32_t));
    for (int64_t i = 0; i < n; i++) out[i] = (uint32_t)i;
    uint32_t *src = out, *dst = tmp;
    for (int64_t width = 1; width < n; width *= 2) {   /* bottom-up stable merge sort */
        for (int64_t lo = 0; lo < n; lo += 2 * width) {
            int64_t mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) dst[o++] = cmp_rows(n_keys, k, src[b], src[a]) < 0 ? src[b++] : src[a++];   /* ties: left first */
            while (a < mid) dst[o++] = src[a++];
            while (b < hi) dst[o++] = src[b++];
        }
        uint32_t* t = src; src = dst; dst = t;
    }
    if (src != out) memcpy(out, src, (size_t)n * sizeof(uint32_t));
    free(tmp);
    for (int i = 0; i < n_keys; i++) { free((void*)k[i].flat); free((void*)k[i].valid); }
    free(k);
    return ORC_OK;
}

int orc_take(int dtype, int64_t n_chunks, const orc_view* chunks, int64_t n_idx, const uint32_t* idx, const uint8_t* idx_validity, orc_out* out) {
    const int is_bool = dtype == ORC_BOOL;
    if (!is_bool && (dtype < 0 || dtype > ORC_F64)) return ORC_UNSUPPORTED;
    const int w = is_bool ? 1 : orc_width(dtype);
    int64_t total = 0;
    int any_validity = idx_validity != NULL;
    for (int64_t c = 0; c < n_chunks; c++) { total += chunks[c].len; any_validity |= chunks[c].validity != NULL; }
    out->len = n_idx; out->null_count = 0; out->has_validity = any_validity;
    if (is_bool) memset(out->values, 0, (size_t)((n_idx + 7) / 8));
    if (any_validity) memset(out->validity, 0, (size_t)((n_idx + 7) / 8));
    for (int64_t i = 0; i < n_idx; i++) {
        int valid = idx_validity ? bit_get(idx_validity, i) : 1;
        if (valid) {
            int64_t r = idx[i];
            if (r >= total) return ORC_PANIC;
            int64_t c = 0;
            while (r >= chunks[c].len) { r -= chunks[c].len; c++; }
            const orc_view* v = &chunks[c];
            valid = v->validity ? bit_get(v->validity, v->offset + r) : 1;
            if (valid) {
                if (is_bool) { if (bit_get((const uint8_t*)v->values, v->offset + r)) bit_set((uint8_t*)out->values, i); }
                else memcpy((char*)out->values + i * w, (const char*)v->values + (v->offset + r) * w, (size_t)w);
            }
        }
        if (!valid && !is_bool) memset((char*)out->values + i * w, 0, (size_t)w);
        if (any_validity) { if (valid) bit_set(out->validity, i); else out->null_count++; }
    }
    return ORC_OK;
}
