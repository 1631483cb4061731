// k_binary.cu -- K1: c[i] = a[i] (op) b[i] over all chunks of a column pair in ONE launch, with the
// validity AND, the output null count and the divide-by-zero flag fused in.
//
// Replaces arrow::compute::{add,subtract,multiply,divide} as called per chunk by
// ScalarFunctions::{add,subtract,multiply,par_multiply,divide} (reference src/functions/scalar.rs:16-103)
// and the math_op binaries atan2/hypot/log (scalar.rs:148,274,291,499-523).
//
// Semantics kept (oracle/oracle.c states them with provenance):
//   * add/sub/mul compute every slot, also under nulls; integers wrap; IEEE ops round once (the
//     __*_rn intrinsics can never be contracted into FMAs);
//   * divide: a VALID slot with a zero divisor (ints and floats) raises the DivideByZero flag; null
//     slots divide by 1; iN::MIN / -1 wraps;
//   * math_op binaries: null slot -> payload 0;
//   * validity = a AND b at bit offset 0, zero padding bits; valid slots are counted per chunk.
//
// Roofline: HBM.  Algorithmic bytes/row = 3*sizeof(T) + (v_in + [v_in>0])/8  (24 B/row for f64, no nulls).
#include "common.cuh"

namespace bdf {

enum : int { OP_ADD = 0, OP_SUB, OP_MUL, OP_DIV, OP_ATAN2, OP_HYPOT, OP_LOG };

template <typename T> struct UnsignedOf { using type = T; };
template <> struct UnsignedOf<int8_t> { using type = uint8_t; };
template <> struct UnsignedOf<int16_t> { using type = uint16_t; };
template <> struct UnsignedOf<int32_t> { using type = uint32_t; };
template <> struct UnsignedOf<int64_t> { using type = uint64_t; };

template <size_t N> struct WideOf { using type = uint32_t; };
template <> struct WideOf<8> { using type = uint64_t; };

template <typename T> struct IsFloat { static constexpr bool value = false; };
template <> struct IsFloat<float> { static constexpr bool value = true; };
template <> struct IsFloat<double> { static constexpr bool value = true; };

// apply(a, b, valid, divzero): valid = output slot is valid; divzero is OR-ed when a valid divisor is 0.
template <typename T, int OP>
__device__ __forceinline__ T bin_apply(T a, T b, bool valid, bool& divzero) {
    if constexpr (IsFloat<T>::value) {
        if constexpr (sizeof(T) == 8) {
            if constexpr (OP == OP_ADD) return __dadd_rn(a, b);
            else if constexpr (OP == OP_SUB) return __dsub_rn(a, b);
            else if constexpr (OP == OP_MUL) return __dmul_rn(a, b);
            else if constexpr (OP == OP_DIV) {
                divzero |= valid && (b == 0.0);
                return __ddiv_rn(a, valid ? b : 1.0);
            } else if constexpr (OP == OP_ATAN2) return valid ? atan2(a, b) : 0.0;
            else if constexpr (OP == OP_HYPOT) return valid ? hypot(a, b) : 0.0;
            else return valid ? __ddiv_rn(log(a), log(b)) : 0.0;
        } else {
            if constexpr (OP == OP_ADD) return __fadd_rn(a, b);
            else if constexpr (OP == OP_SUB) return __fsub_rn(a, b);
            else if constexpr (OP == OP_MUL) return __fmul_rn(a, b);
            else if constexpr (OP == OP_DIV) {
                divzero |= valid && (b == 0.0f);
                return __fdiv_rn(a, valid ? b : 1.0f);
            } else if constexpr (OP == OP_ATAN2) return valid ? atan2f(a, b) : 0.0f;
            else if constexpr (OP == OP_HYPOT) return valid ? hypotf(a, b) : 0.0f;
            else return valid ? __fdiv_rn(logf(a), logf(b)) : 0.0f;
        }
    } else {
        using U = typename UnsignedOf<T>::type;
        using W = typename WideOf<sizeof(T)>::type;  // >= 32 bit unsigned: no signed-int promotion overflow
        if constexpr (OP == OP_ADD) return (T)(U)((W)(U)a + (W)(U)b);
        else if constexpr (OP == OP_SUB) return (T)(U)((W)(U)a - (W)(U)b);
        else if constexpr (OP == OP_MUL) return (T)(U)((W)(U)a * (W)(U)b);
        else {
            divzero |= valid && (b == (T)0);
            const T d = valid ? b : (T)1;
            if constexpr (((T)-1) < (T)0) {
                if (d == (T)-1) return (T)(U)((U)0 - (U)a);  // also covers MIN / -1 (wraps)
            }
            return (T)(a / (d == (T)0 ? (T)1 : d));  // d == 0 only on the error path; result is discarded
        }
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
k_binary(const BinDesc* __restrict__ descs, int n_chunks, unsigned long long* __restrict__ valid_counts,
         int* __restrict__ flags) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int TILE = kThreads * kUnroll * E;
    constexpr uint32_t FULLMASK = (E == 32) ? 0xffffffffu : ((1u << E) - 1u);
    __shared__ unsigned long long s_red[32];

    const int64_t tile = blockIdx.x;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const T* __restrict__ pa = (const T*)descs[c].a;
    const T* __restrict__ pb = (const T*)descs[c].b;
    T* __restrict__ po = (T*)descs[c].out;
    const uint32_t* __restrict__ va = descs[c].va;
    const uint32_t* __restrict__ vb = descs[c].vb;
    uint32_t* __restrict__ vo = descs[c].vout;
    const int64_t len = descs[c].len;
    const int64_t offa = descs[c].offa, offb = descs[c].offb;
    const int64_t base = (tile - descs[c].tile0) * TILE;

    unsigned int nvalid = 0;
    bool divzero = false;

    if (base + TILE <= len) {
        // ---- full tile: 2 x kUnroll 16-byte loads in flight per thread before first use ----
        Vec<T, E> a[kUnroll], b[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            a[j].load(pa + e0);
            b[j].load(pb + e0);
        }
        uint32_t m[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            m[j] = FULLMASK;
            if (va) m[j] &= load_bits<E>(va, offa + e0);
            if (vb) m[j] &= load_bits<E>(vb, offb + e0);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            Vec<T, E> r;
#pragma unroll
            for (int e = 0; e < E; e++) r.e[e] = bin_apply<T, OP>(a[j].e[e], b[j].e[e], (m[j] >> e) & 1u, divzero);
            r.store(po + e0);
            if (vo) {
                store_bits<E>(vo, e0, m[j], true);
                nvalid += __popc(m[j]);
            }
        }
    } else {
        // ---- tail tile of the chunk: element-wise guards ----
#pragma unroll 1
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            const uint32_t in_range = tail_mask<E>(e0, len);
            uint32_t m = in_range;
            if (in_range) {
                if (va) m &= load_bits<E>(va, offa + e0);
                if (vb) m &= load_bits<E>(vb, offb + e0);
            }
#pragma unroll
            for (int e = 0; e < E; e++) {
                if ((in_range >> e) & 1u) {
                    const T x = pa[e0 + e], y = pb[e0 + e];
                    po[e0 + e] = bin_apply<T, OP>(x, y, (m >> e) & 1u, divzero);
                }
            }
            if (vo) {
                store_bits<E>(vo, e0, m, in_range != 0);
                nvalid += __popc(m);
            }
        }
    }

    if (vo) {
        const unsigned long long total = block_sum_u64(nvalid, s_red);
        if (threadIdx.x == 0) atomicAdd(&valid_counts[c], total);
    }
    if constexpr (OP == OP_DIV) {
        if (divzero) atomicOr(flags, 1);
    }
}

int elems_per_tile(int dtype) { return kTileBytes / dtype_width(dtype); }

template <typename T, int OP>
static cudaError_t launch_one(const BinDesc* d, int n, int64_t tiles, unsigned long long* vc, int* flags, cudaStream_t s) {
    k_binary<T, OP><<<(unsigned)tiles, kThreads, 0, s>>>(d, n, vc, flags);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_wrapping(int op, const BinDesc* d, int n, int64_t tiles, unsigned long long* vc, int* flags,
                                   cudaStream_t s) {
    switch (op) {
        case OP_ADD: return launch_one<T, OP_ADD>(d, n, tiles, vc, flags, s);
        case OP_SUB: return launch_one<T, OP_SUB>(d, n, tiles, vc, flags, s);
        default: return launch_one<T, OP_MUL>(d, n, tiles, vc, flags, s);
    }
}

template <typename T>
static cudaError_t launch_float(int op, const BinDesc* d, int n, int64_t tiles, unsigned long long* vc, int* flags,
                                cudaStream_t s) {
    switch (op) {
        case OP_ADD: return launch_one<T, OP_ADD>(d, n, tiles, vc, flags, s);
        case OP_SUB: return launch_one<T, OP_SUB>(d, n, tiles, vc, flags, s);
        case OP_MUL: return launch_one<T, OP_MUL>(d, n, tiles, vc, flags, s);
        case OP_DIV: return launch_one<T, OP_DIV>(d, n, tiles, vc, flags, s);
        case OP_ATAN2: return launch_one<T, OP_ATAN2>(d, n, tiles, vc, flags, s);
        case OP_HYPOT: return launch_one<T, OP_HYPOT>(d, n, tiles, vc, flags, s);
        default: return launch_one<T, OP_LOG>(d, n, tiles, vc, flags, s);
    }
}

cudaError_t launch_binary(int op, int dtype, const BinDesc* d, int n, int64_t tiles, unsigned long long* vc, int* flags,
                          cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (dtype == T_F64) return launch_float<double>(op, d, n, tiles, vc, flags, s);
    if (dtype == T_F32) return launch_float<float>(op, d, n, tiles, vc, flags, s);
    if (op > OP_DIV) return cudaErrorInvalidValue;
    if (op == OP_DIV) {
        switch (dtype) {
            case T_I8: return launch_one<int8_t, OP_DIV>(d, n, tiles, vc, flags, s);
            case T_I16: return launch_one<int16_t, OP_DIV>(d, n, tiles, vc, flags, s);
            case T_I32: return launch_one<int32_t, OP_DIV>(d, n, tiles, vc, flags, s);
            case T_I64: return launch_one<int64_t, OP_DIV>(d, n, tiles, vc, flags, s);
            case T_U8: return launch_one<uint8_t, OP_DIV>(d, n, tiles, vc, flags, s);
            case T_U16: return launch_one<uint16_t, OP_DIV>(d, n, tiles, vc, flags, s);
            case T_U32: return launch_one<uint32_t, OP_DIV>(d, n, tiles, vc, flags, s);
            default: return launch_one<uint64_t, OP_DIV>(d, n, tiles, vc, flags, s);
        }
    }
    // add/sub/mul wrap: identical bits for signed and unsigned -> one instantiation per width
    switch (dtype_width(dtype)) {
        case 1: return launch_wrapping<uint8_t>(op, d, n, tiles, vc, flags, s);
        case 2: return launch_wrapping<uint16_t>(op, d, n, tiles, vc, flags, s);
        case 4: return launch_wrapping<uint32_t>(op, d, n, tiles, vc, flags, s);
        default: return launch_wrapping<uint64_t>(op, d, n, tiles, vc, flags, s);
    }
}

}  // namespace bdf
