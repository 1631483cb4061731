// k_cast.cu -- K3: numeric x numeric cast over all chunks of a column in ONE launch, with the
// "unrepresentable -> NULL" rule, validity pass-through and null count fused in.
//
// Replaces arrow::compute::cast(&ArrayRef, &DataType) as called per chunk by Function::Cast
// (reference src/evaluation.rs:296-315) -> arrow-rs numeric_cast -> num::cast::cast::<From,To>:
//   null -> null; Some(v') -> v'; None -> NULL (payload 0).  num-traits 0.2 ToPrimitive rules:
//   int->int    Some iff representable;      int->float / float->float  always Some (`as`, RN-even);
//   float->iN   Some iff MIN-1 < v < MAX+1 when the float is wider than the int (both bounds exact),
//               else iff (iN::MIN as f) <= v < 2^(bits-1); truncation toward zero;
//   float->uN   Some iff -1 < v < 2^bits;    NaN / +-inf -> None.
// Bit-exact vs the oracle (including which slots turn NULL).
//
// Roofline: HBM, sizeof(From) + sizeof(To) + 2*[nullable]/8 bytes/row (Int32->Float64 = 12 B/row).
// A thread converts E = 16 / max(sizeof(From), sizeof(To)) elements per step so that the wider side
// moves 16 bytes per lane and both sides stay fully coalesced.
#include "common.cuh"

namespace bdf {

template <typename T> struct NumInfo;
#define BDF_NUMINFO(T, ISF, ISS) template <> struct NumInfo<T> { static constexpr bool is_float = ISF, is_signed = ISS; };
BDF_NUMINFO(int8_t, false, true) BDF_NUMINFO(int16_t, false, true) BDF_NUMINFO(int32_t, false, true)
BDF_NUMINFO(int64_t, false, true) BDF_NUMINFO(uint8_t, false, false) BDF_NUMINFO(uint16_t, false, false)
BDF_NUMINFO(uint32_t, false, false) BDF_NUMINFO(uint64_t, false, false) BDF_NUMINFO(float, true, true)
BDF_NUMINFO(double, true, true)
#undef BDF_NUMINFO

template <typename T> __device__ __forceinline__ constexpr int64_t int_max_of() {
    return NumInfo<T>::is_signed ? (int64_t)((1ull << (8 * sizeof(T) - 1)) - 1ull) : 0;
}
template <typename T> __device__ __forceinline__ constexpr uint64_t uint_max_of() {
    return sizeof(T) == 8 ? ~0ull : ((1ull << (8 * (sizeof(T) & 7))) - 1ull);
}

// num::cast::cast::<F,T>(v): returns false for None.
template <typename F, typename T> struct SameType { static constexpr bool value = false; };
template <typename T> struct SameType<T, T> { static constexpr bool value = true; };

template <typename F, typename T>
__device__ __forceinline__ bool cast_one(F v, T& out) {
    if constexpr (NumInfo<T>::is_float) {
        out = (T)v;  // int->float: round-to-nearest-even; f64->f32: RN, overflow -> +-inf; f32->f64 exact
        return true;
    } else if constexpr (NumInfo<F>::is_float) {
        constexpr int bits = 8 * (int)sizeof(T);
        bool ok;
        if constexpr (NumInfo<T>::is_signed) {
            const F max_p1 = (F)(1ull << (bits - 1));  // 2^(bits-1): exact in f32 and f64
            if constexpr (sizeof(F) > sizeof(T)) ok = (v > -max_p1 - (F)1) && (v < max_p1);
            else ok = (v >= -max_p1) && (v < max_p1);
        } else {
            const F max_p1 = (bits == 64) ? (F)18446744073709551616.0 : (F)(1ull << (bits & 63));
            ok = (v > (F)-1) && (v < max_p1);
        }
        out = ok ? (T)v : (T)0;  // in-window conversion truncates toward zero
        return ok;
    } else if constexpr (NumInfo<F>::is_signed) {
        const int64_t x = (int64_t)v;
        bool ok;
        if constexpr (NumInfo<T>::is_signed) ok = (x >= -int_max_of<T>() - 1) && (x <= int_max_of<T>());
        else ok = (x >= 0) && ((uint64_t)x <= uint_max_of<T>());
        out = ok ? (T)x : (T)0;
        return ok;
    } else {
        const uint64_t x = (uint64_t)v;
        bool ok;
        if constexpr (NumInfo<T>::is_signed) ok = x <= (uint64_t)int_max_of<T>();
        else ok = x <= uint_max_of<T>();
        out = ok ? (T)x : (T)0;
        return ok;
    }
}

template <typename F, typename T>
__global__ void __launch_bounds__(kThreads)
k_cast(const UnDesc* __restrict__ descs, int n_chunks, uint32_t* __restrict__ warp_counts) {
    constexpr int W = sizeof(F) > sizeof(T) ? (int)sizeof(F) : (int)sizeof(T);
    constexpr int E = 16 / W;
    constexpr int TILE = kThreads * kUnroll * E;
    constexpr uint32_t FULLMASK = (1u << E) - 1u;

    const int64_t tile = blockIdx.x;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const F* __restrict__ pi = (const F*)descs[c].in;
    T* __restrict__ po = (T*)descs[c].out;
    const uint32_t* __restrict__ vi = descs[c].vin;
    uint32_t* __restrict__ vo = descs[c].vout;
    const int64_t len = descs[c].len;
    const int64_t off = descs[c].off;
    const int64_t base = (tile - descs[c].tile0) * TILE;

    unsigned int nvalid = 0;
    if (base + TILE <= len) {
        Vec<F, E> x[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) x[j].load(pi + base + (int64_t)(j * kThreads + threadIdx.x) * E);
        MaskRaw<E, kUnroll> rv;  // validity words of all steps in one batch (see common.cuh)
        if (vi) mask_issue<E, kUnroll>(rv, vi, off + base + (int64_t)threadIdx.x * E, (int64_t)kThreads * E);
        uint32_t m[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) m[j] = vi ? mask_get<E, kUnroll>(rv, j) : FULLMASK;
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            Vec<T, E> r;
            uint32_t ok = 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
                T y;
                const bool good = cast_one<F, T>(x[j].e[e], y) && ((m[j] >> e) & 1u);
                r.e[e] = (good || SameType<F, T>::value) ? y : (T)0;  // same type = clone: payloads survive
                ok |= (good ? 1u : 0u) << e;
            }
            r.store(po + e0);
            if (vo) {
                store_bits<E>(vo, e0, ok, true);
                nvalid += __popc(ok);
            }
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            const uint32_t in_range = tail_mask<E>(e0, len);
            uint32_t m = in_range;
            if (in_range && vi) m &= load_bits<E>(vi, off + e0);
            uint32_t ok = 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
                if ((in_range >> e) & 1u) {
                    T y;
                    const bool good = cast_one<F, T>(pi[e0 + e], y) && ((m >> e) & 1u);
                    po[e0 + e] = (good || SameType<F, T>::value) ? y : (T)0;
                    ok |= (good ? 1u : 0u) << e;
                }
            }
            if (vo) {
                store_bits<E>(vo, e0, ok, in_range != 0);
                nvalid += __popc(ok);
            }
        }
    }
    if (vo) {
        const unsigned int wvalid = __reduce_add_sync(0xffffffffu, nvalid);  // per-warp count, plain store (no atomics)
        if ((threadIdx.x & 31) == 0) warp_counts[(int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)] = wvalid;
    }
}

int elems_per_tile_cast(int from, int to) {
    const int w = dtype_width(from) > dtype_width(to) ? dtype_width(from) : dtype_width(to);
    return kTileBytes / w;
}

template <typename F, typename T>
static cudaError_t launch_one(const UnDesc* d, int n, int64_t tiles, uint32_t* vc, cudaStream_t s) {
    k_cast<F, T><<<(unsigned)tiles, kThreads, 0, s>>>(d, n, vc);
    return cudaGetLastError();
}

template <typename F>
static cudaError_t launch_from(int to, const UnDesc* d, int n, int64_t tiles, uint32_t* vc, cudaStream_t s) {
    switch (to) {
        case T_I8: return launch_one<F, int8_t>(d, n, tiles, vc, s);
        case T_I16: return launch_one<F, int16_t>(d, n, tiles, vc, s);
        case T_I32: return launch_one<F, int32_t>(d, n, tiles, vc, s);
        case T_I64: return launch_one<F, int64_t>(d, n, tiles, vc, s);
        case T_U8: return launch_one<F, uint8_t>(d, n, tiles, vc, s);
        case T_U16: return launch_one<F, uint16_t>(d, n, tiles, vc, s);
        case T_U32: return launch_one<F, uint32_t>(d, n, tiles, vc, s);
        case T_U64: return launch_one<F, uint64_t>(d, n, tiles, vc, s);
        case T_F32: return launch_one<F, float>(d, n, tiles, vc, s);
        case T_F64: return launch_one<F, double>(d, n, tiles, vc, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_cast(int from, int to, const UnDesc* d, int n, int64_t tiles, uint32_t* vc, cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    switch (from) {
        case T_I8: return launch_from<int8_t>(to, d, n, tiles, vc, s);
        case T_I16: return launch_from<int16_t>(to, d, n, tiles, vc, s);
        case T_I32: return launch_from<int32_t>(to, d, n, tiles, vc, s);
        case T_I64: return launch_from<int64_t>(to, d, n, tiles, vc, s);
        case T_U8: return launch_from<uint8_t>(to, d, n, tiles, vc, s);
        case T_U16: return launch_from<uint16_t>(to, d, n, tiles, vc, s);
        case T_U32: return launch_from<uint32_t>(to, d, n, tiles, vc, s);
        case T_U64: return launch_from<uint64_t>(to, d, n, tiles, vc, s);
        case T_F32: return launch_from<float>(to, d, n, tiles, vc, s);
        case T_F64: return launch_from<double>(to, d, n, tiles, vc, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bdf
