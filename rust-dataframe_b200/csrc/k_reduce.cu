// k_reduce.cu -- K4: sum / min / max / count of a whole column (all chunks) in ONE pass and ONE launch.
//
// Replaces arrow::compute::{sum,max,(min)} as called per chunk by AggregateFunctions::{sum,max,min}
// and the metadata walk of AggregateFunctions::count (reference src/functions/aggregate.rs:12-31,70-93).
//   * integers: wrapping 64-bit sum (truncated to T by the caller: wrapping add is associative, so the
//     result is bit-identical to the reference's sequential fold), min, max, valid count;
//   * floats: sum accumulated in double, valid count (the reference has no float min/max:
//     T::Native: Ord).  Summation order is FIXED: each thread folds the elements of its tile in index
//     order, xor-shuffle tree inside the warp, warps in warp order, one partial per tile; k_finish folds
//     the partials with a grid and a per-thread assignment that depend only on the tile count and the SM
//     count.  No floating-point atomics anywhere => run-to-run deterministic.
//     The reference folds strictly left-to-right; the difference is covered by the stated tolerance.
//
// Roofline: HBM, sizeof(T) + [nullable]/8 bytes/row (8 B/row for f64/i64, 8.125 with a bitmap).
#include <type_traits>

#include "common.cuh"

namespace bdf {

template <typename T> struct RedInfo;
#define BDF_REDINFO(T, ISF, ISS) template <> struct RedInfo<T> { static constexpr bool is_float = ISF, is_signed = ISS; };
BDF_REDINFO(int8_t, false, true) BDF_REDINFO(int16_t, false, true) BDF_REDINFO(int32_t, false, true)
BDF_REDINFO(int64_t, false, true) BDF_REDINFO(uint8_t, false, false) BDF_REDINFO(uint16_t, false, false)
BDF_REDINFO(uint32_t, false, false) BDF_REDINFO(uint64_t, false, false) BDF_REDINFO(float, true, true)
BDF_REDINFO(double, true, true)
#undef BDF_REDINFO

// Running state of one thread / one CTA.  Integers: 64-bit wrapping sum of the sign-/zero-extended values
// and min/max as order-preserving unsigned keys (extended value ^ 2^63 for signed types), so that the
// type-agnostic k_finish can fold partials with plain unsigned compares.  Floats: sum in double.
template <typename T, bool IsFloat = RedInfo<T>::is_float> struct RedState;

template <typename T>
struct RedState<T, false> {
    // narrow types fold a tile into a 32-bit accumulator (<= 64 elements of <= 16 bits per thread and tile),
    // min/max are compared in T's own width and signedness; keys are formed once per CTA in to_dev
    using Acc = typename std::conditional<(sizeof(T) <= 2), typename std::conditional<RedInfo<T>::is_signed, int, unsigned int>::type,
                                          typename std::conditional<RedInfo<T>::is_signed, long long, unsigned long long>::type>::type;
    static constexpr unsigned long long kFlip = RedInfo<T>::is_signed ? (1ull << 63) : 0ull;
    unsigned long long sum; Acc acc; T mn, mx;
    __device__ __forceinline__ void init() {
        sum = 0; acc = 0;
        mn = RedInfo<T>::is_signed ? (T)((1ull << (8 * sizeof(T) - 1)) - 1ull) : (T)~0ull;
        mx = RedInfo<T>::is_signed ? (T)(-(long long)((1ull << (8 * sizeof(T) - 1)) - 1ull) - 1) : (T)0;
    }
    __device__ __forceinline__ void add(T x, bool valid) {
        acc += valid ? (Acc)x : (Acc)0;
        mn = (valid && x < mn) ? x : mn;
        mx = (valid && x > mx) ? x : mx;
    }
    __device__ __forceinline__ void add_all_valid(T x) { acc += (Acc)x; mn = x < mn ? x : mn; mx = x > mx ? x : mx; }
    __device__ __forceinline__ void end_tile() { sum += (unsigned long long)(long long)acc; acc = 0; }  // sign-/zero-extend, wrap
    __device__ __forceinline__ void merge(const RedState& o) { sum += o.sum; mn = o.mn < mn ? o.mn : mn; mx = o.mx > mx ? o.mx : mx; }
    __device__ __forceinline__ RedState shfl_xor(int o) const {
        RedState r;
        r.acc = 0;
        r.sum = __shfl_xor_sync(0xffffffffu, sum, o);
        if constexpr (sizeof(T) == 8) {
            r.mn = (T)__shfl_xor_sync(0xffffffffu, (long long)mn, o);
            r.mx = (T)__shfl_xor_sync(0xffffffffu, (long long)mx, o);
        } else {
            r.mn = (T)__shfl_xor_sync(0xffffffffu, (int)mn, o);
            r.mx = (T)__shfl_xor_sync(0xffffffffu, (int)mx, o);
        }
        return r;
    }
    __device__ __forceinline__ void to_dev(AggDev* d, unsigned long long cnt) const {
        d->sum_bits = sum;
        d->min_bits = (unsigned long long)(long long)mn ^ kFlip;  // order-preserving unsigned keys
        d->max_bits = (unsigned long long)(long long)mx ^ kFlip;
        d->count = cnt;
    }
};

template <typename T>
struct RedState<T, true> {
    double sum;
    __device__ __forceinline__ void init() { sum = 0.0; }
    __device__ __forceinline__ void add(T x, bool valid) { sum = __dadd_rn(sum, valid ? (double)x : 0.0); }
    __device__ __forceinline__ void add_all_valid(T x) { sum = __dadd_rn(sum, (double)x); }
    __device__ __forceinline__ void end_tile() {}
    __device__ __forceinline__ void merge(const RedState& o) { sum = __dadd_rn(sum, o.sum); }
    __device__ __forceinline__ RedState shfl_xor(int o) const {
        RedState r;
        r.sum = __shfl_xor_sync(0xffffffffu, sum, o);
        return r;
    }
    __device__ __forceinline__ void to_dev(AggDev* d, unsigned long long cnt) const {
        d->sum_bits = (unsigned long long)__double_as_longlong(sum); d->min_bits = ~0ull; d->max_bits = 0; d->count = cnt;
    }
};

// K consecutive tiles per CTA, not persistent (measured on the read-only f64 stream: 7.2 TB/s for K <= 2,
// 6.95 TB/s for grid-stride persistent variants, benchmarks/tune_stream.cu).  Floats use K = 2; integers
// K = 4 to amortise the heavier 3-value block reduction.  One partial per CTA in CTA order; launch_finish
// (k_finish, k_binary.cu) folds the partials with a fixed grid and assignment => deterministic.
template <typename T, int K>
__global__ void __launch_bounds__(kThreads)
k_reduce(const RedDesc* __restrict__ descs, int n_chunks, int64_t total_tiles, AggDev* __restrict__ cta_partials) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int TILE = kThreads * kUnroll * E;
    using S = RedState<T>;
    __shared__ S s_state[kWarpsPerCta];
    __shared__ unsigned int s_cnt[kWarpsPerCta];

    S st; st.init();
    unsigned int cnt = 0;
    int c = -1;
    int64_t c_tile0 = 0, c_tile_end = -1, len = 0, off = 0;
    const T* __restrict__ pi = nullptr;
    const uint32_t* __restrict__ vi = nullptr;

#pragma unroll 1
    for (int kk = 0; kk < K; kk++) {
        const int64_t tile = (int64_t)blockIdx.x * K + kk;
        if (tile >= total_tiles) break;
        if (tile >= c_tile_end) {  // first tile, or moved into the next chunk
            c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
            pi = (const T*)descs[c].in;
            vi = descs[c].vin;
            len = descs[c].len;
            off = descs[c].off;
            c_tile0 = descs[c].tile0;
            c_tile_end = c_tile0 + (len + TILE - 1) / TILE;
        }
        const int64_t base = (tile - c_tile0) * TILE;
        if (base + TILE <= len) {
            Vec<T, E> x[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; j++) x[j].load(pi + base + (int64_t)(j * kThreads + threadIdx.x) * E);
            if (vi) {
                MaskRaw<E, kUnroll> rv;  // validity words of all steps in one batch (see common.cuh)
                mask_issue<E, kUnroll>(rv, vi, off + base + (int64_t)threadIdx.x * E, (int64_t)kThreads * E);
#pragma unroll
                for (int j = 0; j < kUnroll; j++) {
                    const uint32_t m = mask_get<E, kUnroll>(rv, j);
#pragma unroll
                    for (int e = 0; e < E; e++) st.add(x[j].e[e], (m >> e) & 1u);
                    cnt += __popc(m);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kUnroll; j++)
#pragma unroll
                    for (int e = 0; e < E; e++) st.add_all_valid(x[j].e[e]);
                cnt += kUnroll * E;
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < kUnroll; j++) {
                const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
                const uint32_t in_range = tail_mask<E>(e0, len);
                uint32_t m = in_range;
                if (in_range && vi) m &= load_bits<E>(vi, off + e0);
#pragma unroll
                for (int e = 0; e < E; e++)
                    if ((in_range >> e) & 1u) st.add(pi[e0 + e], (m >> e) & 1u);
                cnt += __popc(m);
            }
        }
        st.end_tile();
    }
    // fixed xor-shuffle tree inside each warp, then thread 0 folds the warp results in warp order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { S other = st.shfl_xor(o); st.merge(other); }
    const unsigned int wcnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) { s_state[threadIdx.x >> 5] = st; s_cnt[threadIdx.x >> 5] = wcnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        S t = s_state[0];
        unsigned long long total = s_cnt[0];
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; w++) { t.merge(s_state[w]); total += s_cnt[w]; }
        t.to_dev(&cta_partials[blockIdx.x], total);
    }
}

constexpr int kReduceTilesInt = 4, kReduceTilesFloat = 2;
int64_t reduce_partials(int dtype, int64_t tiles) {
    const int k = dtype_is_float(dtype) ? kReduceTilesFloat : kReduceTilesInt;
    return (tiles + k - 1) / k;
}

template <typename T>
static cudaError_t launch_one(const RedDesc* d, int n, int64_t tiles, AggDev* partials, cudaStream_t s) {
    constexpr int K = RedInfo<T>::is_float ? kReduceTilesFloat : kReduceTilesInt;
    k_reduce<T, K><<<(unsigned)((tiles + K - 1) / K), kThreads, 0, s>>>(d, n, tiles, partials);
    return cudaGetLastError();
}

// Per-tile partials of chunks described by d (k_finish folds them; with tiles == 0 nothing is launched and
// k_finish produces the identity).
cudaError_t launch_reduce(int dtype, const RedDesc* d, int n, int64_t tiles, AggDev* partials, cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    switch (dtype) {
        case T_I8: return launch_one<int8_t>(d, n, tiles, partials, s);
        case T_I16: return launch_one<int16_t>(d, n, tiles, partials, s);
        case T_I32: return launch_one<int32_t>(d, n, tiles, partials, s);
        case T_I64: return launch_one<int64_t>(d, n, tiles, partials, s);
        case T_U8: return launch_one<uint8_t>(d, n, tiles, partials, s);
        case T_U16: return launch_one<uint16_t>(d, n, tiles, partials, s);
        case T_U32: return launch_one<uint32_t>(d, n, tiles, partials, s);
        case T_U64: return launch_one<uint64_t>(d, n, tiles, partials, s);
        case T_F32: return launch_one<float>(d, n, tiles, partials, s);
        case T_F64: return launch_one<double>(d, n, tiles, partials, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bdf
