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
// Roofline: HBM, sizeof(T) + [nullable]/8 bytes/row (8 B/row for f64/i64, 8.125 with a bitmap).  The integer
// instantiations are limited by the ALU pipe (64 lanes/clk/SM: ISETP/SEL/LOP3/VIMNMX), not by memory, unless the
// per-element instruction count is kept down -- round 1 measured 30 instructions per Int64 element (ALU pipe 79 %
// busy, DRAM 69 %).  What each width does about it:
//   8-byte   4 ISETP + 4 SEL per element for min/max and a 64-bit add, all under the validity predicate; the validity
//            words of a tile come off one base pointer (common.cuh mask_issue), which removed a third of the
//            instructions of a nullable tile.
//   4-byte   predicated VIMNMX for min/max, predicated IMAD.WIDE (FMA pipe) for the 64-bit sum.
//   1/2-byte SIMD in a register: a 32-bit word holds 4 or 2 elements; validity bits are expanded to byte / half-word
//            masks with two multiplies, nulls are replaced by the identity with one LOP3 per word, min/max run on
//            packed 16-bit lanes (VIMNMX.U16x2 / .S16x2 -- bytes are split into even and odd lanes first), the sum
//            is IDP.4A / IDP.2A against 0x01010101.
#include <algorithm>
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
    static constexpr bool kSigned = RedInfo<T>::is_signed;
    static constexpr unsigned long long kFlip = kSigned ? (1ull << 63) : 0ull;
    unsigned long long sum; T mn, mx;
    __device__ __forceinline__ static T id_min() { return kSigned ? (T)((1ull << (8 * sizeof(T) - 1)) - 1ull) : (T)~0ull; }
    __device__ __forceinline__ static T id_max() { return kSigned ? (T)(-(long long)((1ull << (8 * sizeof(T) - 1)) - 1ull) - 1) : (T)0; }
    __device__ __forceinline__ void init() { sum = 0; mn = id_min(); mx = id_max(); }
    __device__ __forceinline__ void add(T x, bool valid) {
        if (valid) {
            sum += (unsigned long long)(long long)x;   // sign-/zero-extended, wrapping
            mn = x < mn ? x : mn;
            mx = x > mx ? x : mx;
        }
    }
    __device__ __forceinline__ void merge(const RedState& o) { sum += o.sum; mn = o.mn < mn ? o.mn : mn; mx = o.mx > mx ? o.mx : mx; }
    __device__ __forceinline__ RedState shfl_xor(int o) const {
        RedState r;
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
    __device__ __forceinline__ static unsigned long long key(T v) { return (unsigned long long)(long long)v ^ kFlip; }
    __device__ __forceinline__ static T from_key(unsigned long long k) { return (T)(long long)(k ^ kFlip); }
    __device__ __forceinline__ void to_dev(AggDev* d, unsigned long long cnt) const {
        d->sum_bits = sum; d->min_bits = key(mn); d->max_bits = key(mx); d->count = cnt;
    }
};

template <typename T>
struct RedState<T, true> {
    double sum;
    __device__ __forceinline__ void init() { sum = 0.0; }
    __device__ __forceinline__ void add(T x, bool valid) { sum = __dadd_rn(sum, valid ? (double)x : 0.0); }
    __device__ __forceinline__ void add_all_valid(T x) { sum = __dadd_rn(sum, (double)x); }
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

// ---- one full tile, per element width -----------------------------------------------------------------------------
// x[kUnroll]: the thread's vectors of the tile; HAS_V: rv holds the raw validity words (mask_get(rv, j) = the E bits
// of vector j).  Returns the number of valid elements.

// 8-byte integers: everything under the validity predicate (ISETP pairs + SEL for min/max, a 64-bit add for the sum).
// (A filter-then-update variant with a launch-wide min/max hint was tried in round 2: the filter needs the same four
// ISETP per element as the update and the first wave of CTAs pays for both -- 0.21 ms instead of 0.147 ms.)
template <typename T, bool HAS_V>
__device__ __forceinline__ unsigned int tile_int64(RedState<T>& st, const Vec<T, 2> (&x)[kUnroll], const MaskRaw<2, kUnroll>& rv) {
    unsigned int cnt = 0;
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint32_t m = HAS_V ? mask_get<2, kUnroll>(rv, j) : 3u;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const T v = x[j].e[e];
            const uint32_t ok = HAS_V ? ((m >> e) & 1u) : 1u;
            // sum += v * ok as two multiply-adds (IMAD.WIDE + IMAD: the FMA pipe, which this kernel leaves idle; the ALU
            // pipe is what limits it): low word into the 64-bit sum with carry, high word into its upper half
            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(st.sum) : "r"((uint32_t)(unsigned long long)v), "r"(ok));
            asm("{.reg .b32 l, h;\n mov.b64 {l, h}, %0;\n mad.lo.u32 h, %1, %2, h;\n mov.b64 %0, {l, h};}"
                : "+l"(st.sum) : "r"((uint32_t)((unsigned long long)v >> 32)), "r"(ok));
            if (ok) {
                st.mn = v < st.mn ? v : st.mn;
                st.mx = v > st.mx ? v : st.mx;
            }
        }
        cnt += HAS_V ? __popc(m) : 2;
    }
    return cnt;
}

// 4-byte integers: predicated min/max, 64-bit sum by a widening multiply-add (x * valid + sum on the FMA pipe).
template <typename T, bool HAS_V>
__device__ __forceinline__ unsigned int tile_int32(RedState<T>& st, const Vec<T, 4> (&x)[kUnroll], const MaskRaw<4, kUnroll>& rv) {
    unsigned int cnt = 0;
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint32_t m = HAS_V ? mask_get<4, kUnroll>(rv, j) : 15u;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const T v = x[j].e[e];
            const bool ok = !HAS_V || ((m >> e) & 1u);
            if (ok) {
                if constexpr (RedInfo<T>::is_signed) asm("mad.wide.s32 %0, %1, 1, %0;" : "+l"(st.sum) : "r"((int)v));
                else asm("mad.wide.u32 %0, %1, 1, %0;" : "+l"(st.sum) : "r"((unsigned)v));
                st.mn = v < st.mn ? v : st.mn;
                st.mx = v > st.mx ? v : st.mx;
            }
        }
        cnt += HAS_V ? __popc(m) : 4;
    }
    return cnt;
}

// Packed accumulators of the 1- and 2-byte instantiations (folded into the scalar state once per CTA-thread).
template <typename T>
struct Packed {
    static constexpr bool kSigned = RedInfo<T>::is_signed;
    // 2-byte: mn/mx hold two lanes of T.  1-byte: lanes hold order-preserving UNSIGNED byte keys (value ^ 0x80 for
    // signed) zero-extended to 16 bits -- [0] even bytes, [1] odd bytes.
    uint32_t mn[2], mx[2];
    int acc;   // 32-bit sum of one tile (<= 64 elements of <= 16 bits), flushed into the 64-bit sum per tile
    __device__ __forceinline__ void init() {
        if constexpr (sizeof(T) == 2) { mn[0] = mn[1] = kSigned ? 0x7fff7fffu : 0xffffffffu; mx[0] = mx[1] = kSigned ? 0x80008000u : 0u; }
        else { mn[0] = mn[1] = 0x00ff00ffu; mx[0] = mx[1] = 0u; }
        acc = 0;
    }
    __device__ __forceinline__ void flush(RedState<T>& st) {
        st.sum += kSigned ? (unsigned long long)(long long)acc : (unsigned long long)(unsigned int)acc;
        acc = 0;
    }
    __device__ __forceinline__ void fold_into(RedState<T>& st) const {
        if constexpr (sizeof(T) == 2) {
            const uint32_t a = kSigned ? __vmins2(mn[0], mn[1]) : __vminu2(mn[0], mn[1]);
            const uint32_t b = kSigned ? __vmaxs2(mx[0], mx[1]) : __vmaxu2(mx[0], mx[1]);
            const T a0 = (T)(a & 0xffffu), a1 = (T)(a >> 16), b0 = (T)(b & 0xffffu), b1 = (T)(b >> 16);
            const T lo = a0 < a1 ? a0 : a1, hi = b0 > b1 ? b0 : b1;
            st.mn = lo < st.mn ? lo : st.mn;
            st.mx = hi > st.mx ? hi : st.mx;
        } else {
            const uint32_t a = __vminu2(mn[0], mn[1]), b = __vmaxu2(mx[0], mx[1]);
            const uint32_t klo = min(a & 0xffffu, a >> 16), khi = max(b & 0xffffu, b >> 16);
            // an untouched accumulator holds the identity keys (0xff / 0x00): they map back to T's identities
            const T lo = (T)(klo ^ (kSigned ? 0x80u : 0u)), hi = (T)(khi ^ (kSigned ? 0x80u : 0u));
            st.mn = lo < st.mn ? lo : st.mn;
            st.mx = hi > st.mx ? hi : st.mx;
        }
    }
};

// 2-byte integers: two elements per 32-bit word.
template <typename T, bool HAS_V>
__device__ __forceinline__ unsigned int tile_int16(Packed<T>& pk, const Vec<T, 8> (&x)[kUnroll], const MaskRaw<8, kUnroll>& rv) {
    constexpr bool S = RedInfo<T>::is_signed;
    constexpr uint32_t ID_MIN = S ? 0x7fff7fffu : 0xffffffffu, ID_MAX = S ? 0x80008000u : 0u;
    unsigned int cnt = 0;
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint32_t m = HAS_V ? mask_get<8, kUnroll>(rv, j) : 0xffu;
        const uint32_t w[4] = {x[j].q.x, x[j].q.y, x[j].q.z, x[j].q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t wmin = w[k], wmax = w[k], wsum = w[k];
            if (HAS_V) {
                const uint32_t m16 = (((m >> (2 * k)) * 0x8001u) & 0x00010001u) * 0xffffu;   // 2 bits -> two half-word masks
                wmin = (w[k] & m16) | (~m16 & ID_MIN);
                wmax = (w[k] & m16) | (~m16 & ID_MAX);
                wsum = w[k] & m16;
            }
            if constexpr (S) {
                pk.mn[k & 1] = __vmins2(pk.mn[k & 1], wmin); pk.mx[k & 1] = __vmaxs2(pk.mx[k & 1], wmax);
                pk.acc = __dp2a_lo((int)wsum, 0x0101, pk.acc);
            } else {
                pk.mn[k & 1] = __vminu2(pk.mn[k & 1], wmin); pk.mx[k & 1] = __vmaxu2(pk.mx[k & 1], wmax);
                pk.acc = (int)__dp2a_lo(wsum, 0x0101u, (unsigned)pk.acc);
            }
        }
        cnt += HAS_V ? __popc(m) : 8;
    }
    return cnt;
}

// 1-byte integers: four elements per 32-bit word; min/max on unsigned byte keys split into even / odd 16-bit lanes.
template <typename T, bool HAS_V>
__device__ __forceinline__ unsigned int tile_int8(Packed<T>& pk, const Vec<T, 16> (&x)[kUnroll], const MaskRaw<16, kUnroll>& rv) {
    constexpr bool S = RedInfo<T>::is_signed;
    constexpr uint32_t FLIP = S ? 0x80808080u : 0u;
    unsigned int cnt = 0;
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint32_t m = HAS_V ? mask_get<16, kUnroll>(rv, j) : 0xffffu;
        const uint32_t w[4] = {x[j].q.x, x[j].q.y, x[j].q.z, x[j].q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t kmin = w[k] ^ FLIP, kmax = w[k] ^ FLIP, wsum = w[k];
            if (HAS_V) {
                const uint32_t m8 = ((((m >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;   // 4 bits -> four byte masks
                kmin = (w[k] ^ FLIP) | ~m8;    // nulls -> 0xff, the identity of min over unsigned keys
                kmax = (w[k] ^ FLIP) & m8;     // nulls -> 0x00
                wsum = w[k] & m8;
            }
            pk.mn[0] = __vminu2(pk.mn[0], kmin & 0x00ff00ffu);
            pk.mn[1] = __vminu2(pk.mn[1], __byte_perm(kmin, 0u, 0x4341));   // bytes 1 and 3, zero-extended
            pk.mx[0] = __vmaxu2(pk.mx[0], kmax & 0x00ff00ffu);
            pk.mx[1] = __vmaxu2(pk.mx[1], __byte_perm(kmax, 0u, 0x4341));
            if constexpr (S) pk.acc = __dp4a((int)wsum, 0x01010101, pk.acc);
            else pk.acc = (int)__dp4a(wsum, 0x01010101u, (unsigned)pk.acc);
        }
        cnt += HAS_V ? __popc(m) : 16;
    }
    return cnt;
}

// K consecutive tiles per CTA (K = 2 for floats; 8 for integers, whose three-value block reduction costs ~200
// instructions per thread and has to be amortised), not persistent (measured on the read-only f64 stream:
// 7.2 TB/s for K <= 2, 6.95 TB/s for grid-stride persistent variants, benchmarks/tune_stream.cu).  One partial per CTA in
// CTA order; launch_finish (k_finish, k_binary.cu) folds the partials with a fixed grid and assignment => deterministic.
// Folding inside the launch ("last CTA out" tickets, one or two levels) was tried in round 2 and is slower: the fold code
// costs 22-24 registers (occupancy 66 % -> 36-48 %) and every CTA waits for a fence + atomic round trip before it retires
// (f64 sum 0.125 ms -> 0.134 ms with a 4096-CTA grid, 0.181 ms with two-level tickets; profiles/r2_reduce_notes.md).
// Resident CTAs per SM the register allocator is asked to leave room for: the reductions are latency machines (bytes in
// flight = resident threads x 64 B), round 1 ran the float instantiation at 7 CTAs/SM (34 registers) and the integer ones at 5.
template <typename T> struct RedOcc { static constexpr int value = RedInfo<T>::is_float ? 6 : 5; };

template <typename T, int K>
__global__ void __launch_bounds__(kThreads, RedOcc<T>::value)
k_reduce(const RedDesc* __restrict__ descs, int n_chunks, int64_t total_tiles, AggDev* __restrict__ cta_partials) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int TILE = kThreads * kUnroll * E;
    constexpr bool IS_INT = !RedInfo<T>::is_float;
    using S = RedState<T>;
    __shared__ S s_state[kWarpsPerCta];
    __shared__ unsigned int s_cnt[kWarpsPerCta];

    S st; st.init();
    [[maybe_unused]] Packed<typename std::conditional<(IS_INT && sizeof(T) <= 2), T, int8_t>::type> pk;
    if constexpr (IS_INT && sizeof(T) <= 2) pk.init();
    unsigned int cnt = 0;

    // This CTA's tiles [tile, tile_end) are walked chunk by chunk; inside a chunk the full tiles are a pointer-bumping loop
    // (no 64-bit index arithmetic, no chunk lookup per tile -- that overhead was a fifth of the instructions of the
    // integer instantiations), then at most one partial tile at the end of the chunk.
    int64_t tile = (int64_t)blockIdx.x * K;
    const int64_t tile_end = min(tile + K, total_tiles);
#pragma unroll 1
    while (tile < tile_end) {
        const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
        const T* __restrict__ pi = (const T*)descs[c].in;
        const uint32_t* __restrict__ vi = descs[c].vin;
        const int64_t len = descs[c].len, off = descs[c].off, c_tile0 = descs[c].tile0;
        const int64_t chunk_end = c_tile0 + (len + TILE - 1) / TILE;
        if (tile >= chunk_end) break;   // padding between the columns of a batched launch (columns start on CTA boundaries)
        const int64_t run_end = min(tile_end, chunk_end);   // this CTA's tiles inside chunk c
        const int n_full = (int)max((int64_t)0, min(run_end, c_tile0 + len / TILE) - tile);
        const int64_t base0 = (tile - c_tile0) * TILE;
        const T* __restrict__ p = pi + base0 + (int64_t)threadIdx.x * E;
        const int64_t bit0 = off + base0 + (int64_t)threadIdx.x * E;
        const uint32_t* __restrict__ vp = vi ? vi + (bit0 >> 5) : nullptr;
        const int sh = (int)(bit0 & 31);   // TILE is a multiple of 32 bits: the shift is the same for every tile of the run
#pragma unroll 1
        for (int t = 0; t < n_full; t++) {
            Vec<T, E> x[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; j++) x[j].load(p + (int64_t)j * kThreads * E);
            if (vi) {
                MaskRaw<E, kUnroll> rv;  // validity words of all steps in one batch (see common.cuh)
                mask_issue_at<E, kUnroll>(rv, vp, sh, kThreads * E / 32);
                if constexpr (!IS_INT) {
#pragma unroll
                    for (int j = 0; j < kUnroll; j++) {
                        const uint32_t m = mask_get<E, kUnroll>(rv, j);
#pragma unroll
                        for (int e = 0; e < E; e++) st.add(x[j].e[e], (m >> e) & 1u);
                        cnt += __popc(m);
                    }
                } else if constexpr (sizeof(T) == 8) cnt += tile_int64<T, true>(st, x, rv);
                else if constexpr (sizeof(T) == 4) cnt += tile_int32<T, true>(st, x, rv);
                else if constexpr (sizeof(T) == 2) cnt += tile_int16<T, true>(pk, x, rv);
                else cnt += tile_int8<T, true>(pk, x, rv);
                vp += TILE / 32;
            } else {
                MaskRaw<E, kUnroll> none{};   // not read by the HAS_V = false instantiations
                if constexpr (!IS_INT) {
#pragma unroll
                    for (int j = 0; j < kUnroll; j++)
#pragma unroll
                        for (int e = 0; e < E; e++) st.add_all_valid(x[j].e[e]);
                    cnt += kUnroll * E;
                } else if constexpr (sizeof(T) == 8) cnt += tile_int64<T, false>(st, x, none);
                else if constexpr (sizeof(T) == 4) cnt += tile_int32<T, false>(st, x, none);
                else if constexpr (sizeof(T) == 2) cnt += tile_int16<T, false>(pk, x, none);
                else cnt += tile_int8<T, false>(pk, x, none);
            }
            if constexpr (IS_INT && sizeof(T) <= 2) pk.flush(st);
            p += TILE;
        }
        tile += n_full;
        if (tile < run_end) {   // the partial last tile of the chunk
            const int64_t base = (tile - c_tile0) * TILE;
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
            tile++;
        }
    }
    if constexpr (IS_INT && sizeof(T) <= 2) pk.fold_into(st);
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

constexpr int kReduceTilesInt = 8, kReduceTilesFloat = 2;
static int tiles_per_cta(int dtype) { return dtype_is_float(dtype) ? kReduceTilesFloat : kReduceTilesInt; }
int64_t reduce_partials(int dtype, int64_t tiles) {   // partials launch_reduce writes
    const int k = tiles_per_cta(dtype);
    return (tiles + k - 1) / k;
}
int reduce_tiles_per_cta(int dtype) { return tiles_per_cta(dtype); }

template <typename T>
static cudaError_t launch_one(const RedDesc* d, int n, int64_t tiles, AggDev* partials, cudaStream_t s) {
    constexpr int K = RedInfo<T>::is_float ? kReduceTilesFloat : kReduceTilesInt;
    k_reduce<T, K><<<(unsigned)((tiles + K - 1) / K), kThreads, 0, s>>>(d, n, tiles, partials);
    return cudaGetLastError();
}

// Per-CTA partials of the chunks described by d (k_finish folds them; with tiles == 0 nothing is launched and
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
