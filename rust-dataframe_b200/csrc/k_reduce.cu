// k_reduce.cu -- K4: sum / min / max / count of a whole column (all chunks) in ONE pass and ONE launch.
//
// Replaces arrow::compute::{sum,max,(min)} as called per chunk by AggregateFunctions::{sum,max,min}
// and the metadata walk of AggregateFunctions::count (reference src/functions/aggregate.rs:12-31,70-93).
//   * integers: wrapping 64-bit sum (truncated to T by the caller: wrapping add is associative, so the
//     result is bit-identical to the reference's sequential fold), min, max, valid count;
//   * floats: sum accumulated in double, valid count (the reference has no float min/max:
//     T::Native: Ord).  Summation order is FIXED for a given device: tiles are dealt round-robin to a
//     grid whose size depends only on the SM count, each thread folds its elements in index order,
//     warp-shuffle tree, shared-memory tree over warps, then the last CTA (ticket) folds the per-CTA
//     partials in CTA order.  No floating-point atomics anywhere => run-to-run deterministic.
//     The reference folds strictly left-to-right; the difference is covered by the stated tolerance.
//
// Roofline: HBM, sizeof(T) + [nullable]/8 bytes/row (8 B/row for f64/i64, 8.125 with a bitmap).
#include "common.cuh"

namespace bdf {

template <typename T> struct RedInfo;
#define BDF_REDINFO(T, ISF, ISS) template <> struct RedInfo<T> { static constexpr bool is_float = ISF, is_signed = ISS; };
BDF_REDINFO(int8_t, false, true) BDF_REDINFO(int16_t, false, true) BDF_REDINFO(int32_t, false, true)
BDF_REDINFO(int64_t, false, true) BDF_REDINFO(uint8_t, false, false) BDF_REDINFO(uint16_t, false, false)
BDF_REDINFO(uint32_t, false, false) BDF_REDINFO(uint64_t, false, false) BDF_REDINFO(float, true, true)
BDF_REDINFO(double, true, true)
#undef BDF_REDINFO

template <typename T> __device__ __forceinline__ T type_max() {
    if constexpr (RedInfo<T>::is_signed) return (T)((1ull << (8 * sizeof(T) - 1)) - 1ull);
    else return (T)~0ull;
}
template <typename T> __device__ __forceinline__ T type_min() {
    if constexpr (RedInfo<T>::is_signed) return (T)(-(int64_t)((1ull << (8 * sizeof(T) - 1)) - 1ull) - 1);
    else return (T)0;
}

// Running state of one thread / one CTA.
template <typename T, bool IsFloat = RedInfo<T>::is_float> struct RedState;

template <typename T>
struct RedState<T, false> {
    unsigned long long sum; T mn, mx; unsigned long long cnt;
    __device__ __forceinline__ void init() { sum = 0; mn = type_max<T>(); mx = type_min<T>(); cnt = 0; }
    __device__ __forceinline__ void add(T x, bool valid) {
        if (valid) {
            sum += (unsigned long long)(long long)x;  // sign- or zero-extension, then wrapping add
            mn = x < mn ? x : mn;
            mx = x > mx ? x : mx;
        }
    }
    __device__ __forceinline__ void merge(const RedState& o) {
        sum += o.sum; mn = o.mn < mn ? o.mn : mn; mx = o.mx > mx ? o.mx : mx; cnt += o.cnt;
    }
    __device__ __forceinline__ RedState shfl_xor(int o) const {
        RedState r;
        r.sum = __shfl_xor_sync(0xffffffffu, sum, o);
        r.mn = (T)__shfl_xor_sync(0xffffffffu, (long long)mn, o);
        r.mx = (T)__shfl_xor_sync(0xffffffffu, (long long)mx, o);
        r.cnt = __shfl_xor_sync(0xffffffffu, cnt, o);
        return r;
    }
    __device__ __forceinline__ void to_dev(AggDev* d) const {
        d->sum_bits = sum; d->min_bits = (unsigned long long)(long long)mn; d->max_bits = (unsigned long long)(long long)mx;
        d->count = cnt;
    }
    __device__ __forceinline__ void from_dev(const AggDev* d) {
        sum = __ldcg(&d->sum_bits); mn = (T)(long long)__ldcg(&d->min_bits); mx = (T)(long long)__ldcg(&d->max_bits);
        cnt = __ldcg(&d->count);
    }
};

template <typename T>
struct RedState<T, true> {
    double sum; unsigned long long cnt;
    __device__ __forceinline__ void init() { sum = 0.0; cnt = 0; }
    __device__ __forceinline__ void add(T x, bool valid) { sum = __dadd_rn(sum, valid ? (double)x : 0.0); }
    __device__ __forceinline__ void merge(const RedState& o) { sum = __dadd_rn(sum, o.sum); cnt += o.cnt; }
    __device__ __forceinline__ RedState shfl_xor(int o) const {
        RedState r;
        r.sum = __shfl_xor_sync(0xffffffffu, sum, o);
        r.cnt = __shfl_xor_sync(0xffffffffu, cnt, o);
        return r;
    }
    __device__ __forceinline__ void to_dev(AggDev* d) const {
        d->sum_bits = (unsigned long long)__double_as_longlong(sum); d->min_bits = 0; d->max_bits = 0; d->count = cnt;
    }
    __device__ __forceinline__ void from_dev(const AggDev* d) {
        sum = __longlong_as_double((long long)__ldcg(&d->sum_bits)); cnt = __ldcg(&d->count);
    }
};

// Deterministic block reduction: xor-shuffle tree inside each warp, then warp 0 folds the warp results
// with the same tree.  Result valid in thread 0.
template <typename S>
__device__ __forceinline__ void block_reduce(S& st, S* smem /* >= kThreads/32 */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { S other = st.shfl_xor(o); st.merge(other); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // smem may still be in use by a previous call
    if (lane == 0) smem[warp] = st;
    __syncthreads();
    if (warp == 0) {
        S v; v.init();
        if (lane < kThreads / 32) v = smem[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { S other = v.shfl_xor(o); v.merge(other); }
        st = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_reduce(const RedDesc* __restrict__ descs, int n_chunks, int64_t total_tiles, AggDev* __restrict__ partials,
         unsigned int* __restrict__ ticket, AggDev* __restrict__ result) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int TILE = kThreads * kUnroll * E;
    constexpr uint32_t FULLMASK = (1u << E) - 1u;
    using S = RedState<T>;
    __shared__ S s_state[kThreads / 32];
    __shared__ bool s_last;

    S st; st.init();
    unsigned int cnt = 0;

    int c = 0;
    int64_t c_tile0 = 0, c_tile_end = -1;
    const T* __restrict__ pi = nullptr;
    const uint32_t* __restrict__ vi = nullptr;
    int64_t len = 0, off = 0;

    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (tile >= c_tile_end) {  // moved into another chunk (tiles are visited in increasing order)
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
            MaskRaw<E, kUnroll> rv;  // validity words of all steps in one batch (see common.cuh)
            if (vi) mask_issue<E, kUnroll>(rv, vi, off + base + (int64_t)threadIdx.x * E, (int64_t)kThreads * E);
            uint32_t m[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; j++) m[j] = vi ? mask_get<E, kUnroll>(rv, j) : FULLMASK;
#pragma unroll
            for (int j = 0; j < kUnroll; j++) {
#pragma unroll
                for (int e = 0; e < E; e++) st.add(x[j].e[e], (m[j] >> e) & 1u);
                cnt += __popc(m[j]);
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
    }
    st.cnt = cnt;

    block_reduce(st, s_state);
    if (threadIdx.x == 0) {
        st.to_dev(&partials[blockIdx.x]);
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;

    // Last CTA to finish: fold the per-CTA partials in CTA order (fixed order => deterministic).
    __threadfence();
    S acc; acc.init();
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) {
        S p; p.init(); p.from_dev(&partials[i]);
        acc.merge(p);
    }
    block_reduce(acc, s_state);
    if (threadIdx.x == 0) {
        acc.to_dev(result);  // result may live in device-mapped host memory
        __threadfence_system();
        *ticket = 0;  // ready for the next launch on this stream
    }
}

template <typename T>
static int occupancy_grid(int sm_count) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_reduce<T>, kThreads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    if (per_sm > 8) per_sm = 8;
    return sm_count * per_sm;
}

int reduce_grid(int sm_count) { return sm_count * 8; }  // upper bound used to size the partials buffer

template <typename T>
static cudaError_t launch_one(const RedDesc* d, int n, int64_t tiles, int grid_cap, AggDev* partials, unsigned int* ticket,
                              AggDev* result, cudaStream_t s) {
    static int grid_for_device = 0;  // one device per process (one process per GPU)
    if (grid_for_device == 0) grid_for_device = occupancy_grid<T>(grid_cap / 8);
    int64_t grid = grid_for_device;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    k_reduce<T><<<(unsigned)grid, kThreads, 0, s>>>(d, n, tiles, partials, ticket, result);
    return cudaGetLastError();
}

cudaError_t launch_reduce(int dtype, const RedDesc* d, int n, int64_t tiles, int grid_cap, AggDev* partials,
                          unsigned int* ticket, AggDev* result, cudaStream_t s) {
    switch (dtype) {
        case T_I8: return launch_one<int8_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_I16: return launch_one<int16_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_I32: return launch_one<int32_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_I64: return launch_one<int64_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_U8: return launch_one<uint8_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_U16: return launch_one<uint16_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_U32: return launch_one<uint32_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_U64: return launch_one<uint64_t>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_F32: return launch_one<float>(d, n, tiles, grid_cap, partials, ticket, result, s);
        case T_F64: return launch_one<double>(d, n, tiles, grid_cap, partials, ticket, result, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bdf
