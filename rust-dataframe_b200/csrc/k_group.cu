// k_group.cu -- group-by aggregate on SORTED rows (SURVEY 8(f) tail: `GroupAggregate`, a `panic!` in the reference,
// src/evaluation.rs:73; the shape of the operator is src/expression.rs:114-221 `try_aggregate`).
//
// The host (runtime.cu group_aggregate_dev) sorts the key column with the stable radix lexsort of k_sort.cu and gathers
// key and value columns with k_take, so every group is one contiguous run of rows and the rows of a group keep their
// original order.  Two kernels finish the job:
//   k_group_heads   bit i = 1 iff sorted row i starts a group (key differs from row i-1; nulls form ONE group, after all
//                   values -- the order of DataFrame::sort; NaN == NaN, -0.0 == 0.0 like the sort keys).  The bitmap is a
//                   boolean column: `filter(sorted key, heads)` = the distinct keys, `filter(iota, heads)` = the first row of
//                   every group -- both through the existing stream compaction of k_filter.cu.
//   k_group_reduce  one warp per group: lane l folds rows start + l, start + l + 32, ... in order, then a fixed xor-shuffle
//                   tree: deterministic (no atomics on values), integers wrap like AggregateFunctions::sum, floats are
//                   accumulated in double like K4.  sum / count always exist; min / max (integers: T::Native: Ord) are
//                   NULL for a group without a valid value.  A group of tens of millions of rows is one warp's work --
//                   correct, slow; splitting giant groups over warps is the obvious next step and is not done.
#include "common.cuh"

namespace bdf {

template <typename T>
__device__ __forceinline__ bool group_same_key(T a, T b) {
    if constexpr (IsFloat<T>::value) return a == b || (a != a && b != b);
    else return a == b;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_group_heads(const T* __restrict__ key, const uint32_t* __restrict__ kvalid, long long n, uint32_t* __restrict__ head_words) {
    const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    bool h = false;
    if (i < n) {
        if (i == 0) h = true;
        else {
            const bool vi = kvalid ? ((kvalid[i >> 5] >> (i & 31)) & 1u) : true;
            const bool vp = kvalid ? ((kvalid[(i - 1) >> 5] >> ((i - 1) & 31)) & 1u) : true;
            if (vi != vp) h = true;
            else if (vi) h = !group_same_key(key[i - 1], key[i]);
        }
    }
    const unsigned w = __ballot_sync(0xffffffffu, h);
    if ((threadIdx.x & 31) == 0 && i < n) head_words[i >> 5] = w;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_group_reduce(const T* __restrict__ val, const uint32_t* __restrict__ vvalid, const uint32_t* __restrict__ starts, long long n_groups,
               long long n_rows, T* __restrict__ out_sum, long long* __restrict__ out_count, T* __restrict__ out_min, T* __restrict__ out_max,
               uint32_t* __restrict__ mm_valid) {
    const long long g = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const int lane = threadIdx.x & 31;
    const long long b = starts[g], e = g + 1 < n_groups ? (long long)starts[g + 1] : n_rows;
    FusedAgg<T> a;
    a.init();
    unsigned long long cnt = 0;
    constexpr unsigned long long flip = IsFloat<T>::value ? 0ull : (((T)-1 < (T)0) ? (1ull << (8 * sizeof(T) - 1)) : 0ull);
    for (long long i = b + lane; i < e; i += 32) {
        const bool ok = vvalid ? ((vvalid[i >> 5] >> (i & 31)) & 1u) : true;
        a.add(val[i], ok, flip);
        cnt += ok ? 1ull : 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a.merge_shfl(o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if (lane == 0) {
        out_count[g] = (long long)cnt;
        if constexpr (IsFloat<T>::value) {
            out_sum[g] = (T)a.sum;
        } else {
            using U = typename UnsignedOf<T>::type;
            out_sum[g] = (T)(U)a.sum;   // wrapping
            if (out_min) {
                out_min[g] = cnt ? (T)(U)(a.kmin ^ flip) : (T)0;
                out_max[g] = cnt ? (T)(U)(a.kmax ^ flip) : (T)0;
                if (cnt) atomicOr(&mm_valid[g >> 5], 1u << (g & 31));
            }
        }
    }
}

template <typename T>
static cudaError_t heads_one(const void* key, const uint32_t* kvalid, long long n, uint32_t* words, cudaStream_t s) {
    const long long rows = (n + 31) / 32 * 32;
    k_group_heads<T><<<(unsigned)((rows + kThreads - 1) / kThreads), kThreads, 0, s>>>((const T*)key, kvalid, n, words);
    return cudaGetLastError();
}

cudaError_t launch_group_heads(int dtype, const void* key, const uint32_t* kvalid, long long n, uint32_t* words, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    switch (dtype) {   // equality is bitwise for integers: one instantiation per width
        case T_I8: case T_U8: return heads_one<uint8_t>(key, kvalid, n, words, s);
        case T_I16: case T_U16: return heads_one<uint16_t>(key, kvalid, n, words, s);
        case T_I32: case T_U32: return heads_one<uint32_t>(key, kvalid, n, words, s);
        case T_I64: case T_U64: return heads_one<uint64_t>(key, kvalid, n, words, s);
        case T_F32: return heads_one<float>(key, kvalid, n, words, s);
        case T_F64: return heads_one<double>(key, kvalid, n, words, s);
        default: return cudaErrorInvalidValue;
    }
}

template <typename T>
static cudaError_t reduce_one(const void* val, const uint32_t* vvalid, const uint32_t* starts, long long n_groups, long long n_rows, void* sum,
                              long long* count, void* mn, void* mx, uint32_t* mm_valid, cudaStream_t s) {
    k_group_reduce<T><<<(unsigned)((n_groups + kWarpsPerCta - 1) / kWarpsPerCta), kThreads, 0, s>>>((const T*)val, vvalid, starts, n_groups, n_rows, (T*)sum,
                                                                                                    count, (T*)mn, (T*)mx, mm_valid);
    return cudaGetLastError();
}

cudaError_t launch_group_reduce(int dtype, const void* val, const uint32_t* vvalid, const uint32_t* starts, long long n_groups, long long n_rows,
                                void* sum, long long* count, void* mn, void* mx, uint32_t* mm_valid, cudaStream_t s) {
    if (n_groups <= 0) return cudaSuccess;
    if (n_groups > 0x7fffffffLL * kWarpsPerCta) return cudaErrorInvalidConfiguration;
    switch (dtype) {
        case T_I8: return reduce_one<int8_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_I16: return reduce_one<int16_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_I32: return reduce_one<int32_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_I64: return reduce_one<int64_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_U8: return reduce_one<uint8_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_U16: return reduce_one<uint16_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_U32: return reduce_one<uint32_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_U64: return reduce_one<uint64_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_F32: return reduce_one<float>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        case T_F64: return reduce_one<double>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bdf
