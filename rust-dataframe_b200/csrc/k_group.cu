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
//                   NULL for a group without a valid value.  A group of more than kBigGroup rows is not folded by its warp:
//                   the warp registers it in a work list and k_group_big folds it one CTA per 64 Ki-row segment (thread-strided
//                   fold + fixed block tree), k_group_big_finish then folds a group's segment partials in segment order --
//                   still deterministic, and a single hot key is spread over the whole GPU instead of one warp.
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

// Hot keys.  The work list lives in device memory: {counters, list[]} then the segment partials.
constexpr long long kBigGroup = 65536;   // rows a single warp still folds itself; also the segment length of k_group_big
struct BigGroup { unsigned long long g, begin, end; unsigned int seg0, n_seg; };
struct BigWork { unsigned int n_groups, n_segments; BigGroup list[1]; };   // list[capacity] follows

template <typename T>
__device__ __forceinline__ void group_write(const FusedAgg<T>& a, unsigned long long cnt, long long g, T* out_sum, long long* out_count, T* out_min,
                                            T* out_max, uint32_t* mm_valid, unsigned long long flip) {
    out_count[g] = (long long)cnt;
    if constexpr (IsFloat<T>::value) {
        out_sum[g] = (T)a.sum;
    } else {
        using U = typename UnsignedOf<T>::type;
        out_sum[g] = (T)(U)a.sum;   // wrapping
        if (out_min) {
            out_min[g] = cnt ? (T)(U)(a.kmin ^ flip) : (T)0;
            out_max[g] = cnt ? (T)(U)(a.kmax ^ flip) : (T)0;
            if (cnt && mm_valid) atomicOr(&mm_valid[g >> 5], 1u << (g & 31));
        }
    }
}

template <typename T> struct GroupFlip {
    static constexpr unsigned long long value = IsFloat<T>::value ? 0ull : (((T)-1 < (T)0) ? (1ull << (8 * sizeof(T) - 1)) : 0ull);
};

// One CTA per (hot group, segment): thread t folds rows begin + t, begin + t + 256, ... of the segment, then warp trees, then the
// warps in order -> partial[seg0 + segment].  The grid is fixed; CTAs walk the work items (the counters are only known on the device).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_group_big(const T* __restrict__ val, const uint32_t* __restrict__ vvalid, const BigWork* __restrict__ big, FusedAgg<T>* __restrict__ part,
            unsigned long long* __restrict__ part_cnt) {
    __shared__ FusedAgg<T> s_a[kWarpsPerCta];
    __shared__ unsigned long long s_c[kWarpsPerCta];
    const unsigned int n_big = big->n_groups, n_seg = big->n_segments;
    for (unsigned int w = blockIdx.x; w < n_seg; w += gridDim.x) {
        unsigned int gi = 0;
        while (gi < n_big && !(big->list[gi].seg0 <= w && w < big->list[gi].seg0 + big->list[gi].n_seg)) gi++;   // a handful of hot groups
        if (gi == n_big) continue;
        const BigGroup bg = big->list[gi];
        const long long b = (long long)bg.begin + (long long)(w - bg.seg0) * kBigGroup;
        const long long e = min((long long)bg.end, b + kBigGroup);
        FusedAgg<T> a;
        a.init();
        unsigned long long cnt = 0;
        for (long long i = b + threadIdx.x; i < e; i += kThreads) {
            const bool ok = vvalid ? ((vvalid[i >> 5] >> (i & 31)) & 1u) : true;
            a.add(val[i], ok, GroupFlip<T>::value);
            cnt += ok ? 1ull : 0ull;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a.merge_shfl(o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
        __syncthreads();   // the previous work item's readers are done with the shared arrays
        if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_c[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            FusedAgg<T> t = s_a[0];
            unsigned long long total = s_c[0];
#pragma unroll
            for (int k = 1; k < kWarpsPerCta; k++) { t.merge(s_a[k]); total += s_c[k]; }
            part[w] = t;
            part_cnt[w] = total;
        }
    }
}

// One thread per hot group: its segment partials in segment order.
template <typename T>
__global__ void k_group_big_finish(const BigWork* __restrict__ big, const FusedAgg<T>* __restrict__ part, const unsigned long long* __restrict__ part_cnt,
                                   T* __restrict__ out_sum, long long* __restrict__ out_count, T* __restrict__ out_min, T* __restrict__ out_max,
                                   uint32_t* __restrict__ mm_valid) {
    const unsigned int gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= big->n_groups) return;
    const BigGroup bg = big->list[gi];
    FusedAgg<T> a = part[bg.seg0];
    unsigned long long cnt = part_cnt[bg.seg0];
    for (unsigned int k = 1; k < bg.n_seg; k++) { a.merge(part[bg.seg0 + k]); cnt += part_cnt[bg.seg0 + k]; }
    group_write<T>(a, cnt, (long long)bg.g, out_sum, out_count, out_min, out_max, mm_valid, GroupFlip<T>::value);
}

// LANES lanes per group (32, 8 or 1 -- the launcher picks by the average group length, so ~1e8 one-row groups do not spend a
// warp each).  No lane leaves before the shuffles: a sub-group that is past the end, or defers its group, folds an empty range.
template <typename T, int LANES>
__global__ void __launch_bounds__(kThreads)
k_group_reduce(const T* __restrict__ val, const uint32_t* __restrict__ vvalid, const uint32_t* __restrict__ starts, long long n_groups,
               long long n_rows, T* __restrict__ out_sum, long long* __restrict__ out_count, T* __restrict__ out_min, T* __restrict__ out_max,
               uint32_t* __restrict__ mm_valid, BigWork* __restrict__ big) {
    constexpr int kGroupsPerWarp = 32 / LANES;
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long g = t / LANES;
    const int sub = threadIdx.x & (LANES - 1);
    bool mine = g < n_groups;
    long long b = 0, e = 0;
    if (mine) { b = starts[g]; e = g + 1 < n_groups ? (long long)starts[g + 1] : n_rows; }
    if (mine && big && e - b > kBigGroup) {   // a hot key: leave it to k_group_big (one CTA per segment)
        if (sub == 0) {
            const unsigned int nseg = (unsigned int)((e - b + kBigGroup - 1) / kBigGroup);
            const unsigned int slot = atomicAdd(&big->n_groups, 1u);
            const unsigned int base = atomicAdd(&big->n_segments, nseg);
            big->list[slot] = BigGroup{(unsigned long long)g, (unsigned long long)b, (unsigned long long)e, base, nseg};
        }
        mine = false;
        e = b;
    }
    FusedAgg<T> a;
    a.init();
    unsigned long long cnt = 0;
    constexpr unsigned long long flip = GroupFlip<T>::value;
    for (long long i = b + sub; i < e; i += LANES) {
        const bool ok = vvalid ? ((vvalid[i >> 5] >> (i & 31)) & 1u) : true;
        a.add(val[i], ok, flip);
        cnt += ok ? 1ull : 0ull;
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) { a.merge_shfl(o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    const bool writer = mine && sub == 0;
    if (writer) group_write<T>(a, cnt, g, out_sum, out_count, out_min, out_max, nullptr, flip);
    if constexpr (!IsFloat<T>::value) {
        if (out_min) {   // min/max validity: the warp's groups share one bitmap word -> one store (LANES == 1) or one atomic per warp
            const unsigned bal = __ballot_sync(0xffffffffu, writer && cnt != 0);
            if ((threadIdx.x & 31) == 0) {
                const long long g0 = (t >> 5) * kGroupsPerWarp;
                if (g0 < n_groups) {
                    unsigned word = 0;
#pragma unroll
                    for (int j = 0; j < kGroupsPerWarp; j++) word |= ((bal >> (j * LANES)) & 1u) << (((unsigned)g0 + j) & 31u);
                    if (LANES == 1) mm_valid[g0 >> 5] = word;   // hot groups of this word are OR-ed in later by k_group_big_finish
                    else if (word) atomicOr(&mm_valid[g0 >> 5], word);
                }
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

size_t group_big_scratch_bytes(long long n_rows) {   // work list + segment partials for any distribution of n_rows rows
    const size_t max_groups = (size_t)(n_rows / kBigGroup) + 2, max_segs = 2 * max_groups + 2;
    return 64 + max_groups * sizeof(BigGroup) + max_segs * (sizeof(FusedAgg<long long>) + 8) + 256;
}

template <typename T>
static cudaError_t reduce_one(const void* val, const uint32_t* vvalid, const uint32_t* starts, long long n_groups, long long n_rows, void* sum,
                              long long* count, void* mn, void* mx, uint32_t* mm_valid, void* scratch, int sm_count, cudaStream_t s) {
    BigWork* big = (BigWork*)scratch;
    cudaError_t e = cudaMemsetAsync(big, 0, 8, s);   // the two counters
    if (e != cudaSuccess) return e;
    const long long avg = n_rows / n_groups;   // lanes per group by the average group length
    const int lanes = avg > 64 ? 32 : avg > 4 ? 8 : 1;
    const unsigned grid = (unsigned)((n_groups * lanes + kThreads - 1) / kThreads);
    if (lanes == 32) k_group_reduce<T, 32><<<grid, kThreads, 0, s>>>((const T*)val, vvalid, starts, n_groups, n_rows, (T*)sum, count, (T*)mn, (T*)mx, mm_valid, big);
    else if (lanes == 8) k_group_reduce<T, 8><<<grid, kThreads, 0, s>>>((const T*)val, vvalid, starts, n_groups, n_rows, (T*)sum, count, (T*)mn, (T*)mx, mm_valid, big);
    else k_group_reduce<T, 1><<<grid, kThreads, 0, s>>>((const T*)val, vvalid, starts, n_groups, n_rows, (T*)sum, count, (T*)mn, (T*)mx, mm_valid, big);
    if (n_rows > kBigGroup) {   // hot keys are possible: the two follow-up kernels find nothing to do when there are none
        const size_t max_groups = (size_t)(n_rows / kBigGroup) + 2, max_segs = 2 * max_groups + 2;
        char* p = (char*)scratch + 64 + max_groups * sizeof(BigGroup);
        p = (char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
        FusedAgg<T>* part = (FusedAgg<T>*)p;
        unsigned long long* part_cnt = (unsigned long long*)(p + max_segs * sizeof(FusedAgg<long long>));
        k_group_big<T><<<(unsigned)(sm_count * 4), kThreads, 0, s>>>((const T*)val, vvalid, big, part, part_cnt);
        k_group_big_finish<T><<<(unsigned)((max_groups + 127) / 128), 128, 0, s>>>(big, part, part_cnt, (T*)sum, count, (T*)mn, (T*)mx, mm_valid);
    }
    return cudaGetLastError();
}

// scratch: group_big_scratch_bytes(n_rows) of device memory.
cudaError_t launch_group_reduce(int dtype, const void* val, const uint32_t* vvalid, const uint32_t* starts, long long n_groups, long long n_rows,
                                void* sum, long long* count, void* mn, void* mx, uint32_t* mm_valid, void* scratch, int sm_count, cudaStream_t s) {
    if (n_groups <= 0) return cudaSuccess;
    if (n_groups > 0x7fffffffLL * kWarpsPerCta) return cudaErrorInvalidConfiguration;
    switch (dtype) {
        case T_I8: return reduce_one<int8_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_I16: return reduce_one<int16_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_I32: return reduce_one<int32_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_I64: return reduce_one<int64_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_U8: return reduce_one<uint8_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_U16: return reduce_one<uint16_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_U32: return reduce_one<uint32_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_U64: return reduce_one<uint64_t>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_F32: return reduce_one<float>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        case T_F64: return reduce_one<double>(val, vvalid, starts, n_groups, n_rows, sum, count, mn, mx, mm_valid, scratch, sm_count, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bdf
