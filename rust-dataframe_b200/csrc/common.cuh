// common.cuh -- device helpers shared by the sm_100a kernels of libb200df.
//
// Data layout in HBM (DESIGN.md "Layout"): every chunk of a device column is an Arrow PrimitiveArray
// whose values start 256-byte aligned (element offset already applied at upload) and whose validity
// bitmap (LSB-first, 1 = valid) is a 4-byte-aligned array of 32-bit words with a residual BIT offset
// `off` (0..7 for uploaded slices, 0 for kernel-made columns).  Bitmaps are padded so that one word past
// the last used one may be read.
//
// Work decomposition: a *tile* is THREADS x U vectors of 16 bytes per array (coalesced: lane l of a warp
// touches vector j*THREADS + l).  A thread therefore owns E = 16/sizeof(T) consecutive elements per step
// and E consecutive validity bits; 32/E neighbouring lanes share one bitmap word, which they assemble
// with a log2(32/E)-step shuffle-OR so the warp writes whole 32-bit words.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace bdf {

enum : int { T_I8 = 0, T_I16, T_I32, T_I64, T_U8, T_U16, T_U32, T_U64, T_F32, T_F64, T_NTYPES };

__host__ __device__ inline int dtype_width(int t) {
    switch (t) {
        case T_I8: case T_U8: return 1;
        case T_I16: case T_U16: return 2;
        case T_I32: case T_U32: case T_F32: return 4;
        default: return 8;
    }
}
__host__ __device__ inline bool dtype_is_float(int t) { return t == T_F32 || t == T_F64; }
__host__ __device__ inline bool dtype_is_signed_int(int t) { return t >= T_I8 && t <= T_I64; }

constexpr int kThreads = 256;  // CTA size of every streaming kernel
constexpr int kUnroll = 4;     // 16-byte vectors in flight per thread per array
constexpr int kTileBytes = kThreads * kUnroll * 16;  // bytes of the widest array per tile (16 KiB)
constexpr int kWarpsPerCta = kThreads / 32;          // kernels report valid-slot counts per warp: [tile][warp] u32

// ---- 128-bit streaming loads/stores -----------------------------------------------------------
// Inputs are read exactly once and outputs written exactly once: bypass L1 allocation so the small
// validity words (which ARE re-read by neighbouring lanes) keep the L1.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream4(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint16_t ld_stream2(const void* p) {
    uint16_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream8(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream2(void* p, uint16_t v) {
    asm volatile("st.global.L1::no_allocate.u16 [%0], %1;" :: "l"(p), "h"(v) : "memory");
}

// A register-resident vector of N elements of T moved with one access of N*sizeof(T) bytes (2..16).
template <typename T, int N>
struct Vec {
    static constexpr int kBytes = N * (int)sizeof(T);
    static_assert(kBytes == 16 || kBytes == 8 || kBytes == 4 || kBytes == 2, "unsupported vector size");
    union {
        T e[N];
        uint4 q;
        uint2 d;
        uint32_t w;
        uint16_t h;
    };
    __device__ __forceinline__ void load(const void* p) {
        if constexpr (kBytes == 16) q = ld_stream16(p);
        else if constexpr (kBytes == 8) d = ld_stream8(p);
        else if constexpr (kBytes == 4) w = ld_stream4(p);
        else h = ld_stream2(p);
    }
    __device__ __forceinline__ void store(void* p) const {
        if constexpr (kBytes == 16) st_stream16(p, q);
        else if constexpr (kBytes == 8) st_stream8(p, d);
        else if constexpr (kBytes == 4) st_stream4(p, w);
        else st_stream2(p, h);
    }
};

// ---- validity bitmap access ----------------------------------------------------------------------
// E consecutive bits starting at absolute bit index `bit` of word array v.
template <int E>
__device__ __forceinline__ uint32_t load_bits(const uint32_t* __restrict__ v, int64_t bit) {
    const int64_t w = bit >> 5;
    const int sh = (int)(bit & 31);
    const uint32_t lo = __ldg(v + w);
    const uint32_t hi = (sh + E > 32) ? __ldg(v + w + 1) : 0u;
    const uint32_t r = __funnelshift_r(lo, hi, sh);
    if constexpr (E == 32) return r;
    else return r & ((1u << E) - 1u);
}

// Batched form used on full tiles: issue the raw word loads of all U steps back to back (no consumer in
// between, so all of them are in flight together with the value loads), extract the bits later.  Wrap the
// issue in ONE uniform `if (bitmap present)`; per-step conditionals would make the compiler serialise the
// loads (load -> wait -> shift, U times), which costs U memory latencies per tile.
template <int E, int U>
struct MaskRaw {
    uint32_t lo[U], hi[U];
    int sh[U];
};
template <int E, int U>
__device__ __forceinline__ void mask_issue(MaskRaw<E, U>& r, const uint32_t* __restrict__ v, int64_t bit0, int64_t stride_bits) {
    // Every caller steps by kThreads * E bits, a whole number of 32-bit words: the shift is the same for all U steps and
    // the word addresses are one base pointer plus constants (the 64-bit shift/LEA chain per step that the general form
    // costs is most of the instruction count of a nullable tile).
    const uint32_t* __restrict__ p = v + (bit0 >> 5);
    const int sh = (int)(bit0 & 31);
    const int stride_words = (int)(stride_bits >> 5);
    const bool two = sh + E > 32;
#pragma unroll
    for (int j = 0; j < U; j++) {
        r.sh[j] = sh;
        r.lo[j] = __ldg(p + j * stride_words);
        r.hi[j] = two ? __ldg(p + j * stride_words + 1) : 0u;
    }
}
// The general form (a word index per step).  k_filter.cu keeps it: with three bitmaps in flight the base-pointer form above
// costs it registers (filter scatter 0.226 -> 0.279 ms at 1e8 f64 rows when it was switched over in round 2).
template <int E, int U>
__device__ __forceinline__ void mask_issue_general(MaskRaw<E, U>& r, const uint32_t* __restrict__ v, int64_t bit0, int64_t stride_bits) {
#pragma unroll
    for (int j = 0; j < U; j++) {
        const int64_t bit = bit0 + (int64_t)j * stride_bits;
        const int64_t w = bit >> 5;
        r.sh[j] = (int)(bit & 31);
        r.lo[j] = __ldg(v + w);
        r.hi[j] = (r.sh[j] + E > 32) ? __ldg(v + w + 1) : 0u;
    }
}
// The same with the word pointer and the shift already known (a loop over consecutive tiles advances the pointer by a constant).
template <int E, int U>
__device__ __forceinline__ void mask_issue_at(MaskRaw<E, U>& r, const uint32_t* __restrict__ p, int sh, int stride_words) {
    const bool two = sh + E > 32;
#pragma unroll
    for (int j = 0; j < U; j++) {
        r.sh[j] = sh;
        r.lo[j] = __ldg(p + j * stride_words);
        r.hi[j] = two ? __ldg(p + j * stride_words + 1) : 0u;
    }
}
template <int E, int U>
__device__ __forceinline__ uint32_t mask_get(const MaskRaw<E, U>& r, int j) {
    const uint32_t x = __funnelshift_r(r.lo[j], r.hi[j], r.sh[j]);
    if constexpr (E == 32) return x;
    else return x & ((1u << E) - 1u);
}

// Each lane contributes E bits for elements [e0, e0+E) (e0 a multiple of E, bit offset 0 on output); the
// 32/E lanes that share a word OR their pieces together and the first of them stores the word.
// Must be executed by all 32 lanes of the warp.  `active` = this lane's e0 lies inside the chunk.
template <int E>
__device__ __forceinline__ void store_bits(uint32_t* __restrict__ vout, int64_t e0, uint32_t bits, bool active) {
    constexpr int G = 32 / E;
    const int lane = threadIdx.x & 31;
    uint32_t c = bits << (E * (lane % G));
#pragma unroll
    for (int o = 1; o < G; o <<= 1) c |= __shfl_xor_sync(0xffffffffu, c, o);
    if ((lane % G) == 0 && active) vout[e0 >> 5] = c;
}

// Mask of the elements of [e0, e0+E) that lie below len.
template <int E>
__device__ __forceinline__ uint32_t tail_mask(int64_t e0, int64_t len) {
    const int64_t rem = len - e0;
    if (rem >= E) return (E == 32) ? 0xffffffffu : ((1u << E) - 1u);
    if (rem <= 0) return 0u;
    return (1u << (int)rem) - 1u;
}

// Chunk lookup: descriptors are sorted by tile0 (first tile of the chunk); returns the chunk owning `tile`.
template <typename Desc>
__device__ __forceinline__ int find_chunk(const Desc* __restrict__ descs, int n_chunks, int64_t tile) {
    int lo = 0, hi = n_chunks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&descs[mid].tile0) <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ---- descriptors (one per chunk, device-resident array sorted by tile0) ------------------------
struct BinDesc {
    const void* a; const void* b; void* out;
    const uint32_t* va; const uint32_t* vb; uint32_t* vout;
    int64_t len; int64_t tile0;
    int32_t offa, offb;
};
struct UnDesc {
    const void* in; void* out;
    const uint32_t* vin; uint32_t* vout;
    int64_t len; int64_t tile0;
    int32_t off; int32_t pad;
};
struct RedDesc {
    const void* in; const uint32_t* vin;
    int64_t len; int64_t tile0;
    int32_t off; int32_t pad;
};
struct GenDesc {
    void* out; uint32_t* vout;
    int64_t len; int64_t tile0; int64_t row0;
};

// Partial / final aggregate as the kernels produce it.
struct AggDev {
    unsigned long long sum_bits;  // ints: 64-bit wrapping sum; floats: double bit pattern
    unsigned long long min_bits;  // ints only: order-preserving unsigned key of the minimum (identity ~0)
    unsigned long long max_bits;  // ints only: order-preserving unsigned key of the maximum (identity 0)
    unsigned long long count;     // valid slots
};

template <typename T> struct UnsignedOf { using type = T; };
template <> struct UnsignedOf<int8_t> { using type = uint8_t; };
template <> struct UnsignedOf<int16_t> { using type = uint16_t; };
template <> struct UnsignedOf<int32_t> { using type = uint32_t; };
template <> struct UnsignedOf<int64_t> { using type = uint64_t; };
template <typename T> struct IsFloat { static constexpr bool value = false; };
template <> struct IsFloat<float> { static constexpr bool value = true; };
template <> struct IsFloat<double> { static constexpr bool value = true; };

// ---- K5: aggregate of the OUTPUT fused into the same pass (sum/min/max/count of c while c is written) ----
// Integers: wrapping 64-bit sum of the zero-extended values (only the low sizeof(T) bytes are meaningful),
// min/max over order-preserving unsigned keys (value ^ sign flip).  Floats: sum in double.  One partial per
// CTA (= per tile) in tile order; k_finish folds them in a fixed order => deterministic.
template <typename T, bool F = IsFloat<T>::value> struct FusedAgg;
template <typename T> struct FusedAgg<T, false> {
    unsigned long long sum, kmin, kmax;
    __device__ __forceinline__ void init() { sum = 0; kmin = ~0ull; kmax = 0; }
    __device__ __forceinline__ void add(T x, bool valid, unsigned long long flip) {
        using U = typename UnsignedOf<T>::type;
        const unsigned long long z = (unsigned long long)(U)x;
        if (valid) { sum += z; const unsigned long long k = z ^ flip; kmin = k < kmin ? k : kmin; kmax = k > kmax ? k : kmax; }
    }
    __device__ __forceinline__ void merge_shfl(int o) {
        const unsigned long long s2 = __shfl_xor_sync(0xffffffffu, sum, o), a2 = __shfl_xor_sync(0xffffffffu, kmin, o),
                                 b2 = __shfl_xor_sync(0xffffffffu, kmax, o);
        sum += s2; kmin = a2 < kmin ? a2 : kmin; kmax = b2 > kmax ? b2 : kmax;
    }
    __device__ __forceinline__ void merge(const FusedAgg& o) { sum += o.sum; kmin = o.kmin < kmin ? o.kmin : kmin; kmax = o.kmax > kmax ? o.kmax : kmax; }
    __device__ __forceinline__ void store(AggDev* d, unsigned long long cnt) const { d->sum_bits = sum; d->min_bits = kmin; d->max_bits = kmax; d->count = cnt; }
};
template <typename T> struct FusedAgg<T, true> {
    double sum;
    __device__ __forceinline__ void init() { sum = 0.0; }
    __device__ __forceinline__ void add(T x, bool valid, unsigned long long) { sum = __dadd_rn(sum, valid ? (double)x : 0.0); }
    __device__ __forceinline__ void merge_shfl(int o) { sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o)); }
    __device__ __forceinline__ void merge(const FusedAgg& o) { sum = __dadd_rn(sum, o.sum); }
    __device__ __forceinline__ void store(AggDev* d, unsigned long long cnt) const {
        d->sum_bits = (unsigned long long)__double_as_longlong(sum); d->min_bits = ~0ull; d->max_bits = 0; d->count = cnt;
    }
};

// ---- launchers (defined in k_*.cu) ---------------------------------------------------------------
int elems_per_tile(int dtype);                 // tile size in elements for arrays of dtype
int elems_per_tile_binary(int op, int dtype);
int elems_per_tile_unary(int op, int dtype);   // K2 tile size (libm-class functions use shorter tiles)  // K1 tile size (divide/libm binaries use shorter tiles)
int elems_per_tile_cast(int from, int to);
// tile_partials != nullptr (add/sub/mul/div only): also write one AggDev per tile with the aggregate of the
// OUTPUT (K5); fold them with launch_finish.  Integer min/max partials are unsigned keys (value ^ sign flip).
cudaError_t launch_binary(int op, int dtype, const BinDesc* d_descs, int n_chunks, int64_t total_tiles,
                          uint32_t* d_warp_counts, int* d_flags, cudaStream_t s, AggDev* tile_partials = nullptr);
cudaError_t launch_finish(bool is_float, const AggDev* parts, int64_t n_parts, int sm_count, AggDev* stage, unsigned int* ticket,
                          AggDev* result, cudaStream_t s);
cudaError_t launch_unary(int op, int dtype, const UnDesc* d_descs, int n_chunks, int64_t total_tiles,
                         uint32_t* d_warp_counts, cudaStream_t s);
cudaError_t launch_cast(int from, int to, const UnDesc* d_descs, int n_chunks, int64_t total_tiles,
                        uint32_t* d_warp_counts, cudaStream_t s);
// K4: one AggDev partial per CTA (integer min/max as unsigned keys: extended value ^ 2^63 for signed types);
// fold with launch_finish.
cudaError_t launch_reduce(int dtype, const RedDesc* d_descs, int n_chunks, int64_t total_tiles, AggDev* d_cta_partials,
                          cudaStream_t s);
int64_t reduce_partials(int dtype, int64_t tiles);  // number of partials launch_reduce writes
// Several columns of ONE dtype can share a launch: concatenate their chunk descriptors and start every column's tile
// numbering on a multiple of reduce_tiles_per_cta(dtype) -- a CTA then never mixes columns, and column k's partials are the
// CTA range [tile0_k / K, ceil(tile_end_k / K)).
int reduce_tiles_per_cta(int dtype);
// Fold up to kFinishMany partial ranges in ONE launch (grid.y = range); stage: n x sm_count records, tickets: n zeroed counters.
constexpr int kFinishMany = 64;
struct FinishJob { const AggDev* parts; long long n_parts; AggDev* result; int is_float; int pad; };
cudaError_t launch_finish_many(int n, const FinishJob* jobs, int sm_count, AggDev* stage, unsigned int* tickets, cudaStream_t s);
cudaError_t launch_generate(int dtype, int kind, double lo, double hi, uint64_t seed, uint64_t col, uint32_t null_mod,
                            const GenDesc* d_descs, int n_chunks, int64_t total_tiles,
                            uint32_t* d_warp_counts, cudaStream_t s);
cudaError_t launch_fill(void* p, size_t bytes, cudaStream_t s);

}  // namespace bdf
