// k_generate.cu -- counter-based synthetic columns written straight into HBM (bench / large-config tests).
// Not part of the reference's path: it exists because SURVEY 8(d) configs 2-5 hold 1e8..1e9 rows, which
// must be produced on the device.  Bit-for-bit equal to oracle/oracle.c:orc_generate (tests check it):
//   h = splitmix64(seed ^ (col << 56) ^ row);  u = (h >> 11) * 2^-53
//   kind 0: lo + (hi-lo)*u (two roundings)   kind 1: +-(1+u), sign = h&1
//   kind 2: (int64)h                          kind 3: (int64)(h >> 23) - 2^40
//   null  : null_mod != 0 && splitmix64(h) % null_mod == 0
#include "common.cuh"

namespace bdf {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T, bool IsFloat>
__device__ __forceinline__ T gen_value(int kind, double lo, double span, uint64_t h) {
    const bool is_real = kind <= 1;
    double f = 0.0;
    long long iv = 0;
    if (kind == 0) {
        const double u = __dmul_rn((double)(h >> 11), 0x1.0p-53);
        f = __dadd_rn(lo, __dmul_rn(span, u));
    } else if (kind == 1) {
        const double u = __dmul_rn((double)(h >> 11), 0x1.0p-53);
        f = __dadd_rn(1.0, u);
        if (h & 1ull) f = -f;
    } else if (kind == 2) {
        iv = (long long)h;
    } else {
        iv = (long long)(h >> 23) - (1ll << 40);
    }
    if constexpr (IsFloat) return is_real ? (T)f : (T)iv;
    else return (T)(is_real ? (long long)f : iv);
}

template <typename T, bool IsFloat>
__global__ void __launch_bounds__(kThreads)
k_generate(const GenDesc* __restrict__ descs, int n_chunks, int kind, double lo, double span, uint64_t seed,
           uint64_t colbits, uint32_t null_mod, uint32_t* __restrict__ warp_counts) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int TILE = kThreads * kUnroll * E;
    const int64_t tile = blockIdx.x;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    T* __restrict__ po = (T*)descs[c].out;
    uint32_t* __restrict__ vo = descs[c].vout;
    const int64_t len = descs[c].len;
    const int64_t row0 = descs[c].row0;
    const int64_t base = (tile - descs[c].tile0) * TILE;
    unsigned int nvalid = 0;
#pragma unroll 1
    for (int j = 0; j < kUnroll; j++) {
        const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
        const uint32_t in_range = tail_mask<E>(e0, len);
        Vec<T, E> r;
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < E; e++) {
            const uint64_t h = splitmix64(seed ^ colbits ^ (uint64_t)(row0 + e0 + e));
            r.e[e] = gen_value<T, IsFloat>(kind, lo, span, h);
            const bool valid = (null_mod == 0u) || (splitmix64(h) % null_mod) != 0ull;
            m |= (valid ? 1u : 0u) << e;
        }
        m &= in_range;
        if (in_range == ((1u << E) - 1u)) r.store(po + e0);
        else {
#pragma unroll
            for (int e = 0; e < E; e++) if ((in_range >> e) & 1u) po[e0 + e] = r.e[e];
        }
        if (vo) {
            store_bits<E>(vo, e0, m, in_range != 0);
            nvalid += __popc(m);
        }
    }
    if (vo) {
        const unsigned int wvalid = __reduce_add_sync(0xffffffffu, nvalid);
        if ((threadIdx.x & 31) == 0) warp_counts[(int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)] = wvalid;
    }
}

__global__ void __launch_bounds__(kThreads) k_fill(uint4* __restrict__ p, size_t n_vec, uint32_t v) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
        st_stream16(p + i, make_uint4(v, v, v, v));
}

cudaError_t launch_fill(void* p, size_t bytes, cudaStream_t s) {
    const size_t n_vec = bytes / 16;
    if (n_vec == 0) return cudaSuccess;
    k_fill<<<148 * 8, kThreads, 0, s>>>((uint4*)p, n_vec, 0x5a5a5a5au);
    return cudaGetLastError();
}

template <typename T, bool F>
static cudaError_t launch_one(int kind, double lo, double hi, uint64_t seed, uint64_t col, uint32_t null_mod,
                              const GenDesc* d, int n, int64_t tiles, uint32_t* vc, cudaStream_t s) {
    k_generate<T, F><<<(unsigned)tiles, kThreads, 0, s>>>(d, n, kind, lo, hi - lo, seed, col << 56, null_mod, vc);
    return cudaGetLastError();
}

cudaError_t launch_generate(int dtype, int kind, double lo, double hi, uint64_t seed, uint64_t col, uint32_t null_mod,
                            const GenDesc* d, int n, int64_t tiles, uint32_t* vc, cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    switch (dtype) {
        case T_F64: return launch_one<double, true>(kind, lo, hi, seed, col, null_mod, d, n, tiles, vc, s);
        case T_F32: return launch_one<float, true>(kind, lo, hi, seed, col, null_mod, d, n, tiles, vc, s);
        case T_I8: case T_U8: return launch_one<uint8_t, false>(kind, lo, hi, seed, col, null_mod, d, n, tiles, vc, s);
        case T_I16: case T_U16: return launch_one<uint16_t, false>(kind, lo, hi, seed, col, null_mod, d, n, tiles, vc, s);
        case T_I32: case T_U32: return launch_one<uint32_t, false>(kind, lo, hi, seed, col, null_mod, d, n, tiles, vc, s);
        case T_I64: case T_U64: return launch_one<uint64_t, false>(kind, lo, hi, seed, col, null_mod, d, n, tiles, vc, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bdf
