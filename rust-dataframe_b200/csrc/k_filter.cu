// k_filter.cu -- SURVEY 8(f) N2, the step right after the hot path in a lazy pipeline:
//   BooleanFilter::{Gt,Ge,Eq,Ne,Lt,Le}   reference src/expression.rs:820-852 -> arrow compute::{gt,gt_eq,eq,neq,lt,lt_eq}
//                                         on Float64Arrays (both sides are cast to Float64 first)      -> k_compare
//   compute::{and,or,not} on BooleanArrays (expression.rs:803-819; values op values, validity AND)    -> k_boolean
//   ChunkedArray::filter -> arrow compute::filter per chunk (src/table.rs:97-107)                      -> k_filter_*
// Boolean columns are Arrow BooleanArrays: bit-packed values (LSB first) + optional validity bitmap.
//
// Filter = stream compaction, three launches over ALL chunks of the column:
//   1. k_filter_count   one warp per tile: popcount of (mask values AND mask validity)       (reads bitmaps only)
//   2. k_filter_scan    one CTA per chunk: exclusive scan of its tile counts -> tile offsets, chunk totals
//   3. k_filter_scatter one CTA per tile: per-step block scan of the per-thread counts, selected values are
//      written in order at out[tile offset + rank]; kept validity bits are compacted per 32-slot word
//      (each lane compacts its E bits, the 32/E lanes of a word OR them together) and OR-ed into the output
//      bitmap with at most two atomics per 32 input slots.
// Roofline: HBM, w x (1 + selectivity) + bitmaps bytes/row for the scatter; 16 B/row for a compare.
#include <cmath>
#include <type_traits>

#include "common.cuh"

namespace bdf {

enum : int { CMP_GT = 0, CMP_GE, CMP_EQ, CMP_NE, CMP_LT, CMP_LE };
enum : int { BOOL_AND = 0, BOOL_OR, BOOL_NOT };

template <int OP>
__device__ __forceinline__ bool cmp_apply(double a, double b) {
    if constexpr (OP == CMP_GT) return a > b;
    else if constexpr (OP == CMP_GE) return a >= b;
    else if constexpr (OP == CMP_EQ) return a == b;
    else if constexpr (OP == CMP_NE) return a != b;
    else if constexpr (OP == CMP_LT) return a < b;
    else return a <= b;
}

// ---- compare: cast(left, Float64) (op) cast(right, Float64) | scalar -> boolean values + validity ------------------
// The reference materialises both casts (arrow::compute::cast(.., Float64), expression.rs:823-824) and then compares
// the Float64Arrays; here the cast is fused into the load (any of the 10 numeric types, 2 elements per lane and step,
// fully coalesced), so an Int32 column costs 4 B/row instead of 4 + 8 (cast) + 8 (compare) B/row.  Values are
// computed under nulls exactly as the two-step form would: a non-Float64 null slot compares as 0.0 (the cast's
// payload), a Float64 input is used as is.
// All kUnroll steps of one input in two phases, each under ONE uniform switch on the column type: raw_issue only
// issues the loads (2 elements per lane and step, 2..16 bytes, coalesced), raw_convert turns them into doubles later,
// so the loads of both inputs and the validity words are all in flight before anything waits.
template <int U>
__device__ __forceinline__ void raw_issue(int t, const void* __restrict__ p, int64_t e_first, int64_t stride, uint4 (&q)[U]) {
    switch (dtype_width(t)) {
        case 8:
#pragma unroll
            for (int j = 0; j < U; j++) q[j] = ld_stream16((const char*)p + (e_first + (int64_t)j * stride) * 8);
            break;
        case 4:
#pragma unroll
            for (int j = 0; j < U; j++) { const uint2 d = ld_stream8((const char*)p + (e_first + (int64_t)j * stride) * 4); q[j].x = d.x; q[j].y = d.y; }
            break;
        case 2:
#pragma unroll
            for (int j = 0; j < U; j++) q[j].x = ld_stream4((const char*)p + (e_first + (int64_t)j * stride) * 2);
            break;
        default:
#pragma unroll
            for (int j = 0; j < U; j++) q[j].x = ld_stream2((const char*)p + (e_first + (int64_t)j * stride));
            break;
    }
}
template <int U>
__device__ __forceinline__ void raw_convert(int t, const uint4 (&q)[U], double (&x)[U][2]) {
#define BDF_CONV(C0, C1) _Pragma("unroll") for (int j = 0; j < U; j++) { x[j][0] = C0; x[j][1] = C1; } break;
    switch (t) {
        case T_F64: BDF_CONV(__hiloint2double((int)q[j].y, (int)q[j].x), __hiloint2double((int)q[j].w, (int)q[j].z))
        case T_I64: BDF_CONV((double)(long long)(((unsigned long long)q[j].y << 32) | q[j].x), (double)(long long)(((unsigned long long)q[j].w << 32) | q[j].z))
        case T_U64: BDF_CONV((double)(((unsigned long long)q[j].y << 32) | q[j].x), (double)(((unsigned long long)q[j].w << 32) | q[j].z))
        case T_F32: BDF_CONV((double)__uint_as_float(q[j].x), (double)__uint_as_float(q[j].y))
        case T_I32: BDF_CONV((double)(int)q[j].x, (double)(int)q[j].y)
        case T_U32: BDF_CONV((double)q[j].x, (double)q[j].y)
        case T_I16: BDF_CONV((double)(short)(q[j].x & 0xffffu), (double)(short)(q[j].x >> 16))
        case T_U16: BDF_CONV((double)(q[j].x & 0xffffu), (double)(q[j].x >> 16))
        case T_I8: BDF_CONV((double)(signed char)(q[j].x & 0xffu), (double)(signed char)((q[j].x >> 8) & 0xffu))
        default: BDF_CONV((double)(q[j].x & 0xffu), (double)((q[j].x >> 8) & 0xffu))
    }
#undef BDF_CONV
}
__device__ __forceinline__ double load1_as_f64(int t, const void* __restrict__ p, int64_t i) {
    switch (t) {
        case T_F64: return ((const double*)p)[i];
        case T_I64: return (double)((const long long*)p)[i];
        case T_U64: return (double)((const unsigned long long*)p)[i];
        case T_F32: return (double)((const float*)p)[i];
        case T_I32: return (double)((const int*)p)[i];
        case T_U32: return (double)((const unsigned*)p)[i];
        case T_I16: return (double)((const short*)p)[i];
        case T_U16: return (double)((const unsigned short*)p)[i];
        case T_I8: return (double)((const signed char*)p)[i];
        default: return (double)((const unsigned char*)p)[i];
    }
}

template <int OP, bool SCALAR>
__global__ void __launch_bounds__(kThreads)
k_compare(const BinDesc* __restrict__ descs, int n_chunks, int ta, int tb, double scalar, uint32_t* __restrict__ warp_counts) {
    constexpr int E = 2;
    constexpr int TILE = kThreads * kUnroll * E;
    const int64_t tile = blockIdx.x;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const void* __restrict__ pa = descs[c].a;
    const void* __restrict__ pb = descs[c].b;
    uint32_t* __restrict__ po = (uint32_t*)descs[c].out;   // boolean VALUES bitmap
    const uint32_t* __restrict__ va = descs[c].va;
    const uint32_t* __restrict__ vb = descs[c].vb;
    uint32_t* __restrict__ vo = descs[c].vout;
    const int64_t len = descs[c].len;
    const int64_t offa = descs[c].offa, offb = descs[c].offb;
    const int64_t base = (tile - descs[c].tile0) * TILE;
    const bool zero_a = va && ta != T_F64, zero_b = !SCALAR && vb && tb != T_F64;  // cast payload of null slots
    unsigned int nvalid = 0;
    if (base + TILE <= len) {
        double a[kUnroll][E], b[kUnroll][E];
        const int64_t e_first = base + (int64_t)threadIdx.x * E;
        uint4 qa[kUnroll], qb[kUnroll];
        raw_issue<kUnroll>(ta, pa, e_first, (int64_t)kThreads * E, qa);
        if constexpr (!SCALAR) raw_issue<kUnroll>(tb, pb, e_first, (int64_t)kThreads * E, qb);
        MaskRaw<E, kUnroll> ra, rb;
        if (va) mask_issue_general<E, kUnroll>(ra, va, offa + e_first, (int64_t)kThreads * E);
        if (vb) mask_issue_general<E, kUnroll>(rb, vb, offb + e_first, (int64_t)kThreads * E);
        raw_convert<kUnroll>(ta, qa, a);
        if constexpr (!SCALAR) raw_convert<kUnroll>(tb, qb, b);
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            const uint32_t ma = va ? mask_get<E, kUnroll>(ra, j) : 3u;
            const uint32_t mb = vb ? mask_get<E, kUnroll>(rb, j) : 3u;
            uint32_t bits = 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
                const double x = (zero_a && !((ma >> e) & 1u)) ? 0.0 : a[j][e];
                const double y = SCALAR ? scalar : ((zero_b && !((mb >> e) & 1u)) ? 0.0 : b[j][e]);
                bits |= (cmp_apply<OP>(x, y) ? 1u : 0u) << e;
            }
            store_bits<E>(po, e0, bits, true);
            if (vo) {
                const uint32_t m = ma & mb;
                store_bits<E>(vo, e0, m, true);
                nvalid += __popc(m);
            }
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            const uint32_t in_range = tail_mask<E>(e0, len);
            uint32_t ma = in_range, mb = in_range;
            if (in_range) {
                if (va) ma &= load_bits<E>(va, offa + e0);
                if (vb) mb &= load_bits<E>(vb, offb + e0);
            }
            uint32_t bits = 0;
#pragma unroll
            for (int e = 0; e < E; e++)
                if ((in_range >> e) & 1u) {
                    const double x = (zero_a && !((ma >> e) & 1u)) ? 0.0 : load1_as_f64(ta, pa, e0 + e);
                    const double y = SCALAR ? scalar : ((zero_b && !((mb >> e) & 1u)) ? 0.0 : load1_as_f64(tb, pb, e0 + e));
                    bits |= (cmp_apply<OP>(x, y) ? 1u : 0u) << e;
                }
            store_bits<E>(po, e0, bits, in_range != 0);
            if (vo) {
                const uint32_t m = ma & mb;
                store_bits<E>(vo, e0, m, in_range != 0);
                nvalid += __popc(m);
            }
        }
    }
    if (vo) {
        const unsigned int wvalid = __reduce_add_sync(0xffffffffu, nvalid);
        if ((threadIdx.x & 31) == 0) warp_counts[(int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)] = wvalid;
    }
}

// ---- and / or / not on boolean columns: one 32-slot word per thread and step ------------------------------------
// BinDesc reuse: a/b = VALUES bitmaps of the inputs (offa/offb = their bit offsets), va/vb = validity bitmaps whose
// bit offsets travel in the (otherwise unused for this kernel) high halves: see BoolDesc below.
struct BoolDesc {
    const uint32_t* a; const uint32_t* b; uint32_t* out;
    const uint32_t* va; const uint32_t* vb; uint32_t* vout;
    int64_t len; int64_t tile0;
    int32_t offa, offb, voffa, voffb;
};
constexpr int kBoolTile = kThreads * kUnroll * 32;  // slots per tile

template <int OP>
__global__ void __launch_bounds__(kThreads)
k_boolean(const BoolDesc* __restrict__ descs, int n_chunks, uint32_t* __restrict__ warp_counts) {
    const int64_t tile = blockIdx.x;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const BoolDesc d = descs[c];
    const int64_t base = (tile - d.tile0) * kBoolTile;
    unsigned int nvalid = 0;
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * 32;
        const uint32_t in_range = tail_mask<32>(e0, d.len);
        if (in_range) {
            const uint32_t x = load_bits<32>(d.a, d.offa + e0);
            const uint32_t y = (OP == BOOL_NOT) ? 0u : load_bits<32>(d.b, d.offb + e0);
            const uint32_t r = (OP == BOOL_AND) ? (x & y) : (OP == BOOL_OR) ? (x | y) : ~x;
            d.out[e0 >> 5] = r & in_range;
            if (d.vout) {
                uint32_t m = in_range;
                if (d.va) m &= load_bits<32>(d.va, d.voffa + e0);
                if (OP != BOOL_NOT && d.vb) m &= load_bits<32>(d.vb, d.voffb + e0);
                d.vout[e0 >> 5] = m;
                nvalid += __popc(m);
            }
        }
    }
    if (d.vout) {
        const unsigned int wvalid = __reduce_add_sync(0xffffffffu, nvalid);
        if ((threadIdx.x & 31) == 0) warp_counts[(int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)] = wvalid;
    }
}

// ---- filter -----------------------------------------------------------------------------------------------------
struct FilterDesc {
    const void* in; void* out;                     // values
    const uint32_t* vin; uint32_t* vout;           // validity of the values (vout zero-initialised)
    const uint32_t* mval; const uint32_t* mvalid;  // mask values / mask validity (nullptr = all valid)
    int64_t len; int64_t tile0;
    int32_t off, moff, mvoff, pad;
};

// 1. one warp per tile: selected slots of the tile
__global__ void __launch_bounds__(kThreads)
k_filter_count(const FilterDesc* __restrict__ descs, int n_chunks, int64_t total_tiles, int tile_elems, unsigned int* __restrict__ tile_counts) {
    const int64_t tile = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (tile >= total_tiles) return;
    const int lane = threadIdx.x & 31;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const uint32_t* __restrict__ mval = descs[c].mval;
    const uint32_t* __restrict__ mvalid = descs[c].mvalid;
    const int64_t len = descs[c].len;
    const int64_t moff = descs[c].moff, mvoff = descs[c].mvoff;
    const int64_t base = (tile - descs[c].tile0) * tile_elems;
    unsigned int n = 0;
    for (int64_t e0 = base + (int64_t)lane * 32; e0 < base + tile_elems && e0 < len; e0 += 32 * 32) {
        uint32_t sel = load_bits<32>(mval, moff + e0) & tail_mask<32>(e0, len);
        if (mvalid) sel &= load_bits<32>(mvalid, mvoff + e0);
        n += __popc(sel);
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) tile_counts[tile] = n;
}

// 2. one CTA per chunk: exclusive scan of the chunk's tile counts
__global__ void __launch_bounds__(kThreads)
k_filter_scan(const FilterDesc* __restrict__ descs, int tile_elems, const unsigned int* __restrict__ tile_counts,
              long long* __restrict__ tile_offsets, long long* __restrict__ chunk_totals) {
    __shared__ long long s_warp[kWarpsPerCta];
    __shared__ long long s_carry;
    const int c = blockIdx.x;
    const int64_t t0 = descs[c].tile0;
    const int64_t nt = (descs[c].len + tile_elems - 1) / tile_elems;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i0 = 0; i0 < nt; i0 += kThreads) {
        const int64_t i = i0 + threadIdx.x;
        const long long v = i < nt ? (long long)tile_counts[t0 + i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        long long wbase = 0;
        for (int w = 0; w < warp; w++) wbase += s_warp[w];
        const long long carry = s_carry;
        if (i < nt) tile_offsets[t0 + i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_totals[c] = s_carry;
}

// 3. scatter
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_filter_scatter(const FilterDesc* __restrict__ descs, int n_chunks, const long long* __restrict__ tile_offsets) {
    constexpr int E = 16 / (int)sizeof(T);
    constexpr int G = 32 / E;  // lanes per 32-slot word
    constexpr int TILE = kThreads * kUnroll * E;
    constexpr uint32_t FULLMASK = (1u << E) - 1u;
    __shared__ unsigned int s_tot[kUnroll * kWarpsPerCta];

    const int64_t tile = blockIdx.x;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const T* __restrict__ pi = (const T*)descs[c].in;
    T* __restrict__ po = (T*)descs[c].out;
    const uint32_t* __restrict__ vi = descs[c].vin;
    uint32_t* __restrict__ vo = descs[c].vout;
    const uint32_t* __restrict__ mval = descs[c].mval;
    const uint32_t* __restrict__ mvalid = descs[c].mvalid;
    const int64_t len = descs[c].len;
    const int64_t off = descs[c].off, moff = descs[c].moff, mvoff = descs[c].mvoff;
    const int64_t base = (tile - descs[c].tile0) * TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool full = base + TILE <= len;
    const long long out_pos = tile_offsets[tile];  // first output slot of this tile

    Vec<T, E> x[kUnroll];
    uint32_t sel[kUnroll], val[kUnroll];
    if (full) {
#pragma unroll
        for (int j = 0; j < kUnroll; j++) x[j].load(pi + base + (int64_t)(j * kThreads + threadIdx.x) * E);
        MaskRaw<E, kUnroll> rm, rmv, rv;
        const int64_t e_first = base + (int64_t)threadIdx.x * E;
        mask_issue_general<E, kUnroll>(rm, mval, moff + e_first, (int64_t)kThreads * E);
        if (mvalid) mask_issue_general<E, kUnroll>(rmv, mvalid, mvoff + e_first, (int64_t)kThreads * E);
        if (vi) mask_issue_general<E, kUnroll>(rv, vi, off + e_first, (int64_t)kThreads * E);
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            sel[j] = mask_get<E, kUnroll>(rm, j);
            if (mvalid) sel[j] &= mask_get<E, kUnroll>(rmv, j);
            val[j] = vi ? mask_get<E, kUnroll>(rv, j) : FULLMASK;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const int64_t e0 = base + (int64_t)(j * kThreads + threadIdx.x) * E;
            const uint32_t in_range = tail_mask<E>(e0, len);
            sel[j] = 0; val[j] = FULLMASK;
            if (in_range) {
                sel[j] = load_bits<E>(mval, moff + e0) & in_range;
                if (mvalid) sel[j] &= load_bits<E>(mvalid, mvoff + e0);
                if (vi) val[j] = load_bits<E>(vi, off + e0);
#pragma unroll
                for (int e = 0; e < E; e++) if ((in_range >> e) & 1u) x[j].e[e] = pi[e0 + e];
            }
        }
    }

    // Output position of every slot: slots are ordered (step j, thread, element).  Warp-inclusive scans of the
    // per-thread counts of all steps (registers only), barrier, warp 0 turns the kUnroll x 8 per-(step, warp) totals
    // into exclusive bases with one more warp scan, barrier; positions inside the tile are 32-bit.
    unsigned int cnt[kUnroll], incl[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        cnt[j] = __popc(sel[j]);
        incl[j] = cnt[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, incl[j], o); if (lane >= o) incl[j] += u; }
        if (lane == 31) s_tot[j * kWarpsPerCta + warp] = incl[j];
    }
    __syncthreads();
    static_assert(kUnroll * kWarpsPerCta == 32, "one warp scans the per-(step, warp) totals");
    if (warp == 0) {
        const unsigned int v = s_tot[lane];
        unsigned int in2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, in2, o); if (lane >= o) in2 += u; }
        s_tot[lane] = in2 - v;  // exclusive base of (step, warp) inside the tile
    }
    __syncthreads();
    T* __restrict__ pt = po + out_pos;                                   // this tile's slice of the output values
    const int64_t vbit0 = out_pos;                                      // ... and of the output validity bits
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const unsigned int my_pos = s_tot[j * kWarpsPerCta + warp] + incl[j] - cnt[j];
        unsigned int p = my_pos;
#pragma unroll
        for (int e = 0; e < E; e++)
            if ((sel[j] >> e) & 1u) pt[p++] = x[j].e[e];
        // validity: compact this lane's E bits, OR the 32/E lanes of the word together, two atomics at most
        if (vo) {
            uint32_t cbits = 0;
            int k = 0;
#pragma unroll
            for (int e = 0; e < E; e++)
                if ((sel[j] >> e) & 1u) { cbits |= ((val[j] >> e) & 1u) << k; k++; }
            const unsigned int word_pos = __shfl_sync(0xffffffffu, my_pos, lane & ~(G - 1));  // leader's position
            uint32_t wbits = cbits << (my_pos - word_pos);                                    // < 32 selected per word
            unsigned int wcnt = cnt[j];
#pragma unroll
            for (int o = 1; o < G; o <<= 1) { wbits |= __shfl_xor_sync(0xffffffffu, wbits, o); wcnt += __shfl_xor_sync(0xffffffffu, wcnt, o); }
            if ((lane & (G - 1)) == 0 && wbits) {
                const int64_t bit = vbit0 + word_pos;
                const int sh = (int)(bit & 31);
                atomicOr(&vo[bit >> 5], wbits << sh);
                if (sh && sh + (int)wcnt > 32) atomicOr(&vo[(bit >> 5) + 1], wbits >> (32 - sh));
            }
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------------------------
template <int OP>
static cudaError_t cmp_one(const BinDesc* d, int n, int64_t tiles, int ta, int tb, bool scalar_rhs, double scalar, uint32_t* wc, cudaStream_t s) {
    if (scalar_rhs) k_compare<OP, true><<<(unsigned)tiles, kThreads, 0, s>>>(d, n, ta, tb, scalar, wc);
    else k_compare<OP, false><<<(unsigned)tiles, kThreads, 0, s>>>(d, n, ta, tb, scalar, wc);
    return cudaGetLastError();
}

// ---- compare an integer column of at most 32 bits with a SCALAR: no Float64 arithmetic at all --------------------------------
// `cast(x, Float64) OP s` with x an integer is a statement about integers: x > s <=> x >= floor(s) + 1, x >= s <=> x >= ceil(s),
// x < s <=> x <= ceil(s) - 1, x <= s <=> x <= floor(s), x == s <=> s is integral and x == s (a NaN scalar compares false, != true).
// The host turns (op, s) into a closed range [lo, hi] of T (possibly empty) and a negate flag; the kernel is then a range test on
// 8 elements per lane (one or two 16-byte loads), against 2 elements per lane, an I2F.F64 and a DSETP per element in k_compare --
// the generic kernel ran an Int32 column at 0.33 of the roofline (conversion-issue bound).  Null slots compare as 0, the cast's payload.
// Measured and dropped: issuing the loads of all kCmpIntTiles tiles before the first compare (72-80 registers, 3 CTAs per SM instead of 8):
// 0.123 -> 0.172 ms for 1e8 Int32 rows with nulls -- residency beats per-thread load depth here.
constexpr int kCmpIntTiles = 4;
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_compare_int(const BinDesc* __restrict__ descs, int n_chunks, long long total_tiles, long long lo64, long long hi64, int negate,
              uint32_t* __restrict__ warp_counts) {
    constexpr int E = 8;
    constexpr int TILE = kThreads * E;   // == compare_tile_elems(): the host's tile numbering and count layout stay as they are
    using C = typename std::conditional<((T)-1 < (T)0), int, unsigned int>::type;
    const C lo = (C)lo64, hi = (C)hi64;
#pragma unroll 1
    for (int kk = 0; kk < kCmpIntTiles; kk++) {   // a tile is only 2-8 KiB of input here: several per CTA
    const int64_t tile = (int64_t)blockIdx.x * kCmpIntTiles + kk;
    if (tile >= total_tiles) break;
    const int c = (n_chunks == 1) ? 0 : find_chunk(descs, n_chunks, tile);
    const T* __restrict__ pa = (const T*)descs[c].a;
    uint32_t* __restrict__ po = (uint32_t*)descs[c].out;
    const uint32_t* __restrict__ va = descs[c].va;
    uint32_t* __restrict__ vo = descs[c].vout;
    const int64_t len = descs[c].len, offa = descs[c].offa;
    const int64_t base = (tile - descs[c].tile0) * TILE;
    const int64_t e0 = base + (int64_t)threadIdx.x * E;
    const uint32_t in_range = tail_mask<E>(e0, len);
    uint32_t m = in_range;
    T v[E];
    if (in_range == 0xffu) {
        if constexpr (sizeof(T) == 4) {
            const uint4 q0 = ld_stream16(pa + e0), q1 = ld_stream16(pa + e0 + 4);
            v[0] = (T)q0.x; v[1] = (T)q0.y; v[2] = (T)q0.z; v[3] = (T)q0.w; v[4] = (T)q1.x; v[5] = (T)q1.y; v[6] = (T)q1.z; v[7] = (T)q1.w;
        } else if constexpr (sizeof(T) == 2) {
            const uint4 q = ld_stream16(pa + e0);
            v[0] = (T)(q.x & 0xffffu); v[1] = (T)(q.x >> 16); v[2] = (T)(q.y & 0xffffu); v[3] = (T)(q.y >> 16);
            v[4] = (T)(q.z & 0xffffu); v[5] = (T)(q.z >> 16); v[6] = (T)(q.w & 0xffffu); v[7] = (T)(q.w >> 16);
        } else {
            const uint2 q = ld_stream8(pa + e0);
#pragma unroll
            for (int e = 0; e < 4; e++) { v[e] = (T)((q.x >> (8 * e)) & 0xffu); v[4 + e] = (T)((q.y >> (8 * e)) & 0xffu); }
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = ((in_range >> e) & 1u) ? pa[e0 + e] : (T)0;
    }
    if (va && in_range) m &= load_bits<E>(va, offa + e0);
    uint32_t bits = 0;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const C x = (C)(((m >> e) & 1u) || !va ? v[e] : (T)0);   // a null slot compares as the cast's payload, 0
        const bool in = (x >= lo) & (x <= hi);
        bits |= ((in ? 1u : 0u) ^ (unsigned)negate) << e;
    }
    bits &= in_range;
    store_bits<E>(po, e0, bits, in_range != 0);
    unsigned int nvalid = 0;
    if (vo) {
        store_bits<E>(vo, e0, m, in_range != 0);
        nvalid = __popc(m);
        const unsigned int wvalid = __reduce_add_sync(0xffffffffu, nvalid);
        if ((threadIdx.x & 31) == 0) warp_counts[tile * kWarpsPerCta + (threadIdx.x >> 5)] = wvalid;
    }
    }
}

// (op, scalar) -> [lo, hi] within T's range (lo > hi: empty) + negate.  Returns false when the column type has no integer fast path.
static bool compare_int_plan(int op, int ta, double s, long long* lo, long long* hi, int* negate) {
    double tmin, tmax;
    switch (ta) {
        case T_I8: tmin = -128.0; tmax = 127.0; break;
        case T_I16: tmin = -32768.0; tmax = 32767.0; break;
        case T_I32: tmin = -2147483648.0; tmax = 2147483647.0; break;
        case T_U8: tmin = 0.0; tmax = 255.0; break;
        case T_U16: tmin = 0.0; tmax = 65535.0; break;
        case T_U32: tmin = 0.0; tmax = 4294967295.0; break;
        default: return false;
    }
    *negate = 0;
    double lod = tmin, hid = tmax;
    bool empty = false;
    if (s != s) { empty = true; *negate = op == CMP_NE; }
    else switch (op) {
        case CMP_GT: lod = floor(s) + 1.0; break;
        case CMP_GE: lod = ceil(s); break;
        case CMP_LT: hid = ceil(s) - 1.0; break;
        case CMP_LE: hid = floor(s); break;
        default:   // EQ / NE
            if (floor(s) == s && s >= tmin && s <= tmax) { lod = s; hid = s; } else empty = true;
            *negate = op == CMP_NE;
            break;
    }
    if (lod > tmax || hid < tmin || lod > hid) empty = true;
    if (empty) { *lo = 1; *hi = 0; return true; }
    *lo = (long long)(lod < tmin ? tmin : lod);
    *hi = (long long)(hid > tmax ? tmax : hid);
    return true;
}

int compare_tile_elems() { return kThreads * kUnroll * 2; }
static_assert(kThreads * kUnroll * 2 == kThreads * 8, "k_compare_int walks the host's compare tiles");

cudaError_t launch_compare(int op, const BinDesc* d, int n, int64_t tiles, int ta, int tb, bool scalar_rhs, double scalar, uint32_t* wc,
                           cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    long long lo = 0, hi = 0;
    int negate = 0;
    if (scalar_rhs && op >= CMP_GT && op <= CMP_LE && compare_int_plan(op, ta, scalar, &lo, &hi, &negate)) {
        const unsigned g = (unsigned)((tiles + kCmpIntTiles - 1) / kCmpIntTiles);
        switch (ta) {
            case T_I8: k_compare_int<int8_t><<<g, kThreads, 0, s>>>(d, n, tiles, lo, hi, negate, wc); break;
            case T_I16: k_compare_int<int16_t><<<g, kThreads, 0, s>>>(d, n, tiles, lo, hi, negate, wc); break;
            case T_I32: k_compare_int<int32_t><<<g, kThreads, 0, s>>>(d, n, tiles, lo, hi, negate, wc); break;
            case T_U8: k_compare_int<uint8_t><<<g, kThreads, 0, s>>>(d, n, tiles, lo, hi, negate, wc); break;
            case T_U16: k_compare_int<uint16_t><<<g, kThreads, 0, s>>>(d, n, tiles, lo, hi, negate, wc); break;
            default: k_compare_int<uint32_t><<<g, kThreads, 0, s>>>(d, n, tiles, lo, hi, negate, wc); break;
        }
        return cudaGetLastError();
    }
    switch (op) {
        case CMP_GT: return cmp_one<CMP_GT>(d, n, tiles, ta, tb, scalar_rhs, scalar, wc, s);
        case CMP_GE: return cmp_one<CMP_GE>(d, n, tiles, ta, tb, scalar_rhs, scalar, wc, s);
        case CMP_EQ: return cmp_one<CMP_EQ>(d, n, tiles, ta, tb, scalar_rhs, scalar, wc, s);
        case CMP_NE: return cmp_one<CMP_NE>(d, n, tiles, ta, tb, scalar_rhs, scalar, wc, s);
        case CMP_LT: return cmp_one<CMP_LT>(d, n, tiles, ta, tb, scalar_rhs, scalar, wc, s);
        case CMP_LE: return cmp_one<CMP_LE>(d, n, tiles, ta, tb, scalar_rhs, scalar, wc, s);
        default: return cudaErrorInvalidValue;
    }
}

int bool_tile_elems() { return kBoolTile; }

cudaError_t launch_boolean(int op, const void* d, int n, int64_t tiles, uint32_t* wc, cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const BoolDesc* bd = (const BoolDesc*)d;
    switch (op) {
        case BOOL_AND: k_boolean<BOOL_AND><<<(unsigned)tiles, kThreads, 0, s>>>(bd, n, wc); break;
        case BOOL_OR: k_boolean<BOOL_OR><<<(unsigned)tiles, kThreads, 0, s>>>(bd, n, wc); break;
        case BOOL_NOT: k_boolean<BOOL_NOT><<<(unsigned)tiles, kThreads, 0, s>>>(bd, n, wc); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_filter_count(const void* d, int n, int64_t tiles, int tile_elems, unsigned int* tile_counts, long long* tile_offsets,
                                long long* chunk_totals, cudaStream_t s) {
    const FilterDesc* fd = (const FilterDesc*)d;
    if (tiles > 0) {
        if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        k_filter_count<<<(unsigned)((tiles + kWarpsPerCta - 1) / kWarpsPerCta), kThreads, 0, s>>>(fd, n, tiles, tile_elems, tile_counts);
    }
    if (n > 0) k_filter_scan<<<(unsigned)n, kThreads, 0, s>>>(fd, tile_elems, tile_counts, tile_offsets, chunk_totals);
    return cudaGetLastError();
}

cudaError_t launch_filter_scatter(int dtype, const void* d, int n, int64_t tiles, const long long* tile_offsets, cudaStream_t s) {
    if (tiles <= 0) return cudaSuccess;
    const FilterDesc* fd = (const FilterDesc*)d;
    switch (dtype_width(dtype)) {  // a move of opaque w-byte values: one instantiation per width
        case 1: k_filter_scatter<uint8_t><<<(unsigned)tiles, kThreads, 0, s>>>(fd, n, tile_offsets); break;
        case 2: k_filter_scatter<uint16_t><<<(unsigned)tiles, kThreads, 0, s>>>(fd, n, tile_offsets); break;
        case 4: k_filter_scatter<uint32_t><<<(unsigned)tiles, kThreads, 0, s>>>(fd, n, tile_offsets); break;
        default: k_filter_scatter<uint64_t><<<(unsigned)tiles, kThreads, 0, s>>>(fd, n, tile_offsets); break;
    }
    return cudaGetLastError();
}

size_t filter_desc_size() { return sizeof(FilterDesc); }
size_t bool_desc_size() { return sizeof(BoolDesc); }

// host-side fillers (the runtime does not see the struct layouts)
void fill_filter_desc(void* base, int64_t i, const void* in, void* out, const uint32_t* vin, uint32_t* vout, const uint32_t* mval,
                      const uint32_t* mvalid, int64_t len, int64_t tile0, int off, int moff, int mvoff) {
    FilterDesc* d = (FilterDesc*)base + i;
    *d = FilterDesc{in, out, vin, vout, mval, mvalid, len, tile0, off, moff, mvoff, 0};
}
void fill_bool_desc(void* base, int64_t i, const uint32_t* a, const uint32_t* b, uint32_t* out, const uint32_t* va, const uint32_t* vb,
                    uint32_t* vout, int64_t len, int64_t tile0, int offa, int offb, int voffa, int voffb) {
    BoolDesc* d = (BoolDesc*)base + i;
    *d = BoolDesc{a, b, out, va, vb, vout, len, tile0, offa, offb, voffa, voffb};
}

}  // namespace bdf
