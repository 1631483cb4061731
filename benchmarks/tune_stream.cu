// tune_stream.cu -- developer micro-benchmark used to pick the tile shape / load width / grid policy of
// the streaming kernels (K1 add, K4 sum, K5 fused add+sum) on a B200.  Not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo benchmarks/tune_stream.cu -o gpurun_out/tune_stream
//   gpurun -- ./gpurun_out/tune_stream [rows]
// Every variant is timed with CUDA events over inputs far larger than L2 (1e8 f64 rows), best and median of 20.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct V16 { union { uint4 q; double d[2]; }; };
struct V32 { union { uint64_t u[4]; double d[4]; }; };

__device__ __forceinline__ void ld16(V16& v, const void* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.q.x), "=r"(v.q.y), "=r"(v.q.z), "=r"(v.q.w) : "l"(p));
}
__device__ __forceinline__ void st16(void* p, const V16& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.q.x), "r"(v.q.y), "r"(v.q.z), "r"(v.q.w) : "memory");
}
__device__ __forceinline__ void ld32(V32& v, const void* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.u[0]), "=l"(v.u[1]), "=l"(v.u[2]), "=l"(v.u[3]) : "l"(p));
}
__device__ __forceinline__ void st32(void* p, const V32& v) {
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(v.u[0]), "l"(v.u[1]), "l"(v.u[2]), "l"(v.u[3]) : "memory");
}

template <int VB> struct VecSel;
template <> struct VecSel<16> { using T = V16; static constexpr int N = 2; };
template <> struct VecSel<32> { using T = V32; static constexpr int N = 4; };
template <int VB> __device__ __forceinline__ void ldv(typename VecSel<VB>::T& v, const void* p) { if constexpr (VB == 16) ld16(v, p); else ld32(v, p); }
template <int VB> __device__ __forceinline__ void stv(void* p, const typename VecSel<VB>::T& v) { if constexpr (VB == 16) st16(p, v); else st32(p, v); }

// ---- add: c = a + b ----------------------------------------------------------------------------------
template <int THREADS, int U, int VB, bool PERSIST>
__global__ void __launch_bounds__(THREADS) k_add(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c, long n_tiles) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    for (long tile = blockIdx.x; tile < n_tiles; tile += PERSIST ? gridDim.x : n_tiles) {
        const long base = tile * TILE;
        V x[U], y[U];
#pragma unroll
        for (int j = 0; j < U; j++) { const long e = base + (long)(j * THREADS + threadIdx.x) * N; ldv<VB>(x[j], a + e); ldv<VB>(y[j], b + e); }
#pragma unroll
        for (int j = 0; j < U; j++) {
            const long e = base + (long)(j * THREADS + threadIdx.x) * N;
            V r;
#pragma unroll
            for (int k = 0; k < N; k++) r.d[k] = __dadd_rn(x[j].d[k], y[j].d[k]);
            stv<VB>(c + e, r);
        }
    }
}

// ---- sum (read only) and fused add+sum -----------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double r = 0;
    if (warp == 0) {
        r = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
}

template <int THREADS, int U, int VB>
__global__ void __launch_bounds__(THREADS) k_sum(const double* __restrict__ a, long n_tiles, double* __restrict__ partials) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    __shared__ double smem[32];
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0.0;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long base = tile * TILE;
        V x[U];
#pragma unroll
        for (int j = 0; j < U; j++) ldv<VB>(x[j], a + base + (long)(j * THREADS + threadIdx.x) * N);
#pragma unroll
        for (int j = 0; j < U; j++)
#pragma unroll
            for (int k = 0; k < N; k++) acc[k] = __dadd_rn(acc[k], x[j].d[k]);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += acc[k];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

template <int THREADS, int U, int VB>
__global__ void __launch_bounds__(THREADS) k_add_sum(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c, long n_tiles,
                                                      double* __restrict__ partials) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    __shared__ double smem[32];
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0.0;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long base = tile * TILE;
        V x[U], y[U];
#pragma unroll
        for (int j = 0; j < U; j++) { const long e = base + (long)(j * THREADS + threadIdx.x) * N; ldv<VB>(x[j], a + e); ldv<VB>(y[j], b + e); }
#pragma unroll
        for (int j = 0; j < U; j++) {
            const long e = base + (long)(j * THREADS + threadIdx.x) * N;
            V r;
#pragma unroll
            for (int k = 0; k < N; k++) { r.d[k] = __dadd_rn(x[j].d[k], y[j].d[k]); acc[k] = __dadd_rn(acc[k], r.d[k]); }
            stv<VB>(c + e, r);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += acc[k];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// reverse-order sum: visit the most recently written tiles (still in L2 after the add) first
template <int THREADS, int U, int VB>
__global__ void __launch_bounds__(THREADS) k_sum_rev(const double* __restrict__ a, long n_tiles, double* __restrict__ partials) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    __shared__ double smem[32];
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0.0;
    for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long base = (n_tiles - 1 - t) * TILE;
        V x[U];
#pragma unroll
        for (int j = 0; j < U; j++) ldv<VB>(x[j], a + base + (long)(j * THREADS + threadIdx.x) * N);
#pragma unroll
        for (int j = 0; j < U; j++)
#pragma unroll
            for (int k = 0; k < N; k++) acc[k] = __dadd_rn(acc[k], x[j].d[k]);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += acc[k];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// fused add+sum, NOT persistent: each CTA owns K consecutive tiles and writes one partial
template <int THREADS, int U, int VB, int K>
__global__ void __launch_bounds__(THREADS) k_add_sum_np(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c, long n_tiles,
                                                         double* __restrict__ partials) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    __shared__ double smem[32];
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0.0;
#pragma unroll 1
    for (int kk = 0; kk < K; kk++) {
        const long tile = (long)blockIdx.x * K + kk;
        if (tile >= n_tiles) break;
        const long base = tile * TILE;
        V x[U], y[U];
#pragma unroll
        for (int j = 0; j < U; j++) { const long e = base + (long)(j * THREADS + threadIdx.x) * N; ldv<VB>(x[j], a + e); ldv<VB>(y[j], b + e); }
#pragma unroll
        for (int j = 0; j < U; j++) {
            const long e = base + (long)(j * THREADS + threadIdx.x) * N;
            V r;
#pragma unroll
            for (int k = 0; k < N; k++) { r.d[k] = __dadd_rn(x[j].d[k], y[j].d[k]); acc[k] = __dadd_rn(acc[k], r.d[k]); }
            stv<VB>(c + e, r);
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += acc[k];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// sum, NOT persistent: each CTA owns K consecutive tiles and writes one partial
template <int THREADS, int U, int VB, int K>
__global__ void __launch_bounds__(THREADS) k_sum_np(const double* __restrict__ a, long n_tiles, double* __restrict__ partials) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    __shared__ double smem[32];
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0.0;
#pragma unroll 1
    for (int kk = 0; kk < K; kk++) {
        const long tile = (long)blockIdx.x * K + kk;
        if (tile >= n_tiles) break;
        const long base = tile * TILE;
        V x[U];
#pragma unroll
        for (int j = 0; j < U; j++) ldv<VB>(x[j], a + base + (long)(j * THREADS + threadIdx.x) * N);
#pragma unroll
        for (int j = 0; j < U; j++)
#pragma unroll
            for (int k = 0; k < N; k++) acc[k] = __dadd_rn(acc[k], x[j].d[k]);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += acc[k];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// persistent sum with register double buffering: the loads of tile t+1 are issued before tile t is folded
template <int THREADS, int U, int VB>
__global__ void __launch_bounds__(THREADS) k_sum_pipe(const double* __restrict__ a, long n_tiles, double* __restrict__ partials) {
    using V = typename VecSel<VB>::T;
    constexpr int N = VecSel<VB>::N;
    constexpr long TILE = (long)THREADS * U * N;
    __shared__ double smem[32];
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0.0;
    V x[U], y[U];
    long tile = blockIdx.x;
    if (tile < n_tiles) {
#pragma unroll
        for (int j = 0; j < U; j++) ldv<VB>(x[j], a + tile * TILE + (long)(j * THREADS + threadIdx.x) * N);
    }
    while (tile < n_tiles) {
        const long next = tile + gridDim.x;
        if (next < n_tiles) {
#pragma unroll
            for (int j = 0; j < U; j++) ldv<VB>(y[j], a + next * TILE + (long)(j * THREADS + threadIdx.x) * N);
        }
#pragma unroll
        for (int j = 0; j < U; j++)
#pragma unroll
            for (int k = 0; k < N; k++) acc[k] = __dadd_rn(acc[k], x[j].d[k]);
#pragma unroll
        for (int j = 0; j < U; j++) x[j] = y[j];
        tile = next;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < N; k++) s += acc[k];
    s = block_sum(s, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// ---- harness -----------------------------------------------------------------------------------------
static cudaEvent_t e0, e1;
template <typename F>
static void run(const char* name, double bytes, F launch) {
    for (int i = 0; i < 3; i++) launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> ms(20);
    for (auto& m : ms) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&m, e0, e1));
    }
    CK(cudaGetLastError());
    std::sort(ms.begin(), ms.end());
    printf("%-44s best %.4f ms %7.1f GB/s | median %.4f ms %7.1f GB/s\n", name, ms[0], bytes / ms[0] / 1e6, ms[10], bytes / ms[10] / 1e6);
}

int main(int argc, char** argv) {
    const long rows = argc > 1 ? atol(argv[1]) : 100000000L;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, rows %ld\n", prop.name, sms, rows);
    double *a, *b, *c, *partials;
    const size_t bytes = (size_t)rows * 8 + (1 << 20);
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&c, bytes)); CK(cudaMalloc(&partials, 1 << 20));
    CK(cudaMemset(a, 0, bytes)); CK(cudaMemset(b, 0, bytes)); CK(cudaMemset(c, 0, bytes));
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double B_ADD = 24.0 * rows, B_SUM = 8.0 * rows;

    run("memcpy d2d (8 B/row read + 8 write)", 16.0 * rows, [&] { cudaMemcpyAsync(c, a, (size_t)rows * 8, cudaMemcpyDeviceToDevice); });

#define ADD(T, U, VB, P, CPS) { constexpr long TILE = (long)T * U * (VB / 8); const long nt = rows / TILE; \
        const long grid = P ? (long)sms * CPS : nt; char nm[96]; snprintf(nm, 96, "add  T=%d U=%d VB=%d %s", T, U, VB, P ? "persist x" #CPS : "tile/CTA"); \
        run(nm, B_ADD, [&] { k_add<T, U, VB, P><<<(unsigned)grid, T>>>(a, b, c, nt); }); }
    ADD(256, 4, 16, false, 0) ADD(256, 2, 16, false, 0) ADD(256, 8, 16, false, 0) ADD(512, 4, 16, false, 0) ADD(128, 4, 16, false, 0)
    ADD(256, 2, 32, false, 0) ADD(256, 4, 32, false, 0) ADD(512, 2, 32, false, 0) ADD(128, 4, 32, false, 0)
    ADD(256, 4, 16, true, 4) ADD(256, 4, 16, true, 6) ADD(256, 4, 16, true, 8) ADD(256, 2, 32, true, 4) ADD(256, 2, 32, true, 6) ADD(256, 2, 32, true, 8)
    ADD(512, 2, 32, true, 2) ADD(512, 2, 32, true, 3) ADD(512, 2, 32, true, 4) ADD(1024, 1, 32, true, 2) ADD(1024, 2, 16, true, 2)

#define SUM(T, U, VB, CPS) { constexpr long TILE = (long)T * U * (VB / 8); const long nt = rows / TILE; char nm[96]; \
        snprintf(nm, 96, "sum  T=%d U=%d VB=%d grid=SMs x" #CPS, T, U, VB); \
        run(nm, B_SUM, [&] { k_sum<T, U, VB><<<sms * CPS, T>>>(a, nt, partials); }); }
    SUM(256, 4, 16, 5) SUM(256, 4, 16, 8) SUM(256, 8, 16, 4) SUM(256, 8, 16, 8) SUM(256, 16, 16, 4) SUM(512, 4, 16, 4) SUM(512, 8, 16, 2) SUM(512, 8, 16, 4)
    SUM(256, 4, 32, 4) SUM(256, 4, 32, 8) SUM(256, 8, 32, 4) SUM(512, 4, 32, 2) SUM(512, 4, 32, 4) SUM(1024, 4, 32, 1) SUM(1024, 4, 32, 2) SUM(1024, 2, 32, 2)
    SUM(256, 2, 32, 8) SUM(128, 8, 32, 8) SUM(128, 4, 32, 16)

#define ADDSUM(T, U, VB, CPS) { constexpr long TILE = (long)T * U * (VB / 8); const long nt = rows / TILE; char nm[96]; \
        snprintf(nm, 96, "add+sum fused T=%d U=%d VB=%d grid=SMs x" #CPS, T, U, VB); \
        run(nm, B_ADD, [&] { k_add_sum<T, U, VB><<<sms * CPS, T>>>(a, b, c, nt, partials); }); }
    ADDSUM(256, 4, 16, 4) ADDSUM(256, 4, 16, 6) ADDSUM(256, 4, 16, 8) ADDSUM(256, 2, 32, 4) ADDSUM(256, 2, 32, 6) ADDSUM(256, 2, 32, 8)
    ADDSUM(512, 2, 32, 2) ADDSUM(512, 2, 32, 4) ADDSUM(256, 4, 32, 4) ADDSUM(512, 4, 16, 4)
#define ADDSUMNP(T, U, VB, K) { constexpr long TILE = (long)T * U * (VB / 8); const long nt = rows / TILE; char nm[96]; \
        snprintf(nm, 96, "add+sum fused non-persist T=%d U=%d VB=%d K=%d", T, U, VB, K); \
        run(nm, B_ADD, [&] { k_add_sum_np<T, U, VB, K><<<(unsigned)((nt + K - 1) / K), T>>>(a, b, c, nt, partials); }); }
    ADDSUMNP(256, 4, 16, 1) ADDSUMNP(256, 4, 16, 2) ADDSUMNP(256, 4, 16, 4) ADDSUMNP(256, 4, 16, 8) ADDSUMNP(512, 4, 16, 1) ADDSUMNP(512, 4, 16, 2)
    ADDSUMNP(256, 2, 32, 1) ADDSUMNP(256, 2, 32, 4) ADDSUMNP(512, 2, 32, 1) ADDSUMNP(512, 2, 32, 2) ADDSUMNP(512, 2, 32, 4)

#define SUMNP(T, U, VB, K) { constexpr long TILE = (long)T * U * (VB / 8); const long nt = rows / TILE; char nm[96]; \
        snprintf(nm, 96, "sum non-persist T=%d U=%d VB=%d K=%d", T, U, VB, K); \
        run(nm, B_SUM, [&] { k_sum_np<T, U, VB, K><<<(unsigned)((nt + K - 1) / K), T>>>(a, nt, partials); }); }
    SUMNP(256, 4, 16, 1) SUMNP(256, 4, 16, 2) SUMNP(256, 4, 16, 4) SUMNP(256, 4, 16, 8) SUMNP(256, 8, 16, 1) SUMNP(256, 8, 16, 2) SUMNP(256, 8, 16, 4)
    SUMNP(512, 4, 16, 1) SUMNP(512, 4, 16, 2) SUMNP(512, 8, 16, 1) SUMNP(256, 4, 32, 1) SUMNP(256, 4, 32, 2) SUMNP(256, 4, 32, 4) SUMNP(512, 4, 32, 1) SUMNP(512, 4, 32, 2)
#define SUMPIPE(T, U, VB, CPS) { constexpr long TILE = (long)T * U * (VB / 8); const long nt = rows / TILE; char nm[96]; \
        snprintf(nm, 96, "sum pipelined T=%d U=%d VB=%d grid=SMs x" #CPS, T, U, VB); \
        run(nm, B_SUM, [&] { k_sum_pipe<T, U, VB><<<sms * CPS, T>>>(a, nt, partials); }); }
    SUMPIPE(256, 4, 16, 4) SUMPIPE(256, 4, 16, 5) SUMPIPE(256, 4, 16, 6) SUMPIPE(256, 2, 32, 4) SUMPIPE(512, 4, 16, 2) SUMPIPE(256, 8, 16, 2) SUMPIPE(256, 8, 16, 4)

    // sequence as the product runs it: add, then sum over c (events around the sum only)
    {
        constexpr long TILE = 256L * 4 * 2; const long nt = rows / TILE;
        auto seq = [&](const char* name, bool rev) {
            std::vector<float> ms(20);
            for (auto& m : ms) {
                k_add<256, 4, 16, false><<<(unsigned)nt, 256>>>(a, b, c, nt);
                CK(cudaEventRecord(e0));
                if (rev) k_sum_rev<256, 4, 16><<<sms * 5, 256>>>(c, nt, partials); else k_sum<256, 4, 16><<<sms * 5, 256>>>(c, nt, partials);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&m, e0, e1));
            }
            std::sort(ms.begin(), ms.end());
            printf("%-44s best %.4f ms %7.1f GB/s | median %.4f ms %7.1f GB/s\n", name, ms[0], B_SUM / ms[0] / 1e6, ms[10], B_SUM / ms[10] / 1e6);
        };
        seq("sum right after add (forward)", false);
        seq("sum right after add (reverse tile order)", true);
    }
    return 0;
}
