// gather_probe.cu -- how many bytes does a random 8-byte gather cost on this GPU, per load flavour?
// out[i] = vals[idx[i]] over n = 2^27 rows (1 GiB of f64), idx a pseudo-random permutation.  The time of each variant x 1 /
// (measured copy bandwidth) bounds the DRAM bytes it moved; ncu reads 120 B per gathered row for the plain load of k_take.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kBits = 27;
constexpr uint32_t kMask = (1u << kBits) - 1u;

__device__ __forceinline__ uint32_t perm(uint32_t x) {   // bijection on [0, 2^27)
    x = (x * 0x9E3779B1u) & kMask; x ^= x >> 13; x = (x * 0x85EBCA6Bu) & kMask; x ^= x >> 11; x = (x * 0xC2B2AE35u) & kMask; x ^= x >> 15;
    return x;
}
__global__ void k_fill(double* v, uint32_t* idx, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { v[i] = (double)i; idx[i] = perm((uint32_t)i); }
}

template <int V> __device__ __forceinline__ double load(const double* p) {
    double r;
    if constexpr (V == 0) r = *p;
    else if constexpr (V == 1) r = __ldg(p);
    else if constexpr (V == 2) r = __ldcs(p);
    else if constexpr (V == 3) r = __ldcv(p);
    else if constexpr (V == 4) asm volatile("ld.global.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if constexpr (V == 5) asm volatile("ld.global.L2::128B.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if constexpr (V == 6) asm volatile("ld.global.L2::256B.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if constexpr (V == 7) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if constexpr (V == 8) asm volatile("ld.global.L1::evict_first.L2::64B.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if constexpr (V == 9) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    } else if constexpr (V == 10) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    }
    return r;
}

template <int V, int R>
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ vals, const uint32_t* __restrict__ idx, double* __restrict__ out, long long n) {
    const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * R;
    if (base + R > n) return;
    uint32_t r[R];
    if constexpr (R == 4) { const uint4 q = *reinterpret_cast<const uint4*>(idx + base); r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }
    else if constexpr (R == 2) { const uint2 q = *reinterpret_cast<const uint2*>(idx + base); r[0] = q.x; r[1] = q.y; }
    else r[0] = idx[base];
    double v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = load<V>(vals + r[k]);
    if constexpr (R == 4) { *reinterpret_cast<double2*>(out + base) = make_double2(v[0], v[1]); *reinterpret_cast<double2*>(out + base + 2) = make_double2(v[2], v[3]); }
    else if constexpr (R == 2) *reinterpret_cast<double2*>(out + base) = make_double2(v[0], v[1]);
    else out[base] = v[0];
}

template <int V, int R>
static int run(const char* name, const double* v, const uint32_t* idx, double* out, long long n, int limit) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const unsigned grid = (unsigned)(n / R / 256);
    k_gather<V, R><<<grid, 256>>>(v, idx, out, n);
    CK(cudaDeviceSynchronize());
    float best = 1e9f, sum = 0;
    for (int it = 0; it < 5; it++) {
        CK(cudaEventRecord(a));
        k_gather<V, R><<<grid, 256>>>(v, idx, out, n);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best; sum += ms;
    }
    // spot check
    double h[4]; uint32_t hi[4];
    CK(cudaMemcpy(h, out + 12345 * 4, sizeof h, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hi, idx + 12345 * 4, sizeof hi, cudaMemcpyDeviceToHost));
    const bool ok = h[0] == (double)hi[0] && h[3] == (double)hi[3];
    printf("fetch=%-3d %-44s rows/thread %d  %.4f ms avg  %.4f ms best  %.2e rows/s  %s\n", limit, name, R, sum / 5, best, n / (best * 1e-3), ok ? "ok" : "BAD");
    CK(cudaMemset(out, 0, 64));
    return 0;
}

int main() {
    const long long n = 1ll << kBits;
    double *v, *out; uint32_t* idx;
    CK(cudaMalloc(&v, n * 8)); CK(cudaMalloc(&out, n * 8)); CK(cudaMalloc(&idx, n * 4));
    k_fill<<<(unsigned)(n / 256), 256>>>(v, idx, n);
    CK(cudaDeviceSynchronize());
    for (int limit : {0, 32}) {
        if (limit) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)limit));
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity = %zu\n", got);
        if (run<0, 1>("plain ld.global", v, idx, out, n, limit)) return 1;
        if (run<0, 2>("plain ld.global", v, idx, out, n, limit)) return 1;
        if (run<0, 4>("plain ld.global", v, idx, out, n, limit)) return 1;
        if (run<1, 4>("__ldg (ld.global.nc)", v, idx, out, n, limit)) return 1;
        if (run<2, 4>("__ldcs", v, idx, out, n, limit)) return 1;
        if (run<3, 4>("__ldcv", v, idx, out, n, limit)) return 1;
        if (run<4, 4>("ld.global.L2::64B", v, idx, out, n, limit)) return 1;
        if (run<5, 4>("ld.global.L2::128B", v, idx, out, n, limit)) return 1;
        if (run<6, 4>("ld.global.L2::256B", v, idx, out, n, limit)) return 1;
        if (run<7, 4>("ld.global.nc.L1::no_allocate.L2::64B", v, idx, out, n, limit)) return 1;
        if (run<8, 4>("ld.global.L1::evict_first.L2::64B", v, idx, out, n, limit)) return 1;
        if (run<9, 4>("L2::cache_hint evict_first", v, idx, out, n, limit)) return 1;
        if (run<10, 4>("nc.no_allocate + evict_first hint + L2::64B", v, idx, out, n, limit)) return 1;
        if (run<7, 1>("ld.global.nc.L1::no_allocate.L2::64B", v, idx, out, n, limit)) return 1;
    }
    return 0;
}
