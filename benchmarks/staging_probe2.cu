// staging_probe2.cu -- how fast can PAGEABLE host memory (what an Arrow MutableBuffer is) be moved to / from the GPU?
// Explores the design space of the library's staged copies (csrc/runtime.cu "pageable host memory"): worker count,
// job size, slots per worker, non-temporal stores into the pinned slot, fresh (never touched) destination pages with and
// without MADV_POPULATE_WRITE / MADV_HUGEPAGE, and both directions at once (what c = a + b does: 2 columns up, 1 down).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o benchmarks/bin/staging_probe2 benchmarks/staging_probe2.cu -lpthread
//   benchmarks/bin/staging_probe2 [MB per column, default 800]
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void copy_nt(char* dst, const char* src, size_t n) {   // 64-byte aligned dst; streaming stores (no read-for-ownership)
#if defined(__AVX2__)
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence();
#else
    memcpy(dst, src, n);
#endif
}

struct Slot { char* p; cudaEvent_t ev; bool busy; };

struct Engine {
    int threads, slots_per_thread;
    size_t job;
    bool nt;
    std::vector<std::vector<Slot>> slots;
    std::vector<cudaStream_t> streams;
    Engine(int t, int spt, size_t j, bool nt_, int n_streams) : threads(t), slots_per_thread(spt), job(j), nt(nt_) {
        slots.resize(t);
        for (auto& v : slots)
            for (int i = 0; i < spt; i++) {
                Slot s{nullptr, nullptr, false};
                CK(cudaHostAlloc((void**)&s.p, job, cudaHostAllocDefault));
                CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
                memset(s.p, 1, job);
                v.push_back(s);
            }
        streams.resize(n_streams);
        for (auto& s : streams) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    }
    ~Engine() {
        for (auto& v : slots) for (auto& s : v) { cudaFreeHost(s.p); cudaEventDestroy(s.ev); }
        for (auto& s : streams) cudaStreamDestroy(s);
    }
    // host (pageable) -> device
    void up(char* dev, const char* host, size_t bytes) {
        std::atomic<size_t> next{0};
        const size_t n_jobs = (bytes + job - 1) / job;
        auto work = [&](int w) {
            int k = 0;
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= n_jobs) break;
                Slot& s = slots[w][k++ % slots_per_thread];
                if (s.busy) CK(cudaEventSynchronize(s.ev));
                const size_t off = j * job, n = std::min(job, bytes - off);
                if (nt) copy_nt(s.p, host + off, n); else memcpy(s.p, host + off, n);
                cudaStream_t st = streams[w % streams.size()];
                CK(cudaMemcpyAsync(dev + off, s.p, n, cudaMemcpyHostToDevice, st));
                CK(cudaEventRecord(s.ev, st));
                s.busy = true;
            }
        };
        std::vector<std::thread> th;
        for (int w = 1; w < threads; w++) th.emplace_back(work, w);
        work(0);
        for (auto& t : th) t.join();
        for (auto& s : streams) CK(cudaStreamSynchronize(s));
        for (auto& v : slots) for (auto& s : v) s.busy = false;
    }
    // device -> host (pageable); populate: 0 none, 1 MADV_POPULATE_WRITE per job before the copy
    void down(char* host, const char* dev, size_t bytes, int populate) {
        std::atomic<size_t> next{0};
        const size_t n_jobs = (bytes + job - 1) / job;
        auto work = [&](int w) {
            struct Pend { Slot* s; size_t off, n; };
            std::vector<Pend> pend;
            size_t head = 0;
            int k = 0;
            auto retire = [&]() {
                Pend p = pend[head++];
                CK(cudaEventSynchronize(p.s->ev));
                memcpy(host + p.off, p.s->p, p.n);
                p.s->busy = false;
            };
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= n_jobs) break;
                if ((int)(pend.size() - head) >= slots_per_thread) retire();
                Slot& s = slots[w][k++ % slots_per_thread];
                const size_t off = j * job, n = std::min(job, bytes - off);
                cudaStream_t st = streams[w % streams.size()];
                CK(cudaMemcpyAsync(s.p, dev + off, n, cudaMemcpyDeviceToHost, st));
                CK(cudaEventRecord(s.ev, st));
                s.busy = true;
                if (populate == 1) {   // fault the destination pages in while the DMA is in flight
                    const uintptr_t a = ((uintptr_t)(host + off) + 4095) & ~(uintptr_t)4095, b = ((uintptr_t)(host + off + n)) & ~(uintptr_t)4095;
                    if (b > a) madvise((void*)a, b - a, MADV_POPULATE_WRITE);
                }
                pend.push_back(Pend{&s, off, n});
            }
            while (head < pend.size()) retire();
        };
        std::vector<std::thread> th;
        for (int w = 1; w < threads; w++) th.emplace_back(work, w);
        work(0);
        for (auto& t : th) t.join();
    }
};

static void bind_to_gpu_node() {
    char bus[32];
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, 0) != cudaSuccess) return;
    for (char* p = bus; *p; p++) *p = (char)tolower(*p);
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    if (node < 0) { printf("# numa node of the GPU unknown\n"); return; }
    path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    f = fopen(path.c_str(), "r");
    if (!f) return;
    char buf[4096];
    if (!fgets(buf, sizeof buf, f)) { fclose(f); return; }
    fclose(f);
    cpu_set_t set; CPU_ZERO(&set);
    int n = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo, hi;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) { for (int c = lo; c <= hi; c++) { CPU_SET(c, &set); n++; } }
        else if (sscanf(tok, "%d", &lo) == 1) { CPU_SET(lo, &set); n++; }
    }
    if (n && sched_setaffinity(0, sizeof set, &set) == 0) printf("# bound to NUMA node %d (%d cpus)\n", node, n);
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 800) * 1000 * 1000;
    CK(cudaSetDevice(0));
    bind_to_gpu_node();
    printf("# %zu MB per column, %u hardware threads\n", bytes / 1000000, std::thread::hardware_concurrency());
    char *dev, *dev2, *dev3;
    CK(cudaMalloc((void**)&dev, bytes)); CK(cudaMalloc((void**)&dev2, bytes)); CK(cudaMalloc((void**)&dev3, bytes));
    CK(cudaMemset(dev, 3, bytes)); CK(cudaMemset(dev3, 5, bytes));
    char* src = (char*)aligned_alloc(4096, bytes);
    char* src2 = (char*)aligned_alloc(4096, bytes);
    char* dst = (char*)aligned_alloc(4096, bytes);
    memset(src, 7, bytes); memset(src2, 9, bytes); memset(dst, 0, bytes);
    char* pin;
    CK(cudaHostAlloc((void**)&pin, bytes, cudaHostAllocDefault));
    memset(pin, 2, bytes);
    cudaStream_t s0;
    CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    auto rate = [&](double dt) { return bytes / dt / 1e9; };
    // baselines
    for (int rep = 0; rep < 2; rep++) {
        double t0 = now(); CK(cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, s0)); CK(cudaStreamSynchronize(s0)); double t1 = now();
        CK(cudaMemcpyAsync(pin, dev, bytes, cudaMemcpyDeviceToHost, s0)); CK(cudaStreamSynchronize(s0)); double t2 = now();
        CK(cudaMemcpy(dev, src, bytes, cudaMemcpyHostToDevice)); double t3 = now();
        CK(cudaMemcpy(dst, dev, bytes, cudaMemcpyDeviceToHost)); double t4 = now();
        if (rep) printf("baseline: pinned H2D %.1f GB/s, pinned D2H %.1f GB/s, cudaMemcpy pageable H2D %.1f GB/s, pageable D2H %.1f GB/s\n",
                        rate(t1 - t0), rate(t2 - t1), rate(t3 - t2), rate(t4 - t3));
    }
    { double t0 = now(); memcpy(dst, src, bytes); double t1 = now(); printf("single-thread memcpy %.1f GB/s\n", rate(t1 - t0)); }
    CK(cudaMemset(dev, 3, bytes));   // the baselines overwrote it; the download checks below expect 3
    struct Cfg { int threads, spt; size_t job; bool nt; int streams; };
    std::vector<Cfg> cfgs;
    for (int t : {4, 8, 12, 16, 24, 32})
        for (size_t j : {(size_t)1 << 20, (size_t)2 << 20, (size_t)4 << 20})
            for (bool nt : {false, true}) cfgs.push_back(Cfg{t, 3, j, nt, 2});
    cfgs.push_back(Cfg{16, 3, (size_t)2 << 20, true, 1});
    cfgs.push_back(Cfg{16, 3, (size_t)2 << 20, true, 4});
    cfgs.push_back(Cfg{16, 6, (size_t)1 << 20, true, 2});
    for (const Cfg& c : cfgs) {
        Engine e(c.threads, c.spt, c.job, c.nt, c.streams);
        double up_best = 0, down_best = 0, fresh_best = 0, fresh_pop_best = 0, fresh_huge_best = 0;
        for (int rep = 0; rep < 3; rep++) {
            double t0 = now(); e.up(dev2, src, bytes); double t1 = now();
            up_best = std::max(up_best, rate(t1 - t0));
            t0 = now(); e.down(dst, dev, bytes, 0); t1 = now();
            down_best = std::max(down_best, rate(t1 - t0));
        }
        for (int variant = 0; variant < 3; variant++)
            for (int rep = 0; rep < 2; rep++) {   // a destination nobody has touched yet (MutableBuffer::new)
                char* fresh = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                if (fresh == MAP_FAILED) { perror("mmap"); return 1; }
                if (variant == 2) madvise(fresh, bytes, MADV_HUGEPAGE);
                double t0 = now(); e.down(fresh, dev, bytes, variant == 1 ? 1 : 0); double t1 = now();
                double& best = variant == 0 ? fresh_best : variant == 1 ? fresh_pop_best : fresh_huge_best;
                best = std::max(best, rate(t1 - t0));
                if (fresh[bytes / 2] != 3) { printf("BAD DATA\n"); return 1; }
                munmap(fresh, bytes);
            }
        printf("threads %2d job %zu MiB slots/thread %d nt %d streams %d: up %.1f GB/s, down (touched dst) %.1f, down (fresh dst) %.1f, fresh+populate %.1f, fresh+hugepage %.1f\n",
               c.threads, c.job >> 20, c.spt, (int)c.nt, c.streams, up_best, down_best, fresh_best, fresh_pop_best, fresh_huge_best);
        fflush(stdout);
    }
    // both directions at once, as c = a + b needs: two columns up, one down (fresh destination), each with its own engine
    for (int t : {8, 12, 16}) {
        Engine ea(t, 3, (size_t)2 << 20, true, 2), eb(t, 3, (size_t)2 << 20, true, 2);
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            char* fresh = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            double t0 = now();
            std::thread down([&] { eb.down(fresh, dev3, bytes, 0); });
            ea.up(dev, src, bytes); ea.up(dev2, src2, bytes);
            down.join();
            double t1 = now();
            best = std::min(best, t1 - t0);
            munmap(fresh, bytes);
        }
        printf("duplex: 2 columns up + 1 down (fresh), %d+%d threads: %.1f ms -> %.1f GB/s over PCIe\n", t, t, best * 1e3, 3.0 * bytes / best / 1e9);
    }
    return 0;
}
