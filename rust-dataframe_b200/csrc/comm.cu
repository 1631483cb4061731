// comm.cu -- NCCL combine of the per-GPU partial aggregates (see comm.cuh).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already carries an NCCL (torch's) shares
// that copy instead of loading a second one, and the single-GPU path has no NCCL dependency at all.  Only the
// C API of nccl.h is used; the header supplies the types.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "comm.cuh"

namespace bdf {

namespace {

struct NcclApi {
    void* handle = nullptr;
    int version = 0;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* (*GetLastError)(ncclComm_t) = nullptr;
    std::string load_error;
};

NcclApi g_api;
std::once_flag g_api_once;

template <typename F>
bool bind(void* h, const char* name, F* out, std::string* err) {
    *out = reinterpret_cast<F>(dlsym(h, name));
    if (!*out) { *err = std::string("libnccl.so.2 lacks ") + name; return false; }
    return true;
}

void load_api() {
    NcclApi& a = g_api;
    const char* override_path = getenv("BDF_NCCL_LIB");
    // a copy that is already mapped into the process wins (one NCCL per process), then the loader's search path
    void* h = override_path ? dlopen(override_path, RTLD_NOW | RTLD_LOCAL) : dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
    if (!h && !override_path) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        const char* d = dlerror();
        a.load_error = std::string("cannot load NCCL (libnccl.so.2): ") + (d ? d : "unknown error");
        return;
    }
    std::string err;
    bool ok = bind(h, "ncclGetVersion", &a.GetVersion, &err) && bind(h, "ncclGetUniqueId", &a.GetUniqueId, &err) &&
              bind(h, "ncclCommInitRank", &a.CommInitRank, &err) && bind(h, "ncclCommInitAll", &a.CommInitAll, &err) &&
              bind(h, "ncclCommDestroy", &a.CommDestroy, &err) && bind(h, "ncclAllReduce", &a.AllReduce, &err) &&
              bind(h, "ncclAllGather", &a.AllGather, &err) && bind(h, "ncclGroupStart", &a.GroupStart, &err) &&
              bind(h, "ncclGroupEnd", &a.GroupEnd, &err) && bind(h, "ncclGetErrorString", &a.GetErrorString, &err);
    if (!ok) { a.load_error = err; dlclose(h); return; }
    a.GetLastError = reinterpret_cast<const char* (*)(ncclComm_t)>(dlsym(h, "ncclGetLastError"));  // optional (2.13+)
    a.GetVersion(&a.version);
    a.handle = h;
}

NcclApi* api(std::string* err) {
    std::call_once(g_api_once, load_api);
    if (!g_api.handle) { if (err) *err = g_api.load_error; return nullptr; }
    return &g_api;
}

constexpr int kSlots = 64;                       // collectives that may be in flight on the streams at once
constexpr int kFields = 7;                       // sum, count, rows, panics, chunks (ncclSum) | min (ncclMin) | max (ncclMax)
constexpr int kSlotWords = kFields * kCommMaxCols;

struct Extra { unsigned long long rows[kCommMaxCols]; unsigned int panics[kCommMaxCols]; unsigned int chunks[kCommMaxCols]; };

// AggDev records (as k_finish writes them) -> the send record, one field per contiguous run of n words.
__global__ void k_comm_pack(const AggDev* __restrict__ local, int n, Extra ex, unsigned long long* __restrict__ send) {
    const int i = threadIdx.x;
    if (i >= n) return;
    send[0 * n + i] = local[i].sum_bits;
    send[1 * n + i] = local[i].count;
    send[2 * n + i] = ex.rows[i];
    send[3 * n + i] = ex.panics[i];
    send[4 * n + i] = ex.chunks[i];
    send[5 * n + i] = local[i].min_bits;
    send[6 * n + i] = local[i].max_bits;
}

// The combined record -> result[2i] (AggDev) and result[2i+1] ({rows, panics, chunks}); float columns (bit i of
// float_mask) fold the gathered partial sums in rank order.  `result` is device-mapped host memory: fence at
// system scope after the stores.
__global__ void k_comm_unpack(const unsigned long long* __restrict__ recv, const double* __restrict__ gathered, int n, int world,
                              unsigned long long float_mask, AggDev* __restrict__ result) {
    const int i = threadIdx.x;
    if (i < n) {
        AggDev a;
        if ((float_mask >> i) & 1ull) {
            double s = 0.0;
            for (int r = 0; r < world; r++) s = __dadd_rn(s, gathered[(size_t)r * n + i]);
            a.sum_bits = (unsigned long long)__double_as_longlong(s);
            a.min_bits = ~0ull; a.max_bits = 0ull;
        } else {
            a.sum_bits = recv[0 * n + i];
            a.min_bits = recv[5 * n + i];
            a.max_bits = recv[6 * n + i];
        }
        a.count = recv[1 * n + i];
        AggDev b;
        b.sum_bits = recv[2 * n + i]; b.min_bits = recv[3 * n + i]; b.max_bits = recv[4 * n + i]; b.count = 0;
        result[2 * i] = a;
        result[2 * i + 1] = b;
    }
    __threadfence_system();
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    unsigned long long* d_send = nullptr;   // kSlots records
    unsigned long long* d_recv = nullptr;
    double* d_gather = nullptr;             // kSlots x world x kCommMaxCols
    double* d_host_stage = nullptr;         // blocking host helpers
    size_t host_stage_bytes = 0;
    int next = 0;
};

static std::string nccl_msg(NcclApi* a, ncclComm_t comm, const char* what, ncclResult_t r) {
    std::string m = std::string(what) + " failed: " + a->GetErrorString(r);
    if (a->GetLastError && comm) { const char* d = a->GetLastError(comm); if (d && *d) m += std::string(" (") + d + ")"; }
    return m;
}

#define NCCL_TRY(call, what)                                                        \
    do {                                                                            \
        ncclResult_t _r = (call);                                                   \
        if (_r != ncclSuccess) { if (err) *err = nccl_msg(a, c ? c->comm : nullptr, what, _r); return cudaErrorUnknown; } \
    } while (0)

int comm_version() {
    NcclApi* a = api(nullptr);
    return a ? a->version : 0;
}

int comm_unique_id(unsigned char* id, std::string* err) {
    NcclApi* a = api(err);
    if (!a) return 1;
    static_assert(sizeof(ncclUniqueId) == kCommIdBytes, "ncclUniqueId size");
    ncclUniqueId u;
    ncclResult_t r = a->GetUniqueId(&u);
    if (r != ncclSuccess) { if (err) *err = nccl_msg(a, nullptr, "ncclGetUniqueId", r); return 1; }
    memcpy(id, u.internal, kCommIdBytes);
    return 0;
}

static bool comm_alloc_scratch(Comm* c, std::string* err) {
    cudaError_t e = cudaMalloc((void**)&c->d_send, (size_t)kSlots * kSlotWords * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_recv, (size_t)kSlots * kSlotWords * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_gather, (size_t)kSlots * c->world * kCommMaxCols * sizeof(double));
    c->host_stage_bytes = (size_t)1 << 20;
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_host_stage, c->host_stage_bytes);
    if (e != cudaSuccess) { if (err) *err = std::string("device allocation for the communicator failed: ") + cudaGetErrorString(e); return false; }
    return true;
}

Comm* comm_create(const unsigned char* id, int rank, int world, std::string* err) {
    NcclApi* a = api(err);
    if (!a) return nullptr;
    Comm* c = new Comm();
    c->rank = rank; c->world = world;
    cudaGetDevice(&c->device);
    ncclUniqueId u;
    memcpy(u.internal, id, kCommIdBytes);
    ncclResult_t r = a->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { if (err) *err = nccl_msg(a, nullptr, "ncclCommInitRank", r); delete c; return nullptr; }
    if (!comm_alloc_scratch(c, err)) { comm_destroy(c); return nullptr; }
    return c;
}

int comm_create_all(int n, const int* devices, Comm** out, std::string* err) {
    NcclApi* a = api(err);
    if (!a) return 1;
    std::vector<ncclComm_t> comms((size_t)n);
    ncclResult_t r = a->CommInitAll(comms.data(), n, devices);
    if (r != ncclSuccess) { if (err) *err = nccl_msg(a, nullptr, "ncclCommInitAll", r); return 1; }
    int prev = 0;
    cudaGetDevice(&prev);
    bool ok = true;
    for (int i = 0; i < n; i++) {
        Comm* c = new Comm();
        c->comm = comms[i]; c->rank = i; c->world = n; c->device = devices[i];
        out[i] = c;
        cudaSetDevice(devices[i]);
        ok = ok && comm_alloc_scratch(c, err);
    }
    cudaSetDevice(prev);
    if (!ok) { for (int i = 0; i < n; i++) { comm_destroy(out[i]); out[i] = nullptr; } return 1; }
    return 0;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    NcclApi* a = api(nullptr);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (a && c->comm) a->CommDestroy(c->comm);
    if (c->d_send) cudaFree(c->d_send);
    if (c->d_recv) cudaFree(c->d_recv);
    if (c->d_gather) cudaFree(c->d_gather);
    if (c->d_host_stage) cudaFree(c->d_host_stage);
    cudaSetDevice(prev);
    delete c;
}

int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }

cudaError_t comm_combine(Comm* c, unsigned long long float_mask, int n, const AggDev* d_local, const unsigned long long* local_rows,
                         const unsigned int* local_panics, const unsigned int* local_chunks, AggDev* result, cudaStream_t s,
                         std::string* err) {
    NcclApi* a = api(err);
    if (!a) return cudaErrorUnknown;
    if (n < 1 || n > kCommMaxCols) { if (err) *err = "too many aggregates for one combine"; return cudaErrorInvalidValue; }
    const int slot = c->next++ % kSlots;
    unsigned long long* send = c->d_send + (size_t)slot * kSlotWords;
    unsigned long long* recv = c->d_recv + (size_t)slot * kSlotWords;
    double* gathered = c->d_gather + (size_t)slot * c->world * kCommMaxCols;
    Extra ex;
    for (int i = 0; i < n; i++) { ex.rows[i] = local_rows[i]; ex.panics[i] = local_panics ? local_panics[i] : 0u; ex.chunks[i] = local_chunks ? local_chunks[i] : 1u; }
    k_comm_pack<<<1, kCommMaxCols, 0, s>>>(d_local, n, ex, send);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // ONE group: the reductions are fused into a single NCCL launch.  Integer sums wrap (u64 ncclSum) and min/max
    // travel as order-preserving unsigned keys, so the result does not depend on the world size or the ring order.
    NCCL_TRY(a->GroupStart(), "ncclGroupStart");
    if (float_mask) NCCL_TRY(a->AllGather(send, gathered, (size_t)n, ncclFloat64, c->comm, s), "ncclAllGather");
    NCCL_TRY(a->AllReduce(send, recv, (size_t)5 * n, ncclUint64, ncclSum, c->comm, s), "ncclAllReduce(sum)");
    NCCL_TRY(a->AllReduce(send + 5 * n, recv + 5 * n, (size_t)n, ncclUint64, ncclMin, c->comm, s), "ncclAllReduce(min)");
    NCCL_TRY(a->AllReduce(send + 6 * n, recv + 6 * n, (size_t)n, ncclUint64, ncclMax, c->comm, s), "ncclAllReduce(max)");
    NCCL_TRY(a->GroupEnd(), "ncclGroupEnd");
    k_comm_unpack<<<1, kCommMaxCols, 0, s>>>(recv, gathered, n, c->world, float_mask, result);
    return cudaGetLastError();
}

cudaError_t comm_allreduce_max_i32(Comm* c, int* d_inout, int n, cudaStream_t s, std::string* err) {
    NcclApi* a = api(err);
    if (!a) return cudaErrorUnknown;
    NCCL_TRY(a->AllReduce(d_inout, d_inout, (size_t)n, ncclInt32, ncclMax, c->comm, s), "ncclAllReduce(flag)");
    return cudaSuccess;
}

cudaError_t comm_host_allreduce_f64(Comm* c, int op, double* inout, int n, cudaStream_t s, std::string* err) {
    NcclApi* a = api(err);
    if (!a) return cudaErrorUnknown;
    if ((size_t)n * sizeof(double) > c->host_stage_bytes) { if (err) *err = "host all-reduce too large"; return cudaErrorInvalidValue; }
    cudaError_t e = cudaMemcpyAsync(c->d_host_stage, inout, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    const ncclRedOp_t rop = op == 0 ? ncclSum : op == 1 ? ncclMin : ncclMax;
    NCCL_TRY(a->AllReduce(c->d_host_stage, c->d_host_stage, (size_t)n, ncclFloat64, rop, c->comm, s), "ncclAllReduce(host)");
    e = cudaMemcpyAsync(inout, c->d_host_stage, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    return e;
}

cudaError_t comm_host_allgather(Comm* c, const void* in, void* out, size_t bytes, cudaStream_t s, std::string* err) {
    NcclApi* a = api(err);
    if (!a) return cudaErrorUnknown;
    if (bytes * (size_t)c->world + (bytes + 255) / 256 * 256 > c->host_stage_bytes) { if (err) *err = "host all-gather too large"; return cudaErrorInvalidValue; }
    char* d_in = (char*)c->d_host_stage;
    char* d_out = d_in + (bytes + 255) / 256 * 256;
    cudaError_t e = cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    NCCL_TRY(a->AllGather(d_in, d_out, bytes, ncclInt8, c->comm, s), "ncclAllGather(host)");
    e = cudaMemcpyAsync(out, d_out, bytes * (size_t)c->world, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    return e;
}

}  // namespace bdf
