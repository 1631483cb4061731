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

// ---- the same combine without NCCL: every rank pushes its record straight into every peer's mailbox over NVLink ----------
// Mailbox of a rank (device memory, peer-mapped by the others): data[kP2pSlots][world][kSlotWords] then flags[kP2pSlots][world].
// Collective number `seq` uses slot seq % kP2pSlots; rank r writes its n x 7 words into data[slot][r] of EVERY rank (its own
// included), fences at system scope and then releases flags[slot][r] = seq on every rank.  A rank waits until all `world`
// flags of the slot carry seq, folds the records IN RANK ORDER (so float sums are deterministic for a given world size and
// integer results are order independent anyway) and writes the result.  One CTA, one launch, no library call: ~5 us where the
// grouped ncclAllReduce + pack + unpack cost 21 / 24 / 48 us at 2 / 4 / 8 GPUs.  A peer can be at most one collective ahead
// (it needs this rank's flag to finish the next one), so a few slots suffice.
constexpr int kP2pSlots = 8;
constexpr int kP2pThreads = 512;   // >= kSlotWords: one thread per word of the record

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kP2pThreads)
k_p2p_combine(unsigned long long* const* __restrict__ peers, int world, int rank, unsigned long long seq, const AggDev* __restrict__ local, int n,
              Extra ex, unsigned long long float_mask, AggDev* __restrict__ result) {
    const int tid = threadIdx.x;
    const int slot = (int)(seq % kP2pSlots);
    const size_t data_words = (size_t)kP2pSlots * world * kSlotWords;
    const int words = n * kFields;
    if (tid < words) {
        const int field = tid / n, i = tid - field * n;
        unsigned long long v;
        switch (field) {
            case 0: v = local[i].sum_bits; break;
            case 1: v = local[i].count; break;
            case 2: v = ex.rows[i]; break;
            case 3: v = ex.panics[i]; break;
            case 4: v = ex.chunks[i]; break;
            case 5: v = local[i].min_bits; break;
            default: v = local[i].max_bits; break;
        }
        for (int p = 0; p < world; p++) {
            const int q = (rank + p) % world;   // spread the first stores over the links
            peers[q][((size_t)slot * world + rank) * kSlotWords + tid] = v;
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid < world) st_release_sys(peers[tid] + data_words + (size_t)slot * world + rank, seq);
    if (tid < world) {
        const unsigned long long* f = peers[rank] + data_words + (size_t)slot * world + tid;
        while (ld_acquire_sys(f) != seq) __nanosleep(40);
    }
    __syncthreads();
    if (tid < n) {
        const unsigned long long* mine = peers[rank] + (size_t)slot * world * kSlotWords;
        AggDev a, b;
        a.sum_bits = 0; a.count = 0; a.min_bits = ~0ull; a.max_bits = 0;
        b.sum_bits = 0; b.min_bits = 0; b.max_bits = 0; b.count = 0;
        double fs = 0.0;
        const bool is_float = (float_mask >> tid) & 1ull;
        for (int r = 0; r < world; r++) {
            const unsigned long long* rec = mine + (size_t)r * kSlotWords;
            const unsigned long long sv = __ldcv(rec + 0 * n + tid);
            if (is_float) fs = __dadd_rn(fs, __longlong_as_double((long long)sv)); else a.sum_bits += sv;
            a.count += __ldcv(rec + 1 * n + tid);
            b.sum_bits += __ldcv(rec + 2 * n + tid);
            b.min_bits += __ldcv(rec + 3 * n + tid);
            b.max_bits += __ldcv(rec + 4 * n + tid);
            const unsigned long long mn = __ldcv(rec + 5 * n + tid), mx = __ldcv(rec + 6 * n + tid);
            a.min_bits = mn < a.min_bits ? mn : a.min_bits;
            a.max_bits = mx > a.max_bits ? mx : a.max_bits;
        }
        if (is_float) { a.sum_bits = (unsigned long long)__double_as_longlong(fs); a.min_bits = ~0ull; a.max_bits = 0ull; }
        result[2 * tid] = a;
        result[2 * tid + 1] = b;
    }
    __threadfence_system();
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    // NVLink mailboxes (see k_p2p_combine); d_peers[r] = rank r's mailbox as seen from this device
    unsigned long long* mailbox = nullptr;
    unsigned long long** d_peers = nullptr;
    std::vector<void*> ipc_opened;
    bool p2p = false, use_p2p = false;
    unsigned long long seq = 0;
    cudaEvent_t ev_p2p = nullptr;
    cudaStream_t last_stream = nullptr;
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
    if (c->ev_p2p) cudaEventDestroy(c->ev_p2p);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->d_peers) cudaFree(c->d_peers);
    if (c->mailbox) cudaFree(c->mailbox);
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

static size_t mailbox_bytes(int world) { return ((size_t)kP2pSlots * world * kSlotWords + (size_t)kP2pSlots * world) * sizeof(unsigned long long); }

// Process per GPU: exchange CUDA IPC handles of the mailboxes through the communicator itself and map the peers' memory.
int comm_enable_p2p(Comm* c, std::string* err) {
    if (c->p2p) return 0;
    if (c->world < 2) { if (err) *err = "a communicator of one rank has no peers"; return 1; }
    // Every rank runs the SAME two collectives whatever happens locally (a rank that bailed out early would leave the others
    // waiting); the second one carries each rank's verdict and the mailboxes are used only if ALL ranks succeeded.
    cudaStream_t s = nullptr;   // setup time: the legacy stream is fine
    char ok = 1;
    std::string why;
    cudaError_t e = cudaMalloc((void**)&c->mailbox, mailbox_bytes(c->world));
    if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, mailbox_bytes(c->world));
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine, c->mailbox);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the mailbox is zeroed before anybody learns its handle
    if (e != cudaSuccess) { ok = 0; why = std::string("mailbox allocation failed: ") + cudaGetErrorString(e); cudaGetLastError(); }
    std::vector<cudaIpcMemHandle_t> all((size_t)c->world);
    if (comm_host_allgather(c, &mine, all.data(), sizeof mine, s, err) != cudaSuccess) return 1;   // NCCL itself failed: nothing to agree on
    std::vector<unsigned long long*> peers((size_t)c->world, nullptr);
    for (int r = 0; r < c->world && ok; r++) {
        if (r == c->rank) { peers[r] = c->mailbox; continue; }
        void* p = nullptr;
        e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { ok = 0; why = std::string("cannot map the mailbox of rank ") + std::to_string(r) + ": " + cudaGetErrorString(e); cudaGetLastError(); break; }
        c->ipc_opened.push_back(p);
        peers[r] = (unsigned long long*)p;
    }
    if (ok) {
        e = cudaMalloc((void**)&c->d_peers, (size_t)c->world * sizeof(unsigned long long*));
        if (e == cudaSuccess) e = cudaMemcpy(c->d_peers, peers.data(), (size_t)c->world * sizeof(unsigned long long*), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { ok = 0; why = std::string("peer table allocation failed: ") + cudaGetErrorString(e); cudaGetLastError(); }
    }
    std::vector<char> verdicts((size_t)c->world, 0);
    if (comm_host_allgather(c, &ok, verdicts.data(), 1, s, err) != cudaSuccess) return 1;
    for (int r = 0; r < c->world; r++)
        if (!verdicts[r]) {
            if (err) *err = ok ? "rank " + std::to_string(r) + " could not set up its peer mailboxes" : why;
            return 1;
        }
    c->p2p = true;
    return 0;
}

// One process, all GPUs: enable peer access pairwise and use the mailbox pointers directly.
int comm_enable_p2p_all(int n, Comm** comms, std::string* err) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; i++) {
        cudaSetDevice(comms[i]->device);
        for (int j = 0; j < n; j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, comms[i]->device, comms[j]->device);
            if (!can) { if (err) *err = "the GPUs cannot access each other's memory"; cudaSetDevice(prev); return 1; }
            cudaError_t e2 = cudaDeviceEnablePeerAccess(comms[j]->device, 0);
            if (e2 != cudaSuccess && e2 != cudaErrorPeerAccessAlreadyEnabled) e = e2;
            cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMalloc((void**)&comms[i]->mailbox, mailbox_bytes(n));
        if (e == cudaSuccess) e = cudaMemset(comms[i]->mailbox, 0, mailbox_bytes(n));
    }
    for (int i = 0; i < n && e == cudaSuccess; i++) {
        cudaSetDevice(comms[i]->device);
        std::vector<unsigned long long*> peers((size_t)n);
        for (int j = 0; j < n; j++) peers[j] = comms[j]->mailbox;
        e = cudaMalloc((void**)&comms[i]->d_peers, (size_t)n * sizeof(unsigned long long*));
        if (e == cudaSuccess) e = cudaMemcpy(comms[i]->d_peers, peers.data(), (size_t)n * sizeof(unsigned long long*), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) { if (err) *err = std::string("peer mailbox setup failed: ") + cudaGetErrorString(e); cudaGetLastError(); return 1; }
    for (int i = 0; i < n; i++) comms[i]->p2p = true;
    return 0;
}

bool comm_has_p2p(const Comm* c) { return c->p2p; }
bool comm_uses_p2p(const Comm* c) { return c->use_p2p; }
int comm_set_p2p(Comm* c, bool on) {
    if (on && !c->p2p) return 1;
    c->use_p2p = on;
    return 0;
}

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
    if (c->use_p2p) {   // NVLink mailboxes instead of NCCL: one launch, see k_p2p_combine
        c->next--;      // the NCCL slot was not used
        // Collectives must EXECUTE in issue order on every rank (a rank may be at most one collective ahead of a peer, which is
        // what makes a few mailbox slots enough): when this one is enqueued on another stream than the previous one, chain them.
        cudaError_t e = cudaSuccess;
        if (!c->ev_p2p) e = cudaEventCreateWithFlags(&c->ev_p2p, cudaEventDisableTiming);
        if (e == cudaSuccess && c->seq > 0 && c->last_stream != s) e = cudaStreamWaitEvent(s, c->ev_p2p, 0);
        if (e != cudaSuccess) return e;
        k_p2p_combine<<<1, kP2pThreads, 0, s>>>(c->d_peers, c->world, c->rank, ++c->seq, d_local, n, ex, float_mask, result);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_p2p, s);
        c->last_stream = s;
        return e;
    }
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
