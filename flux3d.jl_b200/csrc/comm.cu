// comm.cu — the one data-path exchange of the sharded hot path: an all-reduce of the per-shard scalar
// loss over NCCL (NVLink 5 / NVSwitch).  The reference has no distributed code at all;
// this is what lets chamfer_distance / laplacian_loss shard on the batch axis, one process per GPU.
// NCCL is bound lazily with dlopen so libflux3d_b200.so has no link-time NCCL dependency and a
// single-GPU user never loads it.  If the process already holds an NCCL (e.g. torch's), dlopen by
// soname returns that same copy.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

#include "f3d_common.cuh"

namespace f3d {
namespace {

struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_get_error_string get_error_string = nullptr;
};

constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclChar = 0;     // ncclInt8 / ncclChar
constexpr int kNcclSum = 0;      // ncclSum

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* env = getenv("FLUX3D_B200_NCCL");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.get_unique_id = (fn_get_unique_id)dlsym(api.handle, "ncclGetUniqueId");
            api.comm_init_rank = (fn_comm_init_rank)dlsym(api.handle, "ncclCommInitRank");
            api.all_reduce = (fn_all_reduce)dlsym(api.handle, "ncclAllReduce");
            api.all_gather = (fn_all_gather)dlsym(api.handle, "ncclAllGather");
            api.comm_destroy = (fn_comm_destroy)dlsym(api.handle, "ncclCommDestroy");
            api.get_error_string = (fn_get_error_string)dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy) return nullptr;
    return &api;
}

int32_t nccl_fail(NcclApi* api, int rc, const char* what) {
    return fail(F3D_ERR_NCCL, "%s: NCCL error %d (%s)", what, rc, api->get_error_string ? api->get_error_string(rc) : "?");
}

struct Comm {
    void* nccl_comm;
    int nranks, rank;
    // peer mailboxes (f3d_comm_enable_p2p): the fused cross-rank sum of the chamfer finalize writes straight into them
    unsigned long long* mailbox = nullptr;        // this rank's: [2][nranks] words {step number << 32 | float bits}
    unsigned long long** mailboxes_dev = nullptr; // device array [nranks]: every rank's mailbox as this device addresses it
    void* opened[kMaxPeerRanks] = {};             // cudaIpcOpenMemHandle mappings to close
    unsigned seq = 0;                             // steps issued so far (the same on every rank: collective calls only)
    unsigned* fault_host = nullptr;               // page-locked, device-mapped: the step number whose exchange timed out (0 = none)
    unsigned* fault_dev = nullptr;                // the same word as the device addresses it
    unsigned long long timeout_ns = 0;            // deadline of one exchange; 0 = wait for ever
};

}  // namespace
}  // namespace f3d

using namespace f3d;

extern "C" int32_t f3d_comm_unique_id_host(void* id128_host) {
    if (!id128_host) return fail(F3D_ERR_INVALID, "f3d_comm_unique_id_host: null pointer");
    NcclApi* api = nccl();
    if (!api) return fail(F3D_ERR_NCCL, "f3d_comm_unique_id_host: cannot load libnccl.so.2 (set FLUX3D_B200_NCCL to its path)");
    NcclId id;
    int rc = api->get_unique_id(&id);
    if (rc != 0) return nccl_fail(api, rc, "ncclGetUniqueId");
    memcpy(id128_host, &id, sizeof(id));
    return F3D_OK;
}

extern "C" int32_t f3d_comm_init(int32_t nranks, int32_t rank, const void* id128_host, void** comm) {
    if (!id128_host || !comm) return fail(F3D_ERR_INVALID, "f3d_comm_init: null pointer");
    if (nranks <= 0 || rank < 0 || rank >= nranks) return fail(F3D_ERR_INVALID, "f3d_comm_init: bad rank %d of %d", rank, nranks);
    NcclApi* api = nccl();
    if (!api) return fail(F3D_ERR_NCCL, "f3d_comm_init: cannot load libnccl.so.2 (set FLUX3D_B200_NCCL to its path)");
    NcclId id;
    memcpy(&id, id128_host, sizeof(id));
    void* c = nullptr;
    int rc = api->comm_init_rank(&c, nranks, id, rank);
    if (rc != 0) return nccl_fail(api, rc, "ncclCommInitRank");
    Comm* h = new Comm();
    h->nccl_comm = c; h->nranks = nranks; h->rank = rank;
    *comm = h;
    return F3D_OK;
}

extern "C" int32_t f3d_allreduce_sum_f32(void* comm, float* dev_buf, int32_t count, f3d_stream_t stream) {
    if (!comm || !dev_buf || count <= 0) return fail(F3D_ERR_INVALID, "f3d_allreduce_sum_f32: bad arguments");
    NcclApi* api = nccl();
    if (!api) return fail(F3D_ERR_NCCL, "f3d_allreduce_sum_f32: NCCL not loaded");
    Comm* h = static_cast<Comm*>(comm);
    int rc = api->all_reduce(dev_buf, dev_buf, (size_t)count, kNcclFloat32, kNcclSum, h->nccl_comm, static_cast<cudaStream_t>(stream));
    if (rc != 0) return nccl_fail(api, rc, "ncclAllReduce");
    return F3D_OK;
}

bool f3d::comm_peek_peer_sum(void* comm, ChamferPeerSum* out) {
    Comm* h = static_cast<Comm*>(comm);
    if (!h || !h->mailboxes_dev) {
        set_error("communicator has no peer mailboxes (call f3d_comm_enable_p2p)");
        return false;
    }
    if (h->fault_host && *reinterpret_cast<volatile unsigned*>(h->fault_host) != 0u) {
        set_error("the fused cross-rank sum of step %u timed out (a peer never delivered its loss): the communicator is dead, the loss of that step was NaN",
                  *reinterpret_cast<volatile unsigned*>(h->fault_host));
        return false;
    }
    out->mailboxes = h->mailboxes_dev;
    out->nranks = h->nranks;
    out->rank = h->rank;
    out->seq = h->seq + 1;  // 1, 2, ...: never 0, the value the mailboxes start with
    out->timeout_ns = h->timeout_ns;
    out->fault = h->fault_dev;
    return true;
}

void f3d::comm_commit_peer_sum(void* comm) { ++static_cast<Comm*>(comm)->seq; }

// Peer mailboxes: every rank allocates 2 x nranks 8-byte words, exports them with CUDA IPC, gathers everybody's handle
// over the existing NCCL communicator and maps the peers' mailboxes (NVLink peer access).  Collective; off the hot path.
extern "C" int32_t f3d_comm_enable_p2p(void* comm, f3d_stream_t stream_) {
    if (!comm) return fail(F3D_ERR_INVALID, "f3d_comm_enable_p2p: null communicator");
    NcclApi* api = nccl();
    if (!api || !api->all_gather) return fail(F3D_ERR_NCCL, "f3d_comm_enable_p2p: ncclAllGather not available");
    Comm* h = static_cast<Comm*>(comm);
    if (h->mailboxes_dev) return F3D_OK;
    if (h->nranks > kMaxPeerRanks) return fail(F3D_ERR_INVALID, "f3d_comm_enable_p2p: at most %d ranks", kMaxPeerRanks);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    {
        // how long one exchange waits for a late peer before it gives up and marks the communicator dead: long enough for
        // every routine stall (module load, checkpoint, evaluation on one rank); F3D_PEER_TIMEOUT_S = 0 waits for ever
        double secs = 1800.0;
        if (const char* env = getenv("F3D_PEER_TIMEOUT_S")) secs = atof(env);
        h->timeout_ns = secs > 0.0 ? (unsigned long long)(secs * 1e9) : 0ull;
    }
    // A rank whose LOCAL step fails must still take part in both collectives (with an all-zero handle / a zero vote):
    // otherwise its peers would wait in NCCL for ever.  Every rank then returns the same verdict.
    const size_t words = 2 * (size_t)h->nranks;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    cudaError_t e = cudaSuccess;
    if (!h->fault_host) {
        e = cudaHostAlloc(reinterpret_cast<void**>(&h->fault_host), 64, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) { *h->fault_host = 0u; e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->fault_dev), h->fault_host, 0); }
    }
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->mailbox), words * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->mailbox, 0, words * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine, h->mailbox);
    if (e != cudaSuccess) { memset(&mine, 0, sizeof(mine)); cudaGetLastError(); }
    cudaError_t first_err = e;
    unsigned char* scratch = nullptr;  // [nranks][64] handles, then one float vote
    F3D_CUDA(cudaMalloc(reinterpret_cast<void**>(&scratch), 64 * (size_t)h->nranks + 64));
    F3D_CUDA(cudaMemcpyAsync(scratch + 64 * (size_t)h->rank, &mine, 64, cudaMemcpyHostToDevice, stream));
    int rc = api->all_gather(scratch + 64 * (size_t)h->rank, scratch, 64, kNcclChar, h->nccl_comm, stream);
    if (rc != 0) { cudaFree(scratch); return nccl_fail(api, rc, "ncclAllGather"); }
    cudaIpcMemHandle_t all[kMaxPeerRanks];
    F3D_CUDA(cudaMemcpyAsync(all, scratch, 64 * (size_t)h->nranks, cudaMemcpyDeviceToHost, stream));
    F3D_CUDA(cudaStreamSynchronize(stream));
    bool ok = first_err == cudaSuccess;
    unsigned long long* ptrs[kMaxPeerRanks] = {};
    static const cudaIpcMemHandle_t zero = {};
    for (int r = 0; r < h->nranks && ok; ++r) {
        if (memcmp(&all[r], &zero, sizeof(zero)) == 0) { ok = false; break; }  // that rank could not export a mailbox
        if (r == h->rank) { ptrs[r] = h->mailbox; continue; }
        void* mapped = nullptr;
        e = cudaIpcOpenMemHandle(&mapped, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { if (first_err == cudaSuccess) first_err = e; cudaGetLastError(); ok = false; break; }
        h->opened[r] = mapped;
        ptrs[r] = static_cast<unsigned long long*>(mapped);
    }
    if (ok) {
        e = cudaMalloc(reinterpret_cast<void**>(&h->mailboxes_dev), sizeof(unsigned long long*) * (size_t)h->nranks);
        if (e == cudaSuccess) e = cudaMemcpy(h->mailboxes_dev, ptrs, sizeof(unsigned long long*) * (size_t)h->nranks, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { if (first_err == cudaSuccess) first_err = e; cudaGetLastError(); ok = false; }
    }
    // the vote doubles as the barrier "nobody writes into a mailbox before its owner has zeroed it"
    float* vote = reinterpret_cast<float*>(scratch + 64 * (size_t)h->nranks);
    const float myvote = ok ? 1.0f : 0.0f;
    F3D_CUDA(cudaMemcpyAsync(vote, &myvote, sizeof(float), cudaMemcpyHostToDevice, stream));
    rc = api->all_reduce(vote, vote, 1, kNcclFloat32, kNcclSum, h->nccl_comm, stream);
    if (rc != 0) { cudaFree(scratch); return nccl_fail(api, rc, "ncclAllReduce (vote)"); }
    float votes = 0.0f;
    F3D_CUDA(cudaMemcpyAsync(&votes, vote, sizeof(float), cudaMemcpyDeviceToHost, stream));
    F3D_CUDA(cudaStreamSynchronize(stream));
    F3D_CUDA(cudaFree(scratch));
    if ((int)(votes + 0.5f) != h->nranks) {
        // somebody failed: everybody gives the mailboxes up and reports it (the NCCL all-reduce path still works)
        for (int r = 0; r < kMaxPeerRanks; ++r)
            if (h->opened[r]) { cudaIpcCloseMemHandle(h->opened[r]); h->opened[r] = nullptr; }
        if (h->mailboxes_dev) { cudaFree(h->mailboxes_dev); h->mailboxes_dev = nullptr; }
        if (h->mailbox) { cudaFree(h->mailbox); h->mailbox = nullptr; }
        if (first_err != cudaSuccess) return cuda_fail(first_err, "f3d_comm_enable_p2p (this rank)");
        return fail(F3D_ERR_CUDA, "f3d_comm_enable_p2p: %d of %d ranks could not map the peer mailboxes", h->nranks - (int)(votes + 0.5f), h->nranks);
    }
    return F3D_OK;
}

// f3d_chamfer_fwd on one shard of a batch split across the ranks of `comm`, with the sum of the shard losses fused into
// the finalize kernel (peer mailboxes over NVLink): loss_dev receives the loss of the WHOLE batch on every rank.
extern "C" int32_t f3d_chamfer_fwd_allreduce(void* comm, const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1,
                                             float w2, int32_t B_total, float* loss_dev, int32_t* nnA_dev, int32_t* nnB_dev,
                                             void* ws, size_t ws_bytes, int32_t flags, f3d_stream_t stream_) {
    if (!comm) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd_allreduce: null communicator");
    if (flags & ~F3D_FLAG_CUDA_CORES) return fail(F3D_ERR_INVALID, "f3d_chamfer_fwd_allreduce: only the default sweeps have the fused cross-rank sum");
    ChamferPeerSum peer;
    if (!comm_peek_peer_sum(comm, &peer)) return F3D_ERR_NCCL;  // the error string is set
    // validate and launch first; the step number advances only once the finalize that sends this rank's word is in the stream
    const int32_t rc = chamfer_fwd_launch(A, Bp, B, N, M, w1, w2, B_total, loss_dev, nullptr, nnA_dev, nnB_dev, ws, ws_bytes, flags,
                                          static_cast<cudaStream_t>(stream_), nullptr, &peer);
    if (rc == F3D_OK) comm_commit_peer_sum(comm);
    return rc;
}

extern "C" int32_t f3d_comm_destroy(void* comm) {
    if (!comm) return F3D_OK;
    NcclApi* api = nccl();
    Comm* h = static_cast<Comm*>(comm);
    for (int r = 0; r < kMaxPeerRanks; ++r)
        if (h->opened[r]) cudaIpcCloseMemHandle(h->opened[r]);
    if (h->mailboxes_dev) cudaFree(h->mailboxes_dev);
    if (h->mailbox) cudaFree(h->mailbox);
    if (h->fault_host) cudaFreeHost(h->fault_host);
    int rc = api ? api->comm_destroy(h->nccl_comm) : 0;
    delete h;
    if (rc != 0) return nccl_fail(api, rc, "ncclCommDestroy");
    return F3D_OK;
}
