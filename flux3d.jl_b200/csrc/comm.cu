// comm.cu — the one data-path exchange of the sharded hot path: an all-reduce of the per-shard scalar
// loss over NCCL (NVLink 5 / NVSwitch).  The reference has no distributed code at all;
// this is what lets chamfer_distance / laplacian_loss shard on the batch axis, one process per GPU.
// NCCL is bound lazily with dlopen so libflux3d_b200.so has no link-time NCCL dependency and a
// single-GPU user never loads it.  If the process already holds an NCCL (e.g. torch's), dlopen by
// soname returns that same copy.
#include <dlfcn.h>

#include <cstring>

#include "f3d_common.cuh"

namespace f3d {
namespace {

struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_get_error_string get_error_string = nullptr;
};

constexpr int kNcclFloat32 = 7;  // ncclFloat32
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
    Comm* h = new Comm{c, nranks, rank};
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

extern "C" int32_t f3d_comm_destroy(void* comm) {
    if (!comm) return F3D_OK;
    NcclApi* api = nccl();
    Comm* h = static_cast<Comm*>(comm);
    int rc = api ? api->comm_destroy(h->nccl_comm) : 0;
    delete h;
    if (rc != 0) return nccl_fail(api, rc, "ncclCommDestroy");
    return F3D_OK;
}
