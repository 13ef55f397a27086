// chamfer_pipe.cu — chamfer_distance on HOST arrays: the sweep runs while the clouds are still crossing PCIe.
//
// The reference's array entry points (src/metrics/pcloud.jl:28-37) take host `Array`s and return a host Float32.  A
// drop-in that uploads 12·B·(N+M) bytes, then launches, then reads the scalar back pays the copy and two host round
// trips in full (cfg2: 3.1 MB ≈ 60 µs of PCIe + ≈ 30 µs for the read-back, on top of the sweep).  Here
//   * ONE grid is launched and it uploads for itself.  Tensor-core sweep (problems large enough for it): the two spare warps of
//     every sweep CTA pull the clouds out of page-locked host memory (16-byte loads over PCIe, batch element by batch element,
//     into a staging copy in HBM) and count every element as it lands; the producer waits per element and streams the raw
//     points, the converters centre them and take the norms on the fly — no prepare grid, no copy call, no SM taken from the
//     sweep.  CUDA-core sweep (small problems, F3D_FLAG_CUDA_CORES): the grid's first CTAs upload while the others sweep.
//   * the loss is stored by the grid straight into page-locked, device-mapped host memory and the host spins on that
//     word (falling back to the stream's status): no D2H copy, no driver synchronisation on the critical path.
// Host arrays the device cannot address (pageable memory) take the plain route: two cudaMemcpyAsync on the caller's stream,
// then the same kernels on resident inputs.  Same arithmetic, same bits as f3d_chamfer_fwd in every case.
//
// The handle owns 64 bytes of mapped host memory (created once, off the hot path); every device
// byte — the staging copies of the clouds, the sweep workspace — lives in the caller's workspace.
#include <algorithm>
#include <cstring>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr uint32_t kSentinel = 0x7fc0dead;  // a NaN payload no computation produces: "loss not written yet"
constexpr int kDefaultUploaders = 32;       // x 128 threads x 4 x 16 B = 256 KB in flight: enough to fill PCIe 5 x16

struct Pipe {
    int uploaders;
    int device;
    uint32_t* host;      // mapped page-locked: [0] loss bits
    uint32_t* host_dev;  // the same memory as the device sees it
};

struct PipePlan {
    size_t off_A, off_B, off_ws, ws_chamfer, total;
};

PipePlan make_pipe_plan(int B, int N, int M) {
    PipePlan pl;
    pl.ws_chamfer = align_up(f3d_chamfer_workspace_bytes(B, N, M), 256);
    size_t o = 0;
    pl.off_A = o;  o = align_up(o + sizeof(float) * 3 * (size_t)B * N, 256);
    pl.off_B = o;  o = align_up(o + sizeof(float) * 3 * (size_t)B * M, 256);
    pl.off_ws = o; o += pl.ws_chamfer;
    pl.total = o;
    return pl;
}

// the device's address for a page-locked host array, or null if the device cannot read it in place
const float* device_view(const float* host) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    if ((reinterpret_cast<uintptr_t>(at.devicePointer) & 15u) != 0) return nullptr;
    return static_cast<const float*>(at.devicePointer);
}

}  // namespace
}  // namespace f3d

using namespace f3d;

extern "C" int32_t f3d_chamfer_pipe_create(int32_t uploaders, void** pipe) {
    if (!pipe) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_create: null handle pointer");
    if (uploaders < 0 || uploaders > 1024) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_create: uploaders must be in 0..1024 (got %d)", uploaders);
    Pipe* h = new Pipe();
    memset(h, 0, sizeof(*h));
    h->uploaders = uploaders == 0 ? kDefaultUploaders : uploaders;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&h->host), 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->host_dev), h->host, 0);
    if (e != cudaSuccess) {
        delete h;
        return cuda_fail(e, "f3d_chamfer_pipe_create");
    }
    h->host[0] = kSentinel;
    *pipe = h;
    return F3D_OK;
}

extern "C" size_t f3d_chamfer_pipe_workspace_bytes(int32_t B, int32_t N, int32_t M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return make_pipe_plan(B, N, M).total;
}

extern "C" int32_t f3d_chamfer_pipe_run(void* pipe, const float* A_host, const float* B_host, int32_t B, int32_t N, int32_t M,
                                        float w1, float w2, int32_t B_total, float* loss_dev, float* loss_host, void* ws,
                                        size_t ws_bytes, int32_t flags, void* comm, f3d_stream_t stream_) {
    if (!pipe) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: null pipe handle");
    if (!A_host || !B_host) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: null host array");
    if (!loss_dev && !loss_host) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: give loss_dev, loss_host or both");
    if (B <= 0 || N <= 0 || M <= 0) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: B, N, M must be positive (got %d, %d, %d)", B, N, M);
    if (B > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: B must be <= 65535 per call");
    if (flags & F3D_FLAG_SWEEP_ONLY) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: F3D_FLAG_SWEEP_ONLY is a single-call measurement aid");
    if (B_total == 0) B_total = B;
    if (B_total < B) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: B_total (%d) < B (%d)", B_total, B);
    Pipe* h = static_cast<Pipe*>(pipe);
    const PipePlan pl = make_pipe_plan(B, N, M);
    if (!ws || ws_bytes < pl.total) return fail(F3D_ERR_WORKSPACE, "f3d_chamfer_pipe_run: workspace %zu < required %zu bytes", ws_bytes, pl.total);
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(F3D_ERR_MISALIGNED, "f3d_chamfer_pipe_run: workspace must be 256-byte aligned");
    int dev = -1;
    F3D_CUDA(cudaGetDevice(&dev));
    if (dev != h->device) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: pipe was created on device %d, current device is %d", h->device, dev);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    unsigned char* w = static_cast<unsigned char*>(ws);
    float* dA = reinterpret_cast<float*>(w + pl.off_A);
    float* dB = reinterpret_cast<float*>(w + pl.off_B);
    // where the grid stores the loss: straight into mapped host memory when the host wants it
    float* target = loss_host ? reinterpret_cast<float*>(h->host_dev) : loss_dev;
    if (loss_host) reinterpret_cast<volatile uint32_t*>(h->host)[0] = kSentinel;

    // sharded batch: the loss is summed over the ranks inside the step's last kernel (peer mailboxes over NVLink)
    ChamferPeerSum peer;
    const bool fused_sum = comm != nullptr;
    if (fused_sum) {
        if (flags & (F3D_FLAG_FMA | F3D_FLAG_EXACT_SWEEP)) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: the fused cross-rank sum exists only for the default sweep");
        if (!comm_peek_peer_sum(comm, &peer)) return F3D_ERR_NCCL;  // the error string is set
    }
    // the in-grid upload exists for the filtered sweeps on host memory the device can read in place: the tensor-core sweep pulls the
    // batch with its own spare warps; small problems (or F3D_FLAG_CUDA_CORES) take the CUDA-core sweep, whose first CTAs upload
    ChamferUpload up;
    up.A_host_dev = nullptr; up.B_host_dev = nullptr; up.uploaders = h->uploaders;
    if ((flags & (F3D_FLAG_FMA | F3D_FLAG_EXACT_SWEEP)) == 0) {
        up.A_host_dev = device_view(A_host);
        up.B_host_dev = up.A_host_dev ? device_view(B_host) : nullptr;
    }
    const bool in_grid = up.A_host_dev && up.B_host_dev;
    if (!in_grid) {
        F3D_CUDA(cudaMemcpyAsync(dA, A_host, sizeof(float) * 3 * (size_t)B * N, cudaMemcpyHostToDevice, stream));
        F3D_CUDA(cudaMemcpyAsync(dB, B_host, sizeof(float) * 3 * (size_t)B * M, cudaMemcpyHostToDevice, stream));
    }
    const int32_t rc = chamfer_fwd_launch(dA, dB, B, N, M, w1, w2, B_total, target, nullptr, nullptr, nullptr, w + pl.off_ws, pl.ws_chamfer,
                                          flags, stream, in_grid ? &up : nullptr, fused_sum ? &peer : nullptr);
    if (rc != F3D_OK) return rc;
    if (fused_sum) comm_commit_peer_sum(comm);  // only now: a rank-local failure above must not advance the step number
    if (!loss_host) return F3D_OK;

    if (loss_dev) F3D_CUDA(cudaMemcpyAsync(loss_dev, h->host_dev, sizeof(float), cudaMemcpyDefault, stream));
    // wait for the loss word; the stream's status is polled now and then so that a failed launch cannot hang the host
    volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(h->host);
    for (;;) {
        bool got = false;
        for (int i = 0; i < 2048 && !got; ++i) {
            got = slot[0] != kSentinel;
#if defined(__x86_64__) || defined(__i386__)
            if (!got) __builtin_ia32_pause();
#endif
        }
        if (got) break;
        const cudaError_t q = cudaStreamQuery(stream);
        if (q == cudaSuccess) break;  // finished: whatever is there is the result
        if (q != cudaErrorNotReady) return cuda_fail(q, "f3d_chamfer_pipe_run (waiting for the loss)");
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    if (loss_dev) F3D_CUDA(cudaStreamSynchronize(stream));
    const uint32_t bits = slot[0];
    memcpy(loss_host, &bits, sizeof(float));
    if (in_grid && bits == 0x7fc00000u) return fail(F3D_ERR_CUDA, "f3d_chamfer_pipe_run: a batch element did not arrive within 2 s");
    return F3D_OK;
}

extern "C" int32_t f3d_chamfer_pipe_destroy(void* pipe) {
    if (!pipe) return F3D_OK;
    Pipe* h = static_cast<Pipe*>(pipe);
    const cudaError_t e = cudaFreeHost(h->host);
    delete h;
    if (e != cudaSuccess) return cuda_fail(e, "f3d_chamfer_pipe_destroy");
    return F3D_OK;
}
