// chamfer_pipe.cu — chamfer_distance on HOST arrays: the sweep runs while the clouds are still crossing PCIe.
//
// The reference's array entry points (src/metrics/pcloud.jl:28-37) take host `Array`s and return a host Float32.  A
// drop-in that uploads 12·B·(N+M) bytes, then launches, then reads the scalar back pays the copy and two host round
// trips in full (cfg2: 3.1 MB ≈ 65-90 µs of PCIe + ≈ 30 µs for the read-back, against a ≈ 160 µs sweep).  Here
//   * ONE sweep grid is launched immediately; the batch is uploaded on a copy stream in chunks of 1, 2, 4, ... batch
//     elements, each followed (in stream order) by a 4-byte copy that raises the chunk's arrival flag; the sweep's CTAs
//     — dispatched in batch order — wait on the flag of their batch element.  The copy engine needs no SM, so the
//     waiting CTAs cannot starve it; the first ~100 KB is all that is exposed.  Same kernels, same arithmetic, same
//     bits as f3d_chamfer_fwd on resident inputs.
//   * the loss is stored by the finalize kernel straight into page-locked, device-mapped host memory and the host spins
//     on that word (falling back to the stream's status): no D2H copy, no driver synchronisation on the critical path.
//
// The handle owns a copy stream, two events and 64 bytes of mapped host memory (created once, off the hot path); every
// device byte — the staging copies of the clouds, the sweep workspace — lives in the caller's workspace.
#include <algorithm>
#include <cstring>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr uint32_t kSentinel = 0x7fc0dead;  // a NaN payload no computation produces: "loss not written yet"

struct Pipe {
    int max_chunks;
    int device;
    cudaStream_t copy;
    cudaEvent_t start, reset, copied;
    uint32_t* host;      // mapped page-locked: [0] loss bits, [1] the constant 1 (source of the arrival flags)
    uint32_t* host_dev;  // the same memory as the device sees it
};

struct PipePlan {
    int nchunks, m;
    size_t off_A, off_B, off_ws, ws_chamfer, off_loss, total;
};

PipePlan make_pipe_plan(int B, int N, int M, int max_chunks) {
    PipePlan pl;
    // largest chunk 2^m ≈ B/4: enough pieces to overlap, few enough that issuing the copies never limits the upload
    int m = 0;
    while ((8 << m) <= B) ++m;
    pl.m = m;
    int n = 1;
    while (n < std::min(max_chunks, kArriveMaxChunks) && arrive_chunk_begin(n, m) < B) ++n;
    if (n > 1 && B - arrive_chunk_begin(n - 1, m) < (1 << m) / 2) --n;  // a sliver at the end joins the previous chunk
    pl.nchunks = n;
    pl.ws_chamfer = align_up(f3d_chamfer_workspace_bytes(B, N, M), 256);
    size_t o = 0;
    pl.off_A = o;    o = align_up(o + sizeof(float) * 3 * (size_t)B * N, 256);
    pl.off_B = o;    o = align_up(o + sizeof(float) * 3 * (size_t)B * M, 256);
    pl.off_ws = o;   o += pl.ws_chamfer;
    pl.off_loss = o; o += 256;
    pl.total = o;
    return pl;
}

}  // namespace
}  // namespace f3d

using namespace f3d;

extern "C" int32_t f3d_chamfer_pipe_create(int32_t chunks, void** pipe) {
    if (!pipe) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_create: null handle pointer");
    if (chunks <= 0 || chunks > kArriveMaxChunks) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_create: chunks must be in 1..%d (got %d)", kArriveMaxChunks, chunks);
    Pipe* h = new Pipe();
    memset(h, 0, sizeof(*h));
    h->max_chunks = chunks;
    int lo = 0, hi = 0;
    cudaError_t e = cudaGetDevice(&h->device);
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    // non-blocking: must never synchronise implicitly with a legacy default stream that is running the waiting sweep
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->copy, cudaStreamNonBlocking, hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->reset, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->copied, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&h->host), 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->host_dev), h->host, 0);
    if (e != cudaSuccess) {
        delete h;  // a failure here is fatal for the process' CUDA context anyway; the partially created objects die with it
        return cuda_fail(e, "f3d_chamfer_pipe_create");
    }
    h->host[0] = kSentinel;
    h->host[1] = 1u;
    *pipe = h;
    return F3D_OK;
}

extern "C" size_t f3d_chamfer_pipe_workspace_bytes(int32_t B, int32_t N, int32_t M, int32_t chunks) {
    if (B <= 0 || N <= 0 || M <= 0 || chunks <= 0) return 0;
    return make_pipe_plan(B, N, M, chunks).total;
}

extern "C" int32_t f3d_chamfer_pipe_run(void* pipe, const float* A_host, const float* B_host, int32_t B, int32_t N, int32_t M,
                                        float w1, float w2, int32_t B_total, float* loss_dev, float* loss_host, void* ws,
                                        size_t ws_bytes, int32_t flags, f3d_stream_t stream_) {
    if (!pipe) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: null pipe handle");
    if (!A_host || !B_host) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: null host array");
    if (!loss_dev && !loss_host) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: give loss_dev, loss_host or both");
    if (B <= 0 || N <= 0 || M <= 0) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: B, N, M must be positive (got %d, %d, %d)", B, N, M);
    if (B > 65535) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: B must be <= 65535 per call");
    if (flags & F3D_FLAG_SWEEP_ONLY) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: F3D_FLAG_SWEEP_ONLY is a single-call measurement aid");
    if (B_total == 0) B_total = B;
    if (B_total < B) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: B_total (%d) < B (%d)", B_total, B);
    Pipe* h = static_cast<Pipe*>(pipe);
    // the arrival flags exist for the default (filtered) sweep; the cross-check modes upload first, then sweep
    const bool overlap = (flags & (F3D_FLAG_FMA | F3D_FLAG_EXACT_SWEEP)) == 0;
    const PipePlan pl = make_pipe_plan(B, N, M, overlap ? h->max_chunks : 1);
    if (!ws || ws_bytes < pl.total) return fail(F3D_ERR_WORKSPACE, "f3d_chamfer_pipe_run: workspace %zu < required %zu bytes", ws_bytes, pl.total);
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return fail(F3D_ERR_MISALIGNED, "f3d_chamfer_pipe_run: workspace must be 256-byte aligned");
    int dev = -1;
    F3D_CUDA(cudaGetDevice(&dev));
    if (dev != h->device) return fail(F3D_ERR_INVALID, "f3d_chamfer_pipe_run: pipe was created on device %d, current device is %d", h->device, dev);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    unsigned char* w = static_cast<unsigned char*>(ws);
    float* dA = reinterpret_cast<float*>(w + pl.off_A);
    float* dB = reinterpret_cast<float*>(w + pl.off_B);
    // where the finalize kernel stores the loss: straight into mapped host memory when the host wants it
    float* target = loss_host ? reinterpret_cast<float*>(h->host_dev) : loss_dev;
    if (loss_host) reinterpret_cast<volatile uint32_t*>(h->host)[0] = kSentinel;

    auto upload = [&](int c) -> cudaError_t {
        const int b0 = arrive_chunk_begin(c, pl.m), b1 = (c == pl.nchunks - 1) ? B : std::min(B, arrive_chunk_begin(c + 1, pl.m));
        cudaError_t e = cudaMemcpyAsync(dA + 3 * (size_t)b0 * N, A_host + 3 * (size_t)b0 * N, sizeof(float) * 3 * (size_t)(b1 - b0) * N, cudaMemcpyHostToDevice, h->copy);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dB + 3 * (size_t)b0 * M, B_host + 3 * (size_t)b0 * M, sizeof(float) * 3 * (size_t)(b1 - b0) * M, cudaMemcpyHostToDevice, h->copy);
        return e;
    };

    // the uploads are ordered after the work already queued on the caller's stream (the previous user of the workspace)
    F3D_CUDA(cudaEventRecord(h->start, stream));
    F3D_CUDA(cudaStreamWaitEvent(h->copy, h->start, 0));
    F3D_CUDA(upload(0));
    if (overlap) {
        ChamferArrive arr;
        arr.nchunks = pl.nchunks; arr.m = pl.m; arr.reset_done = h->reset; arr.flags_dev = nullptr;
        // memset of the counters + flags, record `reset`, then the sweep (whose CTAs wait for their chunk) + finalize
        const int32_t rc = chamfer_fwd_launch(dA, dB, B, N, M, w1, w2, B_total, target, nullptr, nullptr, nullptr, w + pl.off_ws, pl.ws_chamfer, flags, stream, &arr);
        if (rc != F3D_OK) return rc;
        F3D_CUDA(cudaStreamWaitEvent(h->copy, h->reset, 0));  // a flag may only be raised after this run's reset
        for (int c = 0; c < pl.nchunks; ++c) {
            if (c > 0) F3D_CUDA(upload(c));
            F3D_CUDA(cudaMemcpyAsync(arr.flags_dev + c, h->host + 1, sizeof(uint32_t), cudaMemcpyHostToDevice, h->copy));
        }
    } else {
        F3D_CUDA(cudaEventRecord(h->copied, h->copy));
        F3D_CUDA(cudaStreamWaitEvent(stream, h->copied, 0));
        const int32_t rc = chamfer_fwd_launch(dA, dB, B, N, M, w1, w2, B_total, target, nullptr, nullptr, nullptr, w + pl.off_ws, pl.ws_chamfer, flags, stream, nullptr);
        if (rc != F3D_OK) return rc;
    }
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
    uint32_t bits = slot[0];
    memcpy(loss_host, &bits, sizeof(float));
    if (overlap && bits == 0x7fc00000u) return fail(F3D_ERR_CUDA, "f3d_chamfer_pipe_run: an upload chunk did not arrive within 2 s (copy stream starved?)");
    return F3D_OK;
}

extern "C" int32_t f3d_chamfer_pipe_destroy(void* pipe) {
    if (!pipe) return F3D_OK;
    Pipe* h = static_cast<Pipe*>(pipe);
    const cudaError_t e = cudaStreamSynchronize(h->copy);
    cudaStreamDestroy(h->copy);
    cudaEventDestroy(h->start);
    cudaEventDestroy(h->reset);
    cudaEventDestroy(h->copied);
    cudaFreeHost(h->host);
    delete h;
    if (e != cudaSuccess) return cuda_fail(e, "f3d_chamfer_pipe_destroy");
    return F3D_OK;
}
