// f3d_common.cuh — shared device helpers and host-side error plumbing for libflux3d_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "flux3d_b200.h"

namespace f3d {

// ---- host: thread-local last-error string (the only mutable library state besides NCCL comms) ----
void set_error(const char* fmt, ...);
int32_t fail(int32_t code, const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what);

#define F3D_CHECK_LAUNCH(what)                                        \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return ::f3d::cuda_fail(e__, what);   \
    } while (0)

#define F3D_CUDA(call)                                                \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return ::f3d::cuda_fail(e__, #call);  \
    } while (0)

// ---- internal: f3d_chamfer_fwd with the optional in-grid upload (chamfer_pipe.cu): A / Bp are then staging buffers that
// the grid's first `uploaders` CTAs fill from page-locked host memory (device-accessible addresses) while the others sweep.
struct ChamferUpload {
    const float* A_host_dev;  // the host arrays as the device addresses them (cudaHostGetDevicePointer)
    const float* B_host_dev;
    int uploaders;
};
// ... and with the optional cross-rank sum of the loss through peer memory (comm.cu: f3d_comm_enable_p2p): mailboxes[r]
// is rank r's mailbox (2 x nranks 8-byte words) as THIS device addresses it; seq is the step number (same on every rank).
constexpr int kMaxPeerRanks = 64;
struct ChamferPeerSum {
    unsigned long long* const* mailboxes;  // device array [nranks]
    int nranks, rank;
    unsigned seq;
    unsigned long long timeout_ns;  // how long the finalize waits for a peer's word; 0 = no deadline (what an NCCL all-reduce does)
    unsigned* fault;                // device view of the communicator's page-locked fault word: set to seq when the deadline passes
};
int32_t chamfer_fwd_launch(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1, float w2,
                           int32_t B_total, float* loss_dev, float* terms_dev, int32_t* nnA_dev, int32_t* nnB_dev,
                           void* ws, size_t ws_bytes, int32_t flags, cudaStream_t stream, const ChamferUpload* upload,
                           const ChamferPeerSum* peer);
// comm.cu: the peer-sum descriptor of a communicator for its NEXT step.  Nothing advances until the launch that carries the
// descriptor has succeeded and comm_commit_peer_sum is called, so a rank-local failure (bad argument, workspace, launch
// error) leaves the step numbers of the ranks in agreement.  false: no peer mailboxes (f3d_comm_enable_p2p not called) or a
// previous step timed out (the error string says which).
bool comm_peek_peer_sum(void* comm, ChamferPeerSum* out);
void comm_commit_peer_sum(void* comm);
// chamfer_tc.cu: the filter sweep on the tensor cores (tcgen05 / TMEM); same results as the CUDA-core sweep of chamfer.cu
size_t chamfer_tc_workspace_bytes(int B, int N, int M);
bool chamfer_tc_possible(int B, int N, int M);   // hard limits of the path
bool chamfer_tc_supported(int B, int N, int M);  // ... and large enough to be the default
bool chamfer_tc_upload_possible(int N, int M);   // ... on host arrays pulled over PCIe by the sweep's own spare warps (ChamferUpload)
int32_t chamfer_tc_launch(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1, float w2, int32_t B_total,
                          float* loss_dev, float* terms_dev, int32_t* nnA_dev, int32_t* nnB_dev, void* ws, size_t ws_bytes, int32_t flags,
                          cudaStream_t stream, const ChamferUpload* upload, const ChamferPeerSum* peer);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device: sm_100 packed-FP32 (f32x2) arithmetic.  Each op is two independently rounded (RN)
// binary32 operations issued as ONE FADD2/FMUL2/FFMA2 instruction. --------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Squared distance in the reference's arithmetic, scalar form: ((dx*dx)+(dy*dy))+(dz*dz), every
// operation separately rounded (never contracted), or the FMA form when kFma.
template <bool kFma>
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    if (kFma) return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- device: the one exchange of the sharded chamfer path, fused into the finalize kernels --------------------------------
// Thread `t` (< nranks) of the finalize's last block stores this rank's shard loss — already divided by the GLOBAL
// N*B_total / M*B_total — into its slot of rank t's mailbox over NVLink as ONE 8-byte word {step number, float bits}
// (value and flag arrive together), then waits for rank t's word in its own mailbox.  The caller adds the nranks values in
// rank order: the same bits on every rank, no NCCL kernel, no extra launch.  Slots are double-buffered by the parity of the
// step number: a rank can only be two steps ahead of a peer after that peer has sent its word for the step in between,
// i.e. after it finished reading the older one.  A peer that is late is simply waited for (first-iteration module loads,
// a checkpoint on one rank, a dataloader stall are routine); only when the communicator's deadline passes (default 30 min,
// F3D_PEER_TIMEOUT_S; 0 = none) the value becomes NaN AND the communicator's fault word is set, so that the next call on
// any entry point that takes the communicator fails with a status instead of training on.
__device__ __forceinline__ float peer_exchange(const ChamferPeerSum& peer, int t, float mine) {
    const unsigned long long word = ((unsigned long long)peer.seq << 32) | (unsigned long long)__float_as_uint(mine);
    const int base = (int)(peer.seq & 1u) * peer.nranks;
    unsigned long long* dst = peer.mailboxes[t] + base + peer.rank;  // my slot in rank t's mailbox
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    const unsigned long long* src = peer.mailboxes[peer.rank] + base + t;  // rank t's slot in my mailbox
    unsigned long long got = 0, t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
        if ((unsigned)(got >> 32) == peer.seq) return __uint_as_float((unsigned)got);
        if (spins > 64) __nanosleep(spins > 4096 ? 1000 : 50);
        if (peer.timeout_ns && (spins & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > peer.timeout_ns) {
                if (peer.fault) { *reinterpret_cast<volatile unsigned*>(peer.fault) = peer.seq; __threadfence_system(); }
                return __int_as_float(0x7fc00000);
            }
        }
    }
}

}  // namespace f3d
