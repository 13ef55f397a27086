// knn.cu — batched exact kNN graph + EdgeConv prologue for sm_100a.
//
// Replaces Flux3D.jl src/models/dgcnn.jl:3-9 (CreateSingleKNNGraph: KD-tree build + N single-point
// knn queries per cloud, CPU only) and :32-45 (batch loop, cat to (F,K,N,B), tile, cat(X, KNN - X)).
//
// Semantics (SURVEY §8a rule 3): for every point the (K+1) nearest points of its own cloud sorted
// ascending by (squared distance, index); positions 2..K+1 are returned, i.e. the first hit — the
// query itself, or its lowest-indexed exact duplicate — is dropped BY POSITION as dgcnn.jl:6 does.
// Distances are s = 0; s = s + (a_d - b_d)^2 for d = 1..F, every operation separately rounded.
//
// Shape of the work: B*N query rows, each against the N candidates of its cloud.  One warp owns
// kQPW queries; a CTA (8 warps) owns 8*kQPW consecutive queries of one cloud and streams the cloud
// through shared memory in tiles of kTileC candidates (rows padded to Fp+4 floats so that LDS.128 is
// bank-conflict-free for row-per-lane access).  Each lane evaluates kCPL candidates x kQPW queries in
// registers.  The running (K+1)-best list of a query lives in registers distributed over the warp
// (element e in lane e%32, slot e/32) as 64-bit keys (distance bits << 32 | index): distances are
// >= 0, so unsigned integer order on the key IS the (distance, index) order.  A candidate is
// inserted only if its key is below the current (K+1)-th key (ballot), by a shuffle-shift.
#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kWarpsK = 8;
constexpr int kThreadsK = 32 * kWarpsK;
constexpr int kQPW = 4;                   // queries per warp
constexpr int kQPC = kWarpsK * kQPW;      // queries per CTA (32)
constexpr int kCPL = 4;                   // candidates per lane per tile
constexpr int kTileC = 32 * kCPL;         // candidates per tile (128)
constexpr u64 kKeyInf = ~0ull;

struct KnnParams {
    const float* X;  // [B][N][F]
    int N, F, Fp, K; // Fp = F rounded up to a multiple of 4 (zero padded: adds +0 exactly)
    int32_t* idx;    // [B][N][K]
    float* dist;     // [B][N][K] or null
    float* gathered; // [B][N][K][F] or null
    float* edge;     // [B][N][K][2F] or null
};

template <int kSlots>
struct TopList {
    u64 v[kSlots];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) v[s] = kKeyInf;
    }
    // element e (0-based rank) broadcast to the warp
    __device__ __forceinline__ u64 get(int e) const {
        u64 r = 0;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            u64 t = __shfl_sync(0xffffffffu, v[s], e & 31);
            if ((e >> 5) == s) r = t;
        }
        return r;
    }
    // insert key (warp-uniform) keeping ascending order; the last element falls off
    __device__ __forceinline__ void insert(u64 key, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) pos += __popc(__ballot_sync(0xffffffffu, v[s] < key));
        u64 carry = key;  // value shifted into lane 0 of the next slot
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            u64 up = __shfl_up_sync(0xffffffffu, v[s], 1);
            u64 last = __shfl_sync(0xffffffffu, v[s], 31);
            if (lane == 0) up = carry;
            carry = last;
            const int e = s * 32 + lane;
            v[s] = (e < pos) ? v[s] : ((e == pos) ? key : up);
        }
    }
};

template <int kSlots>
__global__ void __launch_bounds__(kThreadsK) knn_graph_kernel(KnnParams p) {
    extern __shared__ __align__(16) float smem_k[];
    const int stride = p.Fp + 4;           // floats per staged row
    float* s_q = smem_k;                   // [kQPC][stride]
    float* s_c = s_q + kQPC * stride;      // [kTileC][stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int q0 = blockIdx.x * kQPC;
    const float* Xb = p.X + (size_t)b * p.N * p.F;

    // ---- stage this CTA's queries (zero-padded to Fp, rows past N replicate row N-1: never stored)
    for (int t = tid; t < kQPC * p.Fp; t += kThreadsK) {
        int r = t / p.Fp, d = t - r * p.Fp;
        int qi = min(q0 + r, p.N - 1);
        s_q[r * stride + d] = (d < p.F) ? __ldg(Xb + (size_t)qi * p.F + d) : 0.0f;
    }

    TopList<kSlots> top[kQPW];
    u64 thr[kQPW];
#pragma unroll
    for (int q = 0; q < kQPW; ++q) { top[q].init(); thr[q] = kKeyInf; }

    const float* qrow = s_q + (warp * kQPW) * stride;
    for (int c0 = 0; c0 < p.N; c0 += kTileC) {
        __syncthreads();  // previous tile fully consumed (and s_q visible on the first pass)
        const int nc = min(kTileC, p.N - c0);
        for (int t = tid; t < kTileC * p.Fp; t += kThreadsK) {
            int r = t / p.Fp, d = t - r * p.Fp;
            float v = 0.0f;
            if (r < nc && d < p.F) v = __ldg(Xb + (size_t)(c0 + r) * p.F + d);
            s_c[r * stride + d] = v;
        }
        __syncthreads();

        float acc[kQPW][kCPL];
#pragma unroll
        for (int q = 0; q < kQPW; ++q)
#pragma unroll
            for (int c = 0; c < kCPL; ++c) acc[q][c] = 0.0f;

        for (int d = 0; d < p.Fp; d += 4) {
            float4 cv[kCPL], qv[kQPW];
#pragma unroll
            for (int c = 0; c < kCPL; ++c) cv[c] = *reinterpret_cast<const float4*>(s_c + (c * 32 + lane) * stride + d);
#pragma unroll
            for (int q = 0; q < kQPW; ++q) qv[q] = *reinterpret_cast<const float4*>(qrow + q * stride + d);
#pragma unroll
            for (int q = 0; q < kQPW; ++q)
#pragma unroll
                for (int c = 0; c < kCPL; ++c) {
                    float t;
                    t = __fsub_rn(qv[q].x, cv[c].x); acc[q][c] = __fadd_rn(acc[q][c], __fmul_rn(t, t));
                    t = __fsub_rn(qv[q].y, cv[c].y); acc[q][c] = __fadd_rn(acc[q][c], __fmul_rn(t, t));
                    t = __fsub_rn(qv[q].z, cv[c].z); acc[q][c] = __fadd_rn(acc[q][c], __fmul_rn(t, t));
                    t = __fsub_rn(qv[q].w, cv[c].w); acc[q][c] = __fadd_rn(acc[q][c], __fmul_rn(t, t));
                }
        }

        // ---- threshold-filtered insertion into the distributed sorted lists ------------------------
#pragma unroll
        for (int q = 0; q < kQPW; ++q) {
#pragma unroll
            for (int c = 0; c < kCPL; ++c) {
                const int j = c0 + c * 32 + lane;
                const u64 key = (j < p.N) ? (((u64)__float_as_uint(acc[q][c]) << 32) | (unsigned)j) : kKeyInf;
                unsigned m = __ballot_sync(0xffffffffu, key < thr[q]);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const u64 kk = __shfl_sync(0xffffffffu, key, src);
                    if (kk < thr[q]) {  // warp-uniform; thr may have tightened since the ballot
                        top[q].insert(kk, lane);
                        thr[q] = top[q].get(p.K);
                    }
                }
            }
        }
    }

    // ---- emit ranks 1..K of each query (rank 0 dropped by position, dgcnn.jl:6) --------------------
#pragma unroll
    for (int q = 0; q < kQPW; ++q) {
        const int qi = q0 + warp * kQPW + q;
        if (qi >= p.N) continue;  // warp-uniform
        const size_t obase = ((size_t)b * p.N + qi) * p.K;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            const int e = s * 32 + lane;
            if (e >= 1 && e <= p.K) {
                const u64 key = top[q].v[s];
                p.idx[obase + e - 1] = (int32_t)(unsigned)(key & 0xffffffffu);
                if (p.dist) p.dist[obase + e - 1] = __uint_as_float((unsigned)(key >> 32));
            }
        }
        if (p.gathered || p.edge) {
            const float* xi = Xb + (size_t)qi * p.F;
            for (int k = 0; k < p.K; ++k) {
                const int j = (int)(unsigned)(top[q].get(k + 1) & 0xffffffffu);
                const float* xj = Xb + (size_t)j * p.F;
                for (int d = lane; d < p.F; d += 32) {
                    const float vj = __ldg(xj + d);
                    if (p.gathered) p.gathered[(obase + k) * p.F + d] = vj;
                    if (p.edge) {  // cat(X, KNNGraph - X; dims=1)  dgcnn.jl:45
                        const float vi = __ldg(xi + d);
                        float* o = p.edge + (obase + k) * 2 * p.F;
                        o[d] = vi;
                        o[p.F + d] = __fsub_rn(vj, vi);
                    }
                }
            }
        }
    }
}

size_t knn_smem_bytes(int Fp) { return sizeof(float) * (size_t)(kQPC + kTileC) * (Fp + 4); }

}  // namespace
}  // namespace f3d

extern "C" size_t f3d_knn_graph_workspace_bytes(int32_t, int32_t, int32_t, int32_t) { return 0; }

extern "C" int32_t f3d_knn_graph(const float* X, int32_t B, int32_t N, int32_t F, int32_t K, int32_t* idx,
                                 float* dist, float* gathered, float* edge_feat, void* /*ws*/, size_t /*ws_bytes*/,
                                 int32_t flags, f3d_stream_t stream_) {
    using namespace f3d;
    if (!X || !idx) return fail(F3D_ERR_INVALID, "f3d_knn_graph: null X/idx pointer");
    if (B <= 0 || N <= 0 || F <= 0) return fail(F3D_ERR_INVALID, "f3d_knn_graph: B, N, F must be positive (got %d, %d, %d)", B, N, F);
    if (K < 1 || K >= N) return fail(F3D_ERR_INVALID, "f3d_knn_graph: need 1 <= K < N (K=%d, N=%d)", K, N);
    if (K > 63) return fail(F3D_ERR_INVALID, "f3d_knn_graph: K must be <= 63 (got %d)", K);
    if (F > 256) return fail(F3D_ERR_INVALID, "f3d_knn_graph: F must be <= 256 (got %d)", F);
    if (B > 65535) return fail(F3D_ERR_INVALID, "f3d_knn_graph: B must be <= 65535 per call");
    if (flags != F3D_FLAG_NONE) return fail(F3D_ERR_INVALID, "f3d_knn_graph: no flags are defined for this call");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    KnnParams p;
    p.X = X; p.N = N; p.F = F; p.Fp = (F + 3) / 4 * 4; p.K = K;
    p.idx = idx; p.dist = dist; p.gathered = gathered; p.edge = edge_feat;
    const size_t smem = knn_smem_bytes(p.Fp);
    dim3 grid((N + kQPC - 1) / kQPC, B);
    if (K + 1 <= 32) {
        F3D_CUDA(cudaFuncSetAttribute(knn_graph_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_graph_kernel<1><<<grid, kThreadsK, smem, stream>>>(p);
    } else {
        F3D_CUDA(cudaFuncSetAttribute(knn_graph_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_graph_kernel<2><<<grid, kThreadsK, smem, stream>>>(p);
    }
    F3D_CHECK_LAUNCH("knn_graph_kernel");
    return F3D_OK;
}
