// sampling.cu — sample_points (TriMesh -> PointCloud) for sm_100a.
//
// Replaces Flux3D.jl src/transforms/mesh_func.jl:21-82: per mesh a host loop with a D2H copy of the
// face probabilities, CPU alias sampling (Distributions.Categorical), CPU `rand`, index uploads and
// ~10 broadcast launches.  Here: ONE launch for the whole batch.  A CTA owns (mesh, chunk of
// samples); it rebuilds that mesh's face CDF in shared memory (areas in the reference's Float32
// arithmetic, probabilities area/max(Σarea,eps) in Float64 as at :32-39), then draws with a counter
// RNG (Philox4x32-10, counter = (sample, mesh, offset), key = seed) and inverts the CDF by binary
// search.  The CDF is held in 53-bit FIXED POINT (floor(p_f * 2^53) summed as integers): integer
// addition is associative, so the parallel scan is bit-identical to a sequential one and the face
// chosen for a given draw is defined independently of the scan shape.  The last valid face absorbs
// the rounding residual (the analogue of :36-37).
//
// The reference's own draws (Julia global RNG + alias tables) cannot be reproduced by anyone; with
// inj_face / inj_r1 / inj_r2 the draws are injected and the remaining arithmetic (:60-82) is
// bit-identical to the reference:  u = sqrt(r1); w1 = 1-u; w2 = u*(1-v); w3 = u*v;
// p = ((w1*v1) + (w2*v2)) + (w3*v3).
#include <algorithm>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kST = 256;                       // threads per CTA
constexpr size_t kMaxSmemCdf = 200 * 1024;     // fused path while the CDF fits in shared memory

__device__ __forceinline__ void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

struct SampleParams {
    const float* verts;      // [Nmesh][Vmax][3]
    const int32_t* faces;    // [Nmesh][Fmax][3] local ids
    const int32_t* faces_len;
    int Vmax, Fmax, S;
    double eps;
    unsigned long long seed, offset;
    const unsigned long long* offset_dev;   // optional: added to `offset` at run time (a captured graph draws afresh at every replay)
    const int32_t* inj_face;
    const float* inj_r1;
    const float* inj_r2;
    float* samples;          // [Nmesh][S][3]
    int32_t* face_idx_out;   // [Nmesh][S] or null
    float* bary_out;         // [Nmesh][S][3] or null
    unsigned long long* cdf_ws;  // [Nmesh][Fmax] (two-pass path only)
    int samples_per_cta;
};

// compute_faces_areas (src/rep/mesh.jl:765-780) in the reference's operation order
__device__ __forceinline__ float face_area(const float* __restrict__ V, const int32_t* __restrict__ Fc, int f) {
    const float* v1 = V + 3 * (size_t)__ldg(Fc + 3 * (size_t)f);
    const float* v2 = V + 3 * (size_t)__ldg(Fc + 3 * (size_t)f + 1);
    const float* v3 = V + 3 * (size_t)__ldg(Fc + 3 * (size_t)f + 2);
    const float x1 = __ldg(v1), y1 = __ldg(v1 + 1), z1 = __ldg(v1 + 2);
    const float ax = __fsub_rn(__ldg(v2), x1), ay = __fsub_rn(__ldg(v2 + 1), y1), az = __fsub_rn(__ldg(v2 + 2), z1);
    const float bx = __fsub_rn(__ldg(v3), x1), by = __fsub_rn(__ldg(v3 + 1), y1), bz = __fsub_rn(__ldg(v3 + 2), z1);
    const float cx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    const float cy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    const float cz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
    const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    return __fdiv_rn(n, 2.0f);
}

// CTA-wide: cdf[f] = Σ_{g<=f} floor(area_g / max(Σ area, eps) * 2^53), f < nF.
__device__ void build_cdf(const float* __restrict__ V, const int32_t* __restrict__ Fc, int nF, double eps,
                          unsigned long long* cdf) {
    __shared__ double s_d[kST / 32];
    __shared__ unsigned long long s_u[kST / 32];
    __shared__ double s_tot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 1. areas (kept as double bit patterns in the cdf array) and their Float64 sum.  A double holds any
    //    sum of <2^29 binary32 values of comparable exponent exactly, so the order of this reduction does
    //    not show in the result for real meshes.
    double mine = 0.0;
#pragma unroll 4
    for (int f = tid; f < nF; f += kST) {
        const double a = (double)face_area(V, Fc, f);
        cdf[f] = (unsigned long long)__double_as_longlong(a);
        mine += a;
    }
    mine = warp_sum(mine);
    if (lane == 0) s_d[warp] = mine;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kST / 32; ++w) t += s_d[w];
        s_tot = t;
    }
    __syncthreads();
    const double den = fmax(s_tot, eps);  // max.(sum(...; dims=2), eps)   mesh_func.jl:35
    // 2. fixed-point probabilities + inclusive scan (contiguous chunk per thread, then scan of chunk sums)
    const int per = (nF + kST - 1) / kST;
    const int f0 = min(tid * per, nF), f1 = min(f0 + per, nF);
    unsigned long long local = 0;
    for (int f = f0; f < f1; ++f) {
        const double p = __ddiv_rn(__longlong_as_double((long long)cdf[f]), den);
        const unsigned long long q = (unsigned long long)(p * 9007199254740992.0);  // floor(p * 2^53), p <= 1
        local += q;
        cdf[f] = local;
    }
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_u[warp] = incl;
    __syncthreads();
    unsigned long long base = incl - local;
    for (int w = 0; w < warp; ++w) base += s_u[w];
    for (int f = f0; f < f1; ++f) cdf[f] += base;
    __syncthreads();
}

__device__ __forceinline__ void draw_and_emit(const SampleParams& p, int mesh, int nF, const float* __restrict__ V,
                                              const int32_t* __restrict__ Fc,
                                              const unsigned long long* __restrict__ cdf, int s) {
    const size_t o = (size_t)mesh * p.S + s;
    int face;
    float r1, r2;
    if (p.inj_face) {
        face = __ldg(p.inj_face + o);
        r1 = __ldg(p.inj_r1 + o);
        r2 = __ldg(p.inj_r2 + o);
    } else {
        const unsigned long long off = p.offset + (p.offset_dev ? __ldg(p.offset_dev) : 0ull);
        unsigned c[4] = {(unsigned)s, (unsigned)mesh, (unsigned)off, (unsigned)(off >> 32)};
        philox4x32_10(c, (unsigned)p.seed, (unsigned)(p.seed >> 32));
        const unsigned long long m = (((unsigned long long)c[0] << 32) | c[1]) >> 11;  // 53-bit uniform
        r1 = (float)(c[2] >> 8) * (1.0f / 16777216.0f);
        r2 = (float)(c[3] >> 8) * (1.0f / 16777216.0f);
        int lo = 0, hi = nF - 1;  // smallest f with cdf[f] > m, clamped to the last valid face
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] > m) hi = mid; else lo = mid + 1;
        }
        face = lo;
    }
    float* out = p.samples + 3 * o;
    if (nF <= 0 || face < 0 || face >= nF) {  // empty mesh / bad injected id: defined output, flagged by -1
        out[0] = out[1] = out[2] = 0.0f;
        if (p.face_idx_out) p.face_idx_out[o] = -1;
        if (p.bary_out) { p.bary_out[3 * o] = 0.f; p.bary_out[3 * o + 1] = 0.f; p.bary_out[3 * o + 2] = 0.f; }
        return;
    }
    const float* v1 = V + 3 * (size_t)__ldg(Fc + 3 * (size_t)face);
    const float* v2 = V + 3 * (size_t)__ldg(Fc + 3 * (size_t)face + 1);
    const float* v3 = V + 3 * (size_t)__ldg(Fc + 3 * (size_t)face + 2);
    const float u = __fsqrt_rn(r1), v = r2;               // mesh_func.jl:76-77
    const float w1 = __fsub_rn(1.0f, u);                  // :78
    const float w2 = __fmul_rn(u, __fsub_rn(1.0f, v));    // :79
    const float w3 = __fmul_rn(u, v);                     // :80
#pragma unroll
    for (int d = 0; d < 3; ++d)                           // :71
        out[d] = __fadd_rn(__fadd_rn(__fmul_rn(w1, __ldg(v1 + d)), __fmul_rn(w2, __ldg(v2 + d))), __fmul_rn(w3, __ldg(v3 + d)));
    if (p.face_idx_out) p.face_idx_out[o] = face;
    if (p.bary_out) { p.bary_out[3 * o] = w1; p.bary_out[3 * o + 1] = w2; p.bary_out[3 * o + 2] = w3; }
}

// fused: grid (chunks, Nmesh); dynamic smem = 8*Fmax bytes (0 when the draws are injected)
__global__ void __launch_bounds__(kST) sample_points_fused_kernel(SampleParams p) {
    extern __shared__ __align__(16) unsigned long long s_cdf[];
    const int mesh = blockIdx.y;
    const int nF = min(__ldg(p.faces_len + mesh), p.Fmax);
    const float* V = p.verts + (size_t)mesh * p.Vmax * 3;
    const int32_t* Fc = p.faces + (size_t)mesh * p.Fmax * 3;
    if (!p.inj_face && nF > 0) build_cdf(V, Fc, nF, p.eps, s_cdf);
    const int s0 = blockIdx.x * p.samples_per_cta, s1 = min(s0 + p.samples_per_cta, p.S);
    for (int s = s0 + threadIdx.x; s < s1; s += kST) draw_and_emit(p, mesh, nF, V, Fc, s_cdf, s);
}

// two-pass path for meshes whose CDF does not fit in shared memory
__global__ void __launch_bounds__(kST) sample_points_cdf_kernel(SampleParams p) {
    const int mesh = blockIdx.x;
    const int nF = min(__ldg(p.faces_len + mesh), p.Fmax);
    if (nF > 0)
        build_cdf(p.verts + (size_t)mesh * p.Vmax * 3, p.faces + (size_t)mesh * p.Fmax * 3, nF, p.eps,
                  p.cdf_ws + (size_t)mesh * p.Fmax);
}
__global__ void __launch_bounds__(kST) sample_points_draw_kernel(SampleParams p) {
    const int mesh = blockIdx.y;
    const int nF = min(__ldg(p.faces_len + mesh), p.Fmax);
    const float* V = p.verts + (size_t)mesh * p.Vmax * 3;
    const int32_t* Fc = p.faces + (size_t)mesh * p.Fmax * 3;
    const int s0 = blockIdx.x * p.samples_per_cta, s1 = min(s0 + p.samples_per_cta, p.S);
    for (int s = s0 + threadIdx.x; s < s1; s += kST) draw_and_emit(p, mesh, nF, V, Fc, p.cdf_ws + (size_t)mesh * p.Fmax, s);
}

// pullback: one thread per (sample, corner-coordinate triple)
__global__ void __launch_bounds__(kST) sample_points_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ fidx,
                                                                const float* __restrict__ bary, const int32_t* __restrict__ faces,
                                                                int Vmax, int Fmax, int S, long total, float* __restrict__ gv) {
    const long o = (long)blockIdx.x * kST + threadIdx.x;
    if (o >= total) return;
    const int mesh = (int)(o / S);
    const int face = __ldg(fidx + o);
    if (face < 0 || face >= Fmax) return;
    const float gx = __ldg(g + 3 * o), gy = __ldg(g + 3 * o + 1), gz = __ldg(g + 3 * o + 2);
    const int32_t* fc = faces + ((size_t)mesh * Fmax + face) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int v = __ldg(fc + k);
        if (v < 0 || v >= Vmax) continue;
        const float w = __ldg(bary + 3 * o + k);
        float* dst = gv + ((size_t)mesh * Vmax + v) * 3;
        atomicAdd(dst, w * gx);
        atomicAdd(dst + 1, w * gy);
        atomicAdd(dst + 2, w * gz);
    }
}

}  // namespace
}  // namespace f3d

using namespace f3d;

extern "C" size_t f3d_sample_points_workspace_bytes(int32_t Nmesh, int32_t Fmax) {
    if (Nmesh <= 0 || Fmax <= 0) return 0;
    if (sizeof(unsigned long long) * (size_t)Fmax <= kMaxSmemCdf) return 0;  // fused path: CDF lives in shared memory
    return align_up(sizeof(unsigned long long) * (size_t)Nmesh * Fmax, 256);
}

namespace f3d {
namespace {
__global__ void sample_counter_bump_kernel(unsigned long long* ctr) { *ctr += 1ull; }
}  // namespace
}  // namespace f3d

extern "C" int32_t f3d_sample_points(const float* verts_padded, const int32_t* faces_padded, const int32_t* verts_len,
                                     const int32_t* faces_len, int32_t Nmesh, int32_t Vmax, int32_t Fmax, int32_t S,
                                     double eps, uint64_t seed, uint64_t offset, const int32_t* inj_face,
                                     const float* inj_r1, const float* inj_r2, float* samples, int32_t* face_idx_out,
                                     float* bary_out, void* ws, size_t ws_bytes, f3d_stream_t stream_) {
    return f3d_sample_points_replayable(verts_padded, faces_padded, verts_len, faces_len, Nmesh, Vmax, Fmax, S, eps, seed, offset, nullptr,
                                        inj_face, inj_r1, inj_r2, samples, face_idx_out, bary_out, ws, ws_bytes, stream_);
}

extern "C" int32_t f3d_sample_points_replayable(const float* verts_padded, const int32_t* faces_padded, const int32_t* verts_len,
                                                const int32_t* faces_len, int32_t Nmesh, int32_t Vmax, int32_t Fmax, int32_t S,
                                                double eps, uint64_t seed, uint64_t offset, uint64_t* offset_dev,
                                                const int32_t* inj_face, const float* inj_r1, const float* inj_r2, float* samples,
                                                int32_t* face_idx_out, float* bary_out, void* ws, size_t ws_bytes, f3d_stream_t stream_) {
    using namespace f3d;
    (void)verts_len;  // faces only reference valid vertices; kept in the ABI to mirror verts[:, 1:_verts_len[i], i] (:52)
    if (!verts_padded || !faces_padded || !faces_len || !samples) return fail(F3D_ERR_INVALID, "f3d_sample_points: null pointer");
    if (Nmesh <= 0 || Vmax <= 0 || Fmax <= 0 || S <= 0) return fail(F3D_ERR_INVALID, "f3d_sample_points: Nmesh, Vmax, Fmax, S must be positive (got %d, %d, %d, %d)", Nmesh, Vmax, Fmax, S);
    if (Nmesh > 65535) return fail(F3D_ERR_INVALID, "f3d_sample_points: Nmesh must be <= 65535 per call");
    const bool inj = inj_face != nullptr;
    if (inj && (!inj_r1 || !inj_r2)) return fail(F3D_ERR_INVALID, "f3d_sample_points: inj_face given without inj_r1/inj_r2");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SampleParams p;
    p.verts = verts_padded; p.faces = faces_padded; p.faces_len = faces_len;
    p.Vmax = Vmax; p.Fmax = Fmax; p.S = S; p.eps = eps; p.seed = seed; p.offset = offset; p.offset_dev = reinterpret_cast<const unsigned long long*>(offset_dev);
    p.inj_face = inj_face; p.inj_r1 = inj_r1; p.inj_r2 = inj_r2;
    p.samples = samples; p.face_idx_out = face_idx_out; p.bary_out = bary_out; p.cdf_ws = nullptr;
    // one sample per thread: the draw path is a chain of dependent loads (CDF search -> face ids -> vertices), so
    // the latency is paid once per CTA, not once per sample; rebuilding the CDF per CTA costs nF/256 areas per thread
    int chunks = std::min((S + kST - 1) / kST, 65535);
    p.samples_per_cta = (S + chunks - 1) / chunks;
    chunks = (S + p.samples_per_cta - 1) / p.samples_per_cta;
    const size_t cdf_bytes = sizeof(unsigned long long) * (size_t)Fmax;
    if (inj || cdf_bytes <= kMaxSmemCdf) {
        const size_t smem = inj ? 0 : cdf_bytes;
        if (smem > 48 * 1024) F3D_CUDA(cudaFuncSetAttribute(sample_points_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_points_fused_kernel<<<dim3(chunks, Nmesh), kST, smem, stream>>>(p);
        F3D_CHECK_LAUNCH("sample_points_fused_kernel");
    } else {
        const size_t need = align_up(sizeof(unsigned long long) * (size_t)Nmesh * Fmax, 256);
        if (!ws || ws_bytes < need) return fail(F3D_ERR_WORKSPACE, "f3d_sample_points: workspace %zu < required %zu bytes", ws_bytes, need);
        p.cdf_ws = static_cast<unsigned long long*>(ws);
        sample_points_cdf_kernel<<<Nmesh, kST, 0, stream>>>(p);
        F3D_CHECK_LAUNCH("sample_points_cdf_kernel");
        sample_points_draw_kernel<<<dim3(chunks, Nmesh), kST, 0, stream>>>(p);
        F3D_CHECK_LAUNCH("sample_points_draw_kernel");
    }
    if (offset_dev && !inj) {   // the next launch (or the next replay of a captured graph) draws from a fresh counter block
        sample_counter_bump_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<unsigned long long*>(offset_dev));
        F3D_CHECK_LAUNCH("sample_counter_bump_kernel");
    }
    return F3D_OK;
}

extern "C" int32_t f3d_sample_points_bwd(const float* gsamples, const int32_t* face_idx, const float* bary,
                                         const int32_t* faces_padded, int32_t Nmesh, int32_t Vmax, int32_t Fmax, int32_t S,
                                         float* gverts_padded, f3d_stream_t stream_) {
    if (!gsamples || !face_idx || !bary || !faces_padded || !gverts_padded) return fail(F3D_ERR_INVALID, "f3d_sample_points_bwd: null pointer");
    if (Nmesh <= 0 || Vmax <= 0 || Fmax <= 0 || S <= 0) return fail(F3D_ERR_INVALID, "f3d_sample_points_bwd: Nmesh, Vmax, Fmax, S must be positive (got %d, %d, %d, %d)", Nmesh, Vmax, Fmax, S);
    const long total = (long)Nmesh * S;
    sample_points_bwd_kernel<<<(unsigned)((total + kST - 1) / kST), kST, 0, static_cast<cudaStream_t>(stream_)>>>(
        gsamples, face_idx, bary, faces_padded, Vmax, Fmax, S, total, gverts_padded);
    F3D_CHECK_LAUNCH("sample_points_bwd_kernel");
    return F3D_OK;
}
