// mesh.cu — TriMesh kernels on packed verts/faces for sm_100a, plus the host-side topology build.
//
// Replaces Flux3D.jl src/rep/mesh.jl:589-618 (compute_verts_normals_packed), :689-700
// (compute_faces_normals_packed), :765-780 (compute_faces_areas_packed), :907-1002 (edges /
// Laplacian, built once per topology and cached), src/metrics/mesh.jl:9-15 (laplacian_loss — the
// reference copies verts to the host and runs a CPU SpMM every call) and :24-32 (edge_loss).
//
// All of these are tiny, HBM/latency-bound gathers (cfg4: < 1 MB per call): the design rule is ONE
// launch per op, no host round trip, no float atomics (gather formulations over cached CSR
// adjacency => run-to-run deterministic), scalar losses reduced in-kernel by a last-block pass.
#include <algorithm>
#include <vector>

#include "f3d_common.cuh"

namespace f3d {
namespace {

constexpr int kMT = 256;  // threads per CTA for the mesh kernels

// _lg_cross — src/rep/utils.jl:4-21 (no contraction: every product and difference rounded)
__device__ __forceinline__ void cross3(float ax, float ay, float az, float bx, float by, float bz, float& cx,
                                       float& cy, float& cz) {
    cx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    cy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    cz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
}
// _norm(dims=1) of a 3-vector — src/rep/utils.jl:29: sqrt((x²+y²)+z²)
__device__ __forceinline__ float norm3(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// corner cross product of face f at slot k: (v[k+1]-v[k]) x (v[k+2]-v[k])   mesh.jl:604-615
__device__ __forceinline__ void corner_cross(const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                             int f, int k, float& cx, float& cy, float& cz) {
    const int ka = (k + 1) % 3, kb = (k + 2) % 3;
    const float* v0 = verts + 3 * (size_t)__ldg(faces + 3 * (size_t)f + k);
    const float* va = verts + 3 * (size_t)__ldg(faces + 3 * (size_t)f + ka);
    const float* vb = verts + 3 * (size_t)__ldg(faces + 3 * (size_t)f + kb);
    const float x0 = __ldg(v0), y0 = __ldg(v0 + 1), z0 = __ldg(v0 + 2);
    cross3(__fsub_rn(__ldg(va), x0), __fsub_rn(__ldg(va + 1), y0), __fsub_rn(__ldg(va + 2), z0),
           __fsub_rn(__ldg(vb), x0), __fsub_rn(__ldg(vb + 1), y0), __fsub_rn(__ldg(vb + 2), z0), cx, cy, cz);
}

__global__ void __launch_bounds__(kMT) faces_areas_normals_kernel(const float* __restrict__ verts,
                                                                  const int32_t* __restrict__ faces, int nF,
                                                                  float* __restrict__ areas,
                                                                  float* __restrict__ normals) {
    const int f = blockIdx.x * kMT + threadIdx.x;
    if (f >= nF) return;
    float cx, cy, cz;
    corner_cross(verts, faces, f, 0, cx, cy, cz);
    const float n = norm3(cx, cy, cz);
    if (areas) areas[f] = __fdiv_rn(n, 2.0f);  // mesh.jl:777-779
    if (normals) {                              // _normalize: A ./ max(norm, 1e-6)   utils.jl:23-27
        const float m = fmaxf(n, 1e-6f);
        normals[3 * (size_t)f + 0] = __fdiv_rn(cx, m);
        normals[3 * (size_t)f + 1] = __fdiv_rn(cy, m);
        normals[3 * (size_t)f + 2] = __fdiv_rn(cz, m);
    }
}

// One thread per vertex, walking its incident corners (CSR v2c, ordered by (slot, face)).
__global__ void __launch_bounds__(kMT) verts_normals_kernel(const float* __restrict__ verts,
                                                            const int32_t* __restrict__ faces,
                                                            const int32_t* __restrict__ v2c_rowptr,
                                                            const int32_t* __restrict__ v2c, int nV, int mode,
                                                            float* __restrict__ out) {
    const int v = blockIdx.x * kMT + threadIdx.x;
    if (v >= nV) return;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    const int p0 = __ldg(v2c_rowptr + v), p1 = __ldg(v2c_rowptr + v + 1);
    for (int p = p0; p < p1; ++p) {
        const int c = __ldg(v2c + p);
        const int f = c / 3, k = c - 3 * f;
        if (mode == F3D_NORMALS_REFERENCE_CPU && p + 1 < p1) {
            // gather-add-ASSIGN on a Zygote.Buffer (mesh.jl:604-615): within one slot only the last
            // (highest) face survives — skip every corner that is followed by one of the same slot.
            const int cn = __ldg(v2c + p + 1);
            if (cn - 3 * (cn / 3) == k) continue;
        }
        float cx, cy, cz;
        corner_cross(verts, faces, f, k, cx, cy, cz);
        ax = __fadd_rn(ax, cx);
        ay = __fadd_rn(ay, cy);
        az = __fadd_rn(az, cz);
    }
    const float m = fmaxf(norm3(ax, ay, az), 1e-6f);
    out[3 * (size_t)v + 0] = __fdiv_rn(ax, m);
    out[3 * (size_t)v + 1] = __fdiv_rn(ay, m);
    out[3 * (size_t)v + 2] = __fdiv_rn(az, m);
}

// ---- scalar-loss reduction shared by laplacian_loss / edge_loss -------------------------------------
// Block partials in double (fixed tree), last block sums them in a fixed order: deterministic.
struct ReduceWs {
    double* partial;    // [gridDim.x]
    unsigned* counter;  // zero on entry; reset to zero by the last block (so the workspace is reusable)
};

__device__ __forceinline__ void block_reduce_finish(double mine, ReduceWs ws, double denom, float* loss) {
    __shared__ double s_w[kMT / 32];
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    mine = warp_sum(mine);
    if ((tid & 31) == 0) s_w[tid >> 5] = mine;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kMT / 32; ++w) s += s_w[w];
        ws.partial[blockIdx.x] = s;
        __threadfence();
        s_last = (atomicAdd(ws.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double t = 0.0;
    for (int k = tid; k < (int)gridDim.x; k += kMT) t += __ldcg(ws.partial + k);
    t = warp_sum(t);
    __syncthreads();
    if ((tid & 31) == 0) s_w[tid >> 5] = t;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kMT / 32; ++w) s += s_w[w];
        loss[0] = (float)(s / denom);
        *ws.counter = 0u;
    }
}

// (L v)_i with the reference's SpMM order: ascending columns, C += L[i,j]*v_j, no contraction.
__device__ __forceinline__ void lap_row(const float* __restrict__ verts, const int32_t* __restrict__ rowptr,
                                        const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                        int i, float& ax, float& ay, float& az) {
    ax = ay = az = 0.0f;
    const int p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
    for (int p = p0; p < p1; ++p) {
        const float w = __ldg(vals + p);
        const float* x = verts + 3 * (size_t)__ldg(colidx + p);
        ax = __fadd_rn(ax, __fmul_rn(w, __ldg(x)));
        ay = __fadd_rn(ay, __fmul_rn(w, __ldg(x + 1)));
        az = __fadd_rn(az, __fmul_rn(w, __ldg(x + 2)));
    }
}

__global__ void __launch_bounds__(kMT) laplacian_loss_kernel(const float* __restrict__ verts,
                                                             const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const float* __restrict__ vals, int nV, double denom,
                                                             ReduceWs ws, float* loss) {
    const int i = blockIdx.x * kMT + threadIdx.x;
    double mine = 0.0;
    if (i < nV) {
        float ax, ay, az;
        lap_row(verts, rowptr, colidx, vals, i, ax, ay, az);
        mine = (double)norm3(ax, ay, az);  // _norm(L; dims=2)   metrics/mesh.jl:13
    }
    block_reduce_finish(mine, ws, denom, loss);
}

// backward, pass 1: unit residual directions  n̂_i = (Lv)_i / ‖(Lv)_i‖  (0 where the norm is 0)
__global__ void __launch_bounds__(kMT) laplacian_dir_kernel(const float* __restrict__ verts,
                                                            const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx,
                                                            const float* __restrict__ vals, int nV,
                                                            float* __restrict__ dir) {
    const int i = blockIdx.x * kMT + threadIdx.x;
    if (i >= nV) return;
    float ax, ay, az;
    lap_row(verts, rowptr, colidx, vals, i, ax, ay, az);
    const float n = norm3(ax, ay, az);
    const float inv = n > 0.0f ? 1.0f / n : 0.0f;
    dir[3 * (size_t)i + 0] = ax * inv;
    dir[3 * (size_t)i + 1] = ay * inv;
    dir[3 * (size_t)i + 2] = az * inv;
}

// backward, pass 2: g_j = gout/nV_total * Σ_i L[i,j] n̂_i.  L's pattern is symmetric, so row j's
// columns are exactly the rows i with L[i,j] != 0; L[i,j] = 1/deg(i) for i != j, -1 for i == j.
__global__ void __launch_bounds__(kMT) laplacian_bwd_kernel(const float* __restrict__ dir,
                                                            const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx, int nV,
                                                            float scale_inv_n, const float* __restrict__ gout,
                                                            float* __restrict__ gverts) {
    const int j = blockIdx.x * kMT + threadIdx.x;
    if (j >= nV) return;
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;
    const int p0 = __ldg(rowptr + j), p1 = __ldg(rowptr + j + 1);
    for (int p = p0; p < p1; ++p) {
        const int i = __ldg(colidx + p);
        float w;
        if (i == j) {
            w = -1.0f;
        } else {
            const int deg = __ldg(rowptr + i + 1) - __ldg(rowptr + i) - 1;
            w = (float)(1.0 / (double)deg);  // same value as vals of row i   mesh.jl:985-986
        }
        gx += w * __ldg(dir + 3 * (size_t)i);
        gy += w * __ldg(dir + 3 * (size_t)i + 1);
        gz += w * __ldg(dir + 3 * (size_t)i + 2);
    }
    const float s = __ldg(gout) * scale_inv_n;
    gverts[3 * (size_t)j + 0] = s * gx;
    gverts[3 * (size_t)j + 1] = s * gy;
    gverts[3 * (size_t)j + 2] = s * gz;
}

__global__ void __launch_bounds__(kMT) edge_loss_kernel(const float* __restrict__ verts,
                                                        const int32_t* __restrict__ edges, int nE, float target,
                                                        double denom, ReduceWs ws, float* loss) {
    const int e = blockIdx.x * kMT + threadIdx.x;
    double mine = 0.0;
    if (e < nE) {
        const float* a = verts + 3 * (size_t)__ldg(edges + 2 * (size_t)e);
        const float* b = verts + 3 * (size_t)__ldg(edges + 2 * (size_t)e + 1);
        const float n = norm3(__fsub_rn(__ldg(a), __ldg(b)), __fsub_rn(__ldg(a + 1), __ldg(b + 1)),
                              __fsub_rn(__ldg(a + 2), __ldg(b + 2)));
        const float t = __fsub_rn(n, target);
        mine = (double)__fmul_rn(t, t);  // metrics/mesh.jl:29
    }
    block_reduce_finish(mine, ws, denom, loss);
}

// edge_loss pullback, gathered per vertex over its neighbours (the Laplacian's off-diagonal columns)
__global__ void __launch_bounds__(kMT) edge_loss_bwd_kernel(const float* __restrict__ verts, const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx, int nV, float scale, float target,
                                                            const float* __restrict__ gout, float* __restrict__ gverts) {
    const int i = blockIdx.x * kMT + threadIdx.x;
    if (i >= nV) return;
    const float xi = __ldg(verts + 3 * (size_t)i), yi = __ldg(verts + 3 * (size_t)i + 1), zi = __ldg(verts + 3 * (size_t)i + 2);
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;
    const int p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
    for (int p = p0; p < p1; ++p) {
        const int j = __ldg(colidx + p);
        if (j == i) continue;
        const float dx = xi - __ldg(verts + 3 * (size_t)j), dy = yi - __ldg(verts + 3 * (size_t)j + 1), dz = zi - __ldg(verts + 3 * (size_t)j + 2);
        const float n = norm3(dx, dy, dz);
        const float c = n > 0.0f ? (n - target) / n : 0.0f;
        gx += c * dx; gy += c * dy; gz += c * dz;
    }
    const float s = __ldg(gout) * scale;
    gverts[3 * (size_t)i] = s * gx; gverts[3 * (size_t)i + 1] = s * gy; gverts[3 * (size_t)i + 2] = s * gz;
}

// ---- packed <-> padded (src/rep/utils.jl:131-185): 4-byte elements, so Float32 verts and Int32 faces share the code ----
// packed [ΣL][D], item i = rows offsets[i] .. offsets[i+1]-1;  padded [N][W][D], rows past an item's length hold `fill`.
// delta (optional, [N]): added to every real element on the way to packed, subtracted on the way to padded — the
// global <-> local vertex ids of packed / padded faces (src/rep/mesh.jl:884-896).
__global__ void __launch_bounds__(kMT) packed_to_padded_kernel(const unsigned* __restrict__ packed, const int* __restrict__ offsets,
                                                               const int* __restrict__ delta, int N, int W, int D, unsigned fill,
                                                               unsigned* __restrict__ padded) {
    const size_t e = (size_t)blockIdx.x * kMT + threadIdx.x;  // flat index into padded
    if (e >= (size_t)N * W * D) return;
    const int d = (int)(e % D), w = (int)((e / D) % W), n = (int)(e / ((size_t)D * W));
    const int o0 = __ldg(offsets + n), len = __ldg(offsets + n + 1) - o0;
    unsigned v = fill;
    if (w < len) {
        v = __ldg(packed + (size_t)(o0 + w) * D + d);
        if (delta) v = (unsigned)((int)v - __ldg(delta + n));
    }
    padded[e] = v;
}

__global__ void __launch_bounds__(kMT) padded_to_packed_kernel(const unsigned* __restrict__ padded, const int* __restrict__ offsets,
                                                               const int* __restrict__ delta, int N, int W, int D, int total_rows,
                                                               unsigned* __restrict__ packed) {
    const size_t e = (size_t)blockIdx.x * kMT + threadIdx.x;  // flat index into packed
    if (e >= (size_t)total_rows * D) return;
    const int d = (int)(e % D), r = (int)(e / D);
    int lo = 0, hi = N;  // the item that owns packed row r: largest n with offsets[n] <= r (empty items are skipped)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= r) lo = mid; else hi = mid;
    }
    unsigned v = __ldg(padded + ((size_t)lo * W + (r - __ldg(offsets + lo))) * D + d);
    if (delta) v = (unsigned)((int)v + __ldg(delta + lo));
    packed[e] = v;
}

size_t reduce_ws_bytes(int n) {
    const int blocks = (n + kMT - 1) / kMT;
    return align_up(sizeof(double) * (size_t)blocks, 256) + 256;
}
// Counter lives in the last 256 bytes; it must be zero before the first use (the caller's workspace is
// zero-initialised once by f3d_*_workspace users via cudaMemsetAsync here, cheap and stream-ordered).
ReduceWs reduce_ws(void* ws, int n) {
    const int blocks = (n + kMT - 1) / kMT;
    ReduceWs r;
    r.partial = static_cast<double*>(ws);
    r.counter = reinterpret_cast<unsigned*>(static_cast<unsigned char*>(ws) + align_up(sizeof(double) * (size_t)blocks, 256));
    return r;
}

}  // namespace
}  // namespace f3d

using namespace f3d;

extern "C" int32_t f3d_faces_areas_normals(const float* verts, const int32_t* faces, int32_t nV, int32_t nF,
                                           float* areas, float* normals, f3d_stream_t stream) {
    if (!verts || !faces) return fail(F3D_ERR_INVALID, "f3d_faces_areas_normals: null verts/faces pointer");
    if (nV <= 0 || nF < 0) return fail(F3D_ERR_INVALID, "f3d_faces_areas_normals: bad sizes nV=%d nF=%d", nV, nF);
    if (nF == 0 || (!areas && !normals)) return F3D_OK;
    faces_areas_normals_kernel<<<(nF + kMT - 1) / kMT, kMT, 0, static_cast<cudaStream_t>(stream)>>>(verts, faces, nF, areas, normals);
    F3D_CHECK_LAUNCH("faces_areas_normals_kernel");
    return F3D_OK;
}

extern "C" int32_t f3d_verts_normals(const float* verts, const int32_t* faces, const int32_t* v2c_rowptr,
                                     const int32_t* v2c, int32_t nV, int32_t nF, int32_t mode, float* out,
                                     f3d_stream_t stream) {
    if (!verts || !faces || !v2c_rowptr || !v2c || !out) return fail(F3D_ERR_INVALID, "f3d_verts_normals: null pointer");
    if (nV <= 0 || nF < 0) return fail(F3D_ERR_INVALID, "f3d_verts_normals: bad sizes nV=%d nF=%d", nV, nF);
    if (mode != F3D_NORMALS_REFERENCE_CPU && mode != F3D_NORMALS_ACCUMULATE) return fail(F3D_ERR_INVALID, "f3d_verts_normals: unknown mode %d", mode);
    verts_normals_kernel<<<(nV + kMT - 1) / kMT, kMT, 0, static_cast<cudaStream_t>(stream)>>>(verts, faces, v2c_rowptr, v2c, nV, mode, out);
    F3D_CHECK_LAUNCH("verts_normals_kernel");
    return F3D_OK;
}

extern "C" size_t f3d_laplacian_workspace_bytes(int32_t nV) {
    if (nV <= 0) return 0;
    return reduce_ws_bytes(nV) + align_up(sizeof(float) * 3 * (size_t)nV, 256);  // + n̂ for the backward
}

extern "C" int32_t f3d_laplacian_loss(const float* verts, const int32_t* rowptr, const int32_t* colidx,
                                      const float* vals, int32_t nV, int32_t nV_total, float* loss_dev, void* ws,
                                      size_t ws_bytes, f3d_stream_t stream_) {
    if (!verts || !rowptr || !colidx || !vals || !loss_dev) return fail(F3D_ERR_INVALID, "f3d_laplacian_loss: null pointer");
    if (nV <= 0) return fail(F3D_ERR_INVALID, "f3d_laplacian_loss: nV must be positive (got %d)", nV);
    if (nV_total == 0) nV_total = nV;
    if (nV_total < nV) return fail(F3D_ERR_INVALID, "f3d_laplacian_loss: nV_total (%d) < nV (%d)", nV_total, nV);
    if (!ws || ws_bytes < reduce_ws_bytes(nV)) return fail(F3D_ERR_WORKSPACE, "f3d_laplacian_loss: workspace %zu < required %zu bytes", ws_bytes, reduce_ws_bytes(nV));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ReduceWs r = reduce_ws(ws, nV);
    F3D_CUDA(cudaMemsetAsync(r.counter, 0, sizeof(unsigned), stream));
    laplacian_loss_kernel<<<(nV + kMT - 1) / kMT, kMT, 0, stream>>>(verts, rowptr, colidx, vals, nV, (double)nV_total, r, loss_dev);
    F3D_CHECK_LAUNCH("laplacian_loss_kernel");
    return F3D_OK;
}

extern "C" int32_t f3d_laplacian_loss_bwd(const float* verts, const int32_t* rowptr, const int32_t* colidx,
                                          const float* vals, int32_t nV, int32_t nV_total, const float* gout_dev,
                                          float* gverts, void* ws, size_t ws_bytes, f3d_stream_t stream_) {
    if (!verts || !rowptr || !colidx || !vals || !gout_dev || !gverts) return fail(F3D_ERR_INVALID, "f3d_laplacian_loss_bwd: null pointer");
    if (nV <= 0) return fail(F3D_ERR_INVALID, "f3d_laplacian_loss_bwd: nV must be positive (got %d)", nV);
    if (nV_total == 0) nV_total = nV;
    if (!ws || ws_bytes < f3d_laplacian_workspace_bytes(nV)) return fail(F3D_ERR_WORKSPACE, "f3d_laplacian_loss_bwd: workspace %zu < required %zu bytes", ws_bytes, f3d_laplacian_workspace_bytes(nV));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    float* dir = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + reduce_ws_bytes(nV));
    const int grid = (nV + kMT - 1) / kMT;
    laplacian_dir_kernel<<<grid, kMT, 0, stream>>>(verts, rowptr, colidx, vals, nV, dir);
    F3D_CHECK_LAUNCH("laplacian_dir_kernel");
    laplacian_bwd_kernel<<<grid, kMT, 0, stream>>>(dir, rowptr, colidx, nV, 1.0f / (float)nV_total, gout_dev, gverts);
    F3D_CHECK_LAUNCH("laplacian_bwd_kernel");
    return F3D_OK;
}

extern "C" size_t f3d_edge_loss_workspace_bytes(int32_t nE) { return nE > 0 ? reduce_ws_bytes(nE) : 0; }

extern "C" int32_t f3d_edge_loss(const float* verts, const int32_t* edges, int32_t nE, int32_t nE_total, float target,
                                 float* loss_dev, void* ws, size_t ws_bytes, f3d_stream_t stream_) {
    if (!verts || !edges || !loss_dev) return fail(F3D_ERR_INVALID, "f3d_edge_loss: null pointer");
    if (nE <= 0) return fail(F3D_ERR_INVALID, "f3d_edge_loss: nE must be positive (got %d)", nE);
    if (nE_total == 0) nE_total = nE;
    if (!ws || ws_bytes < reduce_ws_bytes(nE)) return fail(F3D_ERR_WORKSPACE, "f3d_edge_loss: workspace %zu < required %zu bytes", ws_bytes, reduce_ws_bytes(nE));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ReduceWs r = reduce_ws(ws, nE);
    F3D_CUDA(cudaMemsetAsync(r.counter, 0, sizeof(unsigned), stream));
    edge_loss_kernel<<<(nE + kMT - 1) / kMT, kMT, 0, stream>>>(verts, edges, nE, target, (double)nE_total, r, loss_dev);
    F3D_CHECK_LAUNCH("edge_loss_kernel");
    return F3D_OK;
}

extern "C" int32_t f3d_edge_loss_bwd(const float* verts, const int32_t* rowptr, const int32_t* colidx, int32_t nV,
                                     int32_t nE_total, float target, const float* gout_dev, float* gverts,
                                     f3d_stream_t stream_) {
    if (!verts || !rowptr || !colidx || !gout_dev || !gverts) return fail(F3D_ERR_INVALID, "f3d_edge_loss_bwd: null pointer");
    if (nV <= 0 || nE_total <= 0) return fail(F3D_ERR_INVALID, "f3d_edge_loss_bwd: nV and nE_total must be positive (got %d, %d)", nV, nE_total);
    edge_loss_bwd_kernel<<<(nV + kMT - 1) / kMT, kMT, 0, static_cast<cudaStream_t>(stream_)>>>(
        verts, rowptr, colidx, nV, 2.0f / (float)nE_total, target, gout_dev, gverts);
    F3D_CHECK_LAUNCH("edge_loss_bwd_kernel");
    return F3D_OK;
}

// ---------------------------------------------------------------------------------------------------
// Host-side topology build (once per mesh topology; the reference caches the same products in the
// TriMesh struct, src/rep/mesh.jl:93-97).
//   edges   : unique (min,max) pairs, lexicographic order              _compute_edges_packed :907-955
//   f2e     : per face the edge ids of (v2,v3), (v3,v1), (v1,v2)       :943-949
//   Laplacian CSR: diagonal -1, off-diagonals Float32(1/deg), ascending columns   :957-1002
//   v2c     : per vertex its incident corners face*3+slot ordered by (slot, face)
// 64-bit edge keys: the reference hashes in the face index type and overflows UInt32 past 65535
// packed vertices (rep/mesh.jl:928-929); that overflow is deliberately not reproduced.
// ---------------------------------------------------------------------------------------------------
extern "C" int32_t f3d_mesh_topology_build_host(const int32_t* faces, int32_t nV, int32_t nF, int32_t* edges,
                                                int32_t* nE_out, int32_t* f2e, int32_t* lap_rowptr,
                                                int32_t* lap_colidx, float* lap_vals, int32_t* v2c_rowptr,
                                                int32_t* v2c) {
    if (!faces || !nE_out) return fail(F3D_ERR_INVALID, "f3d_mesh_topology_build_host: null faces/nE pointer");
    if (nV <= 0 || nF < 0) return fail(F3D_ERR_INVALID, "f3d_mesh_topology_build_host: bad sizes nV=%d nF=%d", nV, nF);
    for (int64_t t = 0; t < 3 * (int64_t)nF; ++t)
        if (faces[t] < 0 || faces[t] >= nV) return fail(F3D_ERR_INVALID, "f3d_mesh_topology_build_host: face index %d out of range [0,%d)", faces[t], nV);

    auto edge_key = [nV](int32_t a, int32_t b) -> uint64_t {
        const uint64_t lo = (uint64_t)std::min(a, b), hi = (uint64_t)std::max(a, b);
        return lo * (uint64_t)nV + hi;
    };
    std::vector<uint64_t> keys;
    keys.reserve(3 * (size_t)nF);
    for (int f = 0; f < nF; ++f) {
        const int32_t* t = faces + 3 * (size_t)f;
        keys.push_back(edge_key(t[0], t[1]));
        keys.push_back(edge_key(t[1], t[2]));
        keys.push_back(edge_key(t[2], t[0]));
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const int32_t nE = (int32_t)keys.size();
    *nE_out = nE;
    if (edges)
        for (int32_t e = 0; e < nE; ++e) {
            edges[2 * (size_t)e] = (int32_t)(keys[e] / (uint64_t)nV);
            edges[2 * (size_t)e + 1] = (int32_t)(keys[e] % (uint64_t)nV);
        }
    if (f2e)
        for (int f = 0; f < nF; ++f) {
            const int32_t* t = faces + 3 * (size_t)f;
            const uint64_t k3[3] = {edge_key(t[1], t[2]), edge_key(t[2], t[0]), edge_key(t[0], t[1])};
            for (int c = 0; c < 3; ++c)
                f2e[3 * (size_t)f + c] = (int32_t)(std::lower_bound(keys.begin(), keys.end(), k3[c]) - keys.begin());
        }
    if (lap_rowptr && lap_colidx && lap_vals) {
        std::vector<int32_t> deg(nV, 0);
        for (uint64_t k : keys) { deg[k / (uint64_t)nV]++; deg[k % (uint64_t)nV]++; }
        lap_rowptr[0] = 0;
        for (int v = 0; v < nV; ++v) lap_rowptr[v + 1] = lap_rowptr[v] + deg[v] + 1;
        std::vector<int32_t> fill(lap_rowptr, lap_rowptr + nV);
        // Visiting edges in lexicographic order makes every row ascending if the diagonal is slotted in
        // at the right moment: neighbours a < v arrive (as the 'hi' end) before neighbours b > v.
        // Simpler and obviously right: fill, then sort each (short) row.
        for (int v = 0; v < nV; ++v) lap_colidx[fill[v]++] = v;
        for (uint64_t k : keys) {
            const int32_t a = (int32_t)(k / (uint64_t)nV), b = (int32_t)(k % (uint64_t)nV);
            lap_colidx[fill[a]++] = b;
            lap_colidx[fill[b]++] = a;
        }
        for (int v = 0; v < nV; ++v) {
            std::sort(lap_colidx + lap_rowptr[v], lap_colidx + lap_rowptr[v + 1]);
            const float w = deg[v] > 0 ? (float)(1.0 / (double)deg[v]) : 0.0f;  // T.(1/deg) mesh.jl:985-986
            for (int p = lap_rowptr[v]; p < lap_rowptr[v + 1]; ++p) lap_vals[p] = (lap_colidx[p] == v) ? -1.0f : w;
        }
    }
    if (v2c_rowptr && v2c) {
        std::vector<int32_t> cnt(nV + 1, 0);
        for (int64_t t = 0; t < 3 * (int64_t)nF; ++t) cnt[faces[t] + 1]++;
        v2c_rowptr[0] = 0;
        for (int v = 0; v < nV; ++v) v2c_rowptr[v + 1] = v2c_rowptr[v] + cnt[v + 1];
        std::vector<int32_t> fill(v2c_rowptr, v2c_rowptr + nV);
        for (int k = 0; k < 3; ++k)          // slot-major, faces ascending inside a slot
            for (int f = 0; f < nF; ++f) v2c[fill[faces[3 * (size_t)f + k]]++] = 3 * f + k;
    }
    return F3D_OK;
}

extern "C" int32_t f3d_packed_to_padded(const void* packed, const int32_t* offsets, const int32_t* delta, int32_t N, int32_t W,
                                        int32_t D, uint32_t fill_bits, void* padded, f3d_stream_t stream) {
    using namespace f3d;
    if (!packed || !offsets || !padded) return fail(F3D_ERR_INVALID, "f3d_packed_to_padded: null pointer");
    if (N <= 0 || W <= 0 || D <= 0) return fail(F3D_ERR_INVALID, "f3d_packed_to_padded: N, W, D must be positive (got %d, %d, %d)", N, W, D);
    const size_t total = (size_t)N * W * D;
    packed_to_padded_kernel<<<(unsigned)((total + kMT - 1) / kMT), kMT, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned*>(packed), offsets, delta, N, W, D, fill_bits, static_cast<unsigned*>(padded));
    F3D_CHECK_LAUNCH("packed_to_padded_kernel");
    return F3D_OK;
}

extern "C" int32_t f3d_padded_to_packed(const void* padded, const int32_t* offsets, const int32_t* delta, int32_t N, int32_t W,
                                        int32_t D, int32_t total_rows, void* packed, f3d_stream_t stream) {
    using namespace f3d;
    if (!padded || !offsets || !packed) return fail(F3D_ERR_INVALID, "f3d_padded_to_packed: null pointer");
    if (N <= 0 || W <= 0 || D <= 0 || total_rows < 0) return fail(F3D_ERR_INVALID, "f3d_padded_to_packed: bad sizes (N %d, W %d, D %d, rows %d)", N, W, D, total_rows);
    if (total_rows == 0) return F3D_OK;
    const size_t total = (size_t)total_rows * D;
    padded_to_packed_kernel<<<(unsigned)((total + kMT - 1) / kMT), kMT, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned*>(padded), offsets, delta, N, W, D, total_rows, static_cast<unsigned*>(packed));
    F3D_CHECK_LAUNCH("padded_to_packed_kernel");
    return F3D_OK;
}
