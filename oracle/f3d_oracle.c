/*
 * f3d_oracle.c — CPU restatement of the Flux3D.jl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file's shared object.  The product (flux3d.jl_b200/) never links or calls it.
 *
 * Every function cites the reference file:line (relative to FluxML/Flux3D.jl @ v0.1.6) whose
 * arithmetic it restates.  The reference is pure Julia and cannot run in this image (no julia
 * binary), so this file is a *restatement*; it is pinned against the reference's own golden
 * vectors in tests/test_oracle_golden.py (test/rep.jl:178-388 normals/areas, test/rep.jl:112-175
 * edges/laplacian, test/metrics.jl:8-73 dense laplacian identity, test/metrics.jl:94-111
 * naive_chamfer identity, README.md:111-112 laplacian_loss(teapot) = 0.05888283f0).
 * kNN ordering and sample_points draws are NOT pinned by any reference test ("parity unpinned":
 * test/models.jl:24-41 asserts shapes only; test/transforms/mesh_func.jl:4-14 is statistical).
 *
 * Arithmetic policy (Julia never contracts a*b+c into an fma unless muladd/@fastmath is used,
 * and none is on this path): every +,-,*,/ and sqrt is a separately rounded IEEE binary32 (or
 * binary64 where the reference uses Float64) operation.  Build with -ffp-contract=off and
 * without -ffast-math (see oracle/Makefile).
 *
 * Layout convention: a Julia (3,N,B) Float32 array is byte-identical to C [B][N][3].
 * Indices are 0-based int32 at this boundary (the Julia side is 1-based).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_version(void) { return 1; }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Squared Euclidean distance as NearestNeighbors.jl / Distances.jl evaluate it for the
 * reference's KDTree(… ; Euclidean) (call sites src/metrics/pcloud.jl:57,64 and
 * src/models/dgcnn.jl:5-6): s = 0; for d in 1:F  s = s + abs2(a[d]-b[d]).
 * (sqrt is monotone and applied at the end by NearestNeighbors; ordering is decided on s.)
 * ---------------------------------------------------------------------------------------- */
static inline float sqdist(const float* a, const float* b, int F) {
    float s = 0.0f;
    for (int d = 0; d < F; ++d) {
        float t = a[d] - b[d];
        s = s + t * t;
    }
    return s;
}

static inline float sqdist3(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return ((dx * dx) + (dy * dy)) + (dz * dz);
}

/* Julia Base.mapreduce_impl(identity,+,A,ifirst,ilast,blksize=1024): pairwise summation that
 * `mean`/`sum` use on a dense Array{Float32}.  The <1024 leaf is
 *     v = A[ifirst] + A[ifirst+1];  @simd for i = ifirst+2:ilast  v += A[i]  end
 * and @simd licenses LLVM to re-associate: on x86-64 it keeps ORC_SIMD_LANES lane-strided partial
 * sums (8-wide vectors x 4 interleave on AVX2) and reduces them by a halving tree, the scalar
 * remainder is added last.  The lane count is the compiler's choice, so the reference value itself
 * is machine-dependent at the last ulps; every width from 4 to 32 reproduces the published
 * laplacian_loss(teapot) = 0.05888283f0 (README.md:111-112) exactly, the strictly sequential
 * order does not (0.058882844f0) — which is why the leaf is written this way. */
#define ORC_SIMD_LANES 32
static float pairwise_sum_f32(const float* a, long n) {
    if (n <= 0) return 0.0f;
    if (n == 1) return a[0];
    if (n - 1 < 1024) { /* ilast - ifirst < blksize */
        float v = a[0] + a[1];
        float lane[ORC_SIMD_LANES];
        long rest = n - 2, nv = rest / ORC_SIMD_LANES * ORC_SIMD_LANES;
        for (int l = 0; l < ORC_SIMD_LANES; ++l) lane[l] = 0.0f;
        for (long i = 0; i < nv; i += ORC_SIMD_LANES)
            for (int l = 0; l < ORC_SIMD_LANES; ++l) lane[l] = lane[l] + a[2 + i + l];
        for (int w = ORC_SIMD_LANES / 2; w >= 1; w /= 2)
            for (int l = 0; l < w; ++l) lane[l] = lane[l] + lane[l + w];
        v = v + lane[0];
        for (long i = nv; i < rest; ++i) v = v + a[2 + i];
        return v;
    }
    long left = ((n - 1) >> 1) + 1; /* imid = ifirst + ((ilast-ifirst) >> 1); left block = ifirst..imid */
    return pairwise_sum_f32(a, left) + pairwise_sum_f32(a + left, n - left);
}

/* ------------------------------------------------------------------------------------------
 * _nearest_neighbors(x::Array, y::Array)  — src/metrics/pcloud.jl:54-70 (the CPU method is the
 * reference semantics; an exact KD-tree 1-NN returns the metric argmin, so brute force gives
 * the same index except on exact ties, where the rule adopted is "lowest index").
 * x: [B][N][3], y: [B][M][3]; nn_x: [B][N] index into y, nn_y: [B][M] index into x.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_nearest_neighbors(const float* x, const float* y, int B, int N, int M,
                                   int32_t* nn_x, int32_t* nn_y) {
#pragma omp parallel for schedule(dynamic, 16) collapse(2)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < N; ++i) {
            const float* p = x + ((long)b * N + i) * 3;
            const float* q = y + (long)b * M * 3;
            float best = INFINITY;
            int bi = 0;
            for (int j = 0; j < M; ++j) {
                float d = sqdist3(p, q + (long)j * 3);
                if (d < best) { best = d; bi = j; }
            }
            nn_x[(long)b * N + i] = bi;
        }
    }
#pragma omp parallel for schedule(dynamic, 16) collapse(2)
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < M; ++j) {
            const float* p = y + ((long)b * M + j) * 3;
            const float* q = x + (long)b * N * 3;
            float best = INFINITY;
            int bi = 0;
            for (int i = 0; i < N; ++i) {
                /* same operand order as the x→y sweep: (x_i - y_j); squares make the sign irrelevant */
                float d = sqdist3(q + (long)i * 3, p);
                if (d < best) { best = d; bi = i; }
            }
            nn_y[(long)b * M + j] = bi;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * _chamfer_distance(A,B,w1,w2) — src/metrics/pcloud.jl:39-52
 *   dist_A_to_B = mean((A .- B[:, nn_for_A]) .^ 2) * 3f0      (mean over all 3*N*B elements)
 *   dist_B_to_A = mean((B .- A[:, nn_for_B]) .^ 2) * 3f0
 *   distance    = (w1*dist_A_to_B) + (w2*dist_B_to_A)
 * nnA/nnB may be NULL.  Returns the loss; also writes the two un-weighted terms if asked.
 * ---------------------------------------------------------------------------------------- */
ORC_API float orc_chamfer_distance(const float* A, const float* Bp, int B, int N, int M, float w1,
                                   float w2, int32_t* nnA_out, int32_t* nnB_out, float* terms) {
    int32_t* nnA = nnA_out ? nnA_out : (int32_t*)malloc(sizeof(int32_t) * (size_t)B * N);
    int32_t* nnB = nnB_out ? nnB_out : (int32_t*)malloc(sizeof(int32_t) * (size_t)B * M);
    orc_nearest_neighbors(A, Bp, B, N, M, nnA, nnB);

    long nA = (long)B * N * 3, nB = (long)B * M * 3;
    float* e = (float*)malloc(sizeof(float) * (size_t)(nA > nB ? nA : nB));
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            const float* p = A + ((long)b * N + i) * 3;
            const float* q = Bp + ((long)b * M + nnA[(long)b * N + i]) * 3;
            for (int d = 0; d < 3; ++d) {
                float t = p[d] - q[d];
                e[((long)b * N + i) * 3 + d] = t * t;
            }
        }
    float dAB = (pairwise_sum_f32(e, nA) / (float)nA) * 3.0f;
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < M; ++j) {
            const float* p = Bp + ((long)b * M + j) * 3;
            const float* q = A + ((long)b * N + nnB[(long)b * M + j]) * 3;
            for (int d = 0; d < 3; ++d) {
                float t = p[d] - q[d];
                e[((long)b * M + j) * 3 + d] = t * t;
            }
        }
    float dBA = (pairwise_sum_f32(e, nB) / (float)nB) * 3.0f;
    free(e);
    if (!nnA_out) free(nnA);
    if (!nnB_out) free(nnB);
    if (terms) { terms[0] = dAB; terms[1] = dBA; }
    return (w1 * dAB) + (w2 * dBA);
}

/* Zygote pullback of src/metrics/pcloud.jl:47-50 with the indices held constant (@ignore at :45):
 *   dA[:,i,b] = 2*w1*3/(3NB) (A_i - B_nnA(i)) + sum_{j: nnB(j)=i} -2*w2*3/(3MB) (B_j - A_i)
 * (and symmetrically for B).  gA: [B][N][3], gB: [B][M][3].  Accumulation in double, cast at the end:
 * the reference test pins this only to atol 1e-2 / rtol 1e-3 (test/metrics.jl:112-114). */
ORC_API void orc_chamfer_backward(const float* A, const float* Bp, int B, int N, int M, float w1,
                                  float w2, const int32_t* nnA, const int32_t* nnB, float gout,
                                  float* gA, float* gB) {
    double* dA = (double*)calloc((size_t)B * N * 3, sizeof(double));
    double* dB = (double*)calloc((size_t)B * M * 3, sizeof(double));
    double cA = 2.0 * (double)w1 * (double)gout / ((double)N * B);
    double cB = 2.0 * (double)w2 * (double)gout / ((double)M * B);
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < N; ++i) {
            long ia = ((long)b * N + i) * 3, ib = ((long)b * M + nnA[(long)b * N + i]) * 3;
            for (int d = 0; d < 3; ++d) {
                double t = (double)A[ia + d] - (double)Bp[ib + d];
                dA[ia + d] += cA * t;
                dB[ib + d] -= cA * t;
            }
        }
        for (int j = 0; j < M; ++j) {
            long ib = ((long)b * M + j) * 3, ia = ((long)b * N + nnB[(long)b * M + j]) * 3;
            for (int d = 0; d < 3; ++d) {
                double t = (double)Bp[ib + d] - (double)A[ia + d];
                dB[ib + d] += cB * t;
                dA[ia + d] -= cB * t;
            }
        }
    }
    for (long k = 0; k < (long)B * N * 3; ++k) gA[k] = (float)dA[k];
    for (long k = 0; k < (long)B * M * 3; ++k) gB[k] = (float)dB[k];
    free(dA);
    free(dB);
}

/* ------------------------------------------------------------------------------------------
 * CreateSingleKNNGraph(X,K) — src/models/dgcnn.jl:3-7, batched as in EdgeConv :36.
 *   knn(kdtree, X[:,i], K+1, true)[1][2:K+1]  : (K+1)-NN list sorted ascending, first dropped
 *   *by position*.  Order rule adopted for ties: ascending (distance, index).
 * X: [B][N][F].  idx: [B][N][K] (0-based).  dist (opt): [B][N][K] squared distances.
 * gathered (opt): [B][N][K][F]  == Julia (F,K,N,B) column-major.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float d; int32_t j; } orc_cand;

static int cand_cmp(const void* a, const void* b) {
    const orc_cand* x = (const orc_cand*)a;
    const orc_cand* y = (const orc_cand*)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->j > y->j) - (x->j < y->j);
}

ORC_API int orc_knn_graph(const float* X, int B, int N, int F, int K, int32_t* idx, float* dist,
                          float* gathered) {
    if (K < 1 || K + 1 > N) return 1;
#pragma omp parallel
    {
        orc_cand* c = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)N);
#pragma omp for schedule(dynamic, 8) collapse(2)
        for (int b = 0; b < B; ++b) {
            for (int i = 0; i < N; ++i) {
                const float* base = X + (long)b * N * F;
                const float* p = base + (long)i * F;
                for (int j = 0; j < N; ++j) {
                    c[j].d = sqdist(p, base + (long)j * F, F);
                    c[j].j = j;
                }
                qsort(c, (size_t)N, sizeof(orc_cand), cand_cmp);
                for (int k = 0; k < K; ++k) {
                    long o = ((long)b * N + i) * K + k;
                    idx[o] = c[k + 1].j;
                    if (dist) dist[o] = c[k + 1].d;
                    if (gathered) memcpy(gathered + o * F, base + (long)c[k + 1].j * F, sizeof(float) * F);
                }
            }
        }
        free(c);
    }
    return 0;
}

/* EdgeConv prologue — src/models/dgcnn.jl:39-45: cat(X_tiled, KNNGraph - X_tiled; dims=1)
 * → (2F,K,N,B) column-major == C [B][N][K][2F]: first F = x_i, next F = x_j - x_i. */
ORC_API void orc_edge_features(const float* X, const int32_t* idx, int B, int N, int F, int K,
                               float* out) {
    for (long b = 0; b < B; ++b)
        for (long i = 0; i < N; ++i)
            for (long k = 0; k < K; ++k) {
                const float* xi = X + (b * N + i) * F;
                const float* xj = X + (b * N + idx[(b * N + i) * K + k]) * F;
                float* o = out + ((b * N + i) * K + k) * 2 * F;
                for (int f = 0; f < F; ++f) {
                    o[f] = xi[f];
                    o[F + f] = xj[f] - xi[f];
                }
            }
}

/* ------------------------------------------------------------------------------------------
 * _lg_cross — src/rep/utils.jl:4-21:  (a2*b3 - a3*b2, a3*b1 - a1*b3, a1*b2 - a2*b1)
 * ---------------------------------------------------------------------------------------- */
static inline void lg_cross(const float* a, const float* b, float* c) {
    c[0] = (a[1] * b[2]) - (a[2] * b[1]);
    c[1] = (a[2] * b[0]) - (a[0] * b[2]);
    c[2] = (a[0] * b[1]) - (a[1] * b[0]);
}
static inline void sub3(const float* a, const float* b, float* c) {
    c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2];
}
/* _norm(A; dims=1) on a 3-vector — src/rep/utils.jl:29: sqrt(sum(A.^2)) with sequential sum */
static inline float norm3(const float* c) {
    return sqrtf(((c[0] * c[0]) + (c[1] * c[1])) + (c[2] * c[2]));
}

/* compute_faces_areas_packed — src/rep/mesh.jl:765-780 and compute_faces_normals_packed — :689-700
 * verts: [nV][3]; faces: [nF][3] 0-based packed (global) indices.  areas/normals may be NULL. */
ORC_API void orc_faces_areas_normals(const float* verts, const int32_t* faces, int nV, int nF,
                                     float* areas, float* normals) {
    (void)nV;
    for (int f = 0; f < nF; ++f) {
        const float* v1 = verts + 3L * faces[3L * f + 0];
        const float* v2 = verts + 3L * faces[3L * f + 1];
        const float* v3 = verts + 3L * faces[3L * f + 2];
        float e1[3], e2[3], c[3];
        sub3(v2, v1, e1);
        sub3(v3, v1, e2);
        lg_cross(e1, e2, c);
        float n = norm3(c);
        if (areas) areas[f] = n / 2.0f;
        if (normals) { /* _normalize: A ./ max(norm, 1e-6) — src/rep/utils.jl:23-27 */
            float m = fmaxf(n, 1e-6f);
            normals[3L * f + 0] = c[0] / m;
            normals[3L * f + 1] = c[1] / m;
            normals[3L * f + 2] = c[2] / m;
        }
    }
}

/* compute_verts_normals_packed — src/rep/mesh.jl:589-618.
 * mode 0 = REFERENCE_CPU: Zygote.Buffer gather-add-assign, so for each corner slot k only the LAST
 *          face (highest face index) holding the vertex in slot k contributes; slots are applied
 *          in order 1,2,3, each on top of the previous slots' result ((0+c1)+c2)+c3.
 * mode 1 = ACCUMULATE: the documented intent — sum over every incident corner, in face order,
 *          slot 1 of all faces, then slot 2, then slot 3 (the order a serial scatter-add of
 *          :604-615 would use). */
ORC_API void orc_verts_normals(const float* verts, const int32_t* faces, int nV, int nF, int mode,
                               float* out) {
    float* vn = (float*)calloc((size_t)nV * 3, sizeof(float));
    for (int k = 0; k < 3; ++k) {
        int ka = (k + 1) % 3, kb = (k + 2) % 3;
        if (mode == 0) {
            /* gathered = vn[:, faces[k,:]] BEFORE any assignment of this slot */
            float* tmp = (float*)malloc(sizeof(float) * 3 * (size_t)nF);
            for (int f = 0; f < nF; ++f) {
                const float* vk = verts + 3L * faces[3L * f + k];
                float e1[3], e2[3], c[3];
                sub3(verts + 3L * faces[3L * f + ka], vk, e1);
                sub3(verts + 3L * faces[3L * f + kb], vk, e2);
                lg_cross(e1, e2, c);
                const float* cur = vn + 3L * faces[3L * f + k];
                tmp[3L * f + 0] = cur[0] + c[0];
                tmp[3L * f + 1] = cur[1] + c[1];
                tmp[3L * f + 2] = cur[2] + c[2];
            }
            for (int f = 0; f < nF; ++f) memcpy(vn + 3L * faces[3L * f + k], tmp + 3L * f, 3 * sizeof(float));
            free(tmp);
        } else {
            for (int f = 0; f < nF; ++f) {
                const float* vk = verts + 3L * faces[3L * f + k];
                float e1[3], e2[3], c[3];
                sub3(verts + 3L * faces[3L * f + ka], vk, e1);
                sub3(verts + 3L * faces[3L * f + kb], vk, e2);
                lg_cross(e1, e2, c);
                float* cur = vn + 3L * faces[3L * f + k];
                cur[0] = cur[0] + c[0];
                cur[1] = cur[1] + c[1];
                cur[2] = cur[2] + c[2];
            }
        }
    }
    for (int v = 0; v < nV; ++v) {
        float m = fmaxf(norm3(vn + 3L * v), 1e-6f);
        out[3L * v + 0] = vn[3L * v + 0] / m;
        out[3L * v + 1] = vn[3L * v + 1] / m;
        out[3L * v + 2] = vn[3L * v + 2] / m;
    }
    free(vn);
}

/* ------------------------------------------------------------------------------------------
 * _compute_edges_packed — src/rep/mesh.jl:907-955.
 * Unique undirected edges (min,max) sorted lexicographically; faces_to_edges column order
 * (e23, e31, e12).  The reference's hash is computed in the face-index type R and overflows for
 * UInt32 when ΣV > 65535; this restatement uses int64 keys (documented divergence).
 * edges: [maxE][2] with maxE >= 3*nF; f2e: [nF][3] or NULL.  Returns nE.
 * ---------------------------------------------------------------------------------------- */
static int i64_cmp(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

ORC_API int orc_edges_packed(const int32_t* faces, int nV, int nF, int32_t* edges, int32_t* f2e) {
    int64_t H = (int64_t)nV + 1;
    long n3 = 3L * nF;
    int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n3 > 0 ? n3 : 1));
    for (int f = 0; f < nF; ++f) {
        for (int e = 0; e < 3; ++e) { /* e12, e23, e31 (1-based values as in Julia, so hashes match) */
            int64_t a = (int64_t)faces[3L * f + e] + 1, b = (int64_t)faces[3L * f + (e + 1) % 3] + 1;
            int64_t lo = a < b ? a : b, hi = a < b ? b : a;
            keys[(long)e * nF + f] = H * lo + hi;
        }
    }
    qsort(keys, (size_t)n3, sizeof(int64_t), i64_cmp);
    long nE = 0;
    for (long i = 0; i < n3; ++i)
        if (i == 0 || keys[i] != keys[i - 1]) keys[nE++] = keys[i];
    for (long e = 0; e < nE; ++e) {
        edges[2 * e + 0] = (int32_t)(keys[e] / H) - 1;
        edges[2 * e + 1] = (int32_t)(keys[e] % H) - 1;
    }
    if (f2e) {
        for (int f = 0; f < nF; ++f) {
            /* columns: e23, e31, e12 */
            const int pair[3][2] = {{1, 2}, {2, 0}, {0, 1}};
            for (int c = 0; c < 3; ++c) {
                int64_t a = (int64_t)faces[3L * f + pair[c][0]] + 1, b = (int64_t)faces[3L * f + pair[c][1]] + 1;
                int64_t lo = a < b ? a : b, hi = a < b ? b : a;
                int64_t key = H * lo + hi;
                int64_t* hit = (int64_t*)bsearch(&key, keys, (size_t)nE, sizeof(int64_t), i64_cmp);
                f2e[3L * f + c] = hit ? (int32_t)(hit - keys) : -1;
            }
        }
    }
    free(keys);
    return (int)nE;
}

/* ------------------------------------------------------------------------------------------
 * _compute_laplacian_packed — src/rep/mesh.jl:957-1002, as CSR (row = vertex i):
 *   L[i,i] = -1 ; L[i,j] = Float32(1/deg(i)) for every edge (i,j); columns ascending.
 * rowptr: [nV+1]; colidx/vals: [2*nE + nV].
 * ---------------------------------------------------------------------------------------- */
static int i32_cmp(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

ORC_API void orc_laplacian_csr(const int32_t* edges, int nE, int nV, int32_t* rowptr,
                               int32_t* colidx, float* vals) {
    int32_t* deg = (int32_t*)calloc((size_t)nV + 1, sizeof(int32_t));
    for (int e = 0; e < nE; ++e) { deg[edges[2 * e]]++; deg[edges[2 * e + 1]]++; }
    rowptr[0] = 0;
    for (int v = 0; v < nV; ++v) rowptr[v + 1] = rowptr[v] + deg[v] + 1;
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nV > 0 ? nV : 1));
    for (int v = 0; v < nV; ++v) { fill[v] = rowptr[v]; colidx[fill[v]++] = v; }
    for (int e = 0; e < nE; ++e) {
        int a = edges[2 * e], b = edges[2 * e + 1];
        colidx[fill[a]++] = b;
        colidx[fill[b]++] = a;
    }
    for (int v = 0; v < nV; ++v) {
        qsort(colidx + rowptr[v], (size_t)(rowptr[v + 1] - rowptr[v]), sizeof(int32_t), i32_cmp);
        /* T.(x > 0 ? 1/x : x): 1/deg in Float64, then rounded to Float32 — :985-986 */
        float w = deg[v] > 0 ? (float)(1.0 / (double)deg[v]) : 0.0f;
        for (int p = rowptr[v]; p < rowptr[v + 1]; ++p) vals[p] = (colidx[p] == v) ? -1.0f : w;
    }
    free(fill);
    free(deg);
}

/* laplacian_loss — src/metrics/mesh.jl:9-15.  SparseMatrixCSC * dense accumulates, for each output
 * entry, in ascending column order: C[i,k] += L[i,j]*X[j,k]  (SparseArrays mul!, no muladd);
 * then _norm(dims=2) = sqrt((x²+y²)+z²), then mean over ΣV (pairwise). */
ORC_API float orc_laplacian_loss(const float* verts, const int32_t* rowptr, const int32_t* colidx,
                                 const float* vals, int nV) {
    float* nrm = (float*)malloc(sizeof(float) * (size_t)(nV > 0 ? nV : 1));
    for (int v = 0; v < nV; ++v) {
        float acc[3] = {0.0f, 0.0f, 0.0f};
        for (int p = rowptr[v]; p < rowptr[v + 1]; ++p) {
            const float* x = verts + 3L * colidx[p];
            float w = vals[p];
            acc[0] = acc[0] + w * x[0];
            acc[1] = acc[1] + w * x[1];
            acc[2] = acc[2] + w * x[2];
        }
        nrm[v] = norm3(acc);
    }
    float r = pairwise_sum_f32(nrm, nV) / (float)nV;
    free(nrm);
    return r;
}

/* edge_loss — src/metrics/mesh.jl:24-32: mean((‖v1-v2‖ - target)^2) over packed unique edges. */
ORC_API float orc_edge_loss(const float* verts, const int32_t* edges, int nE, float target) {
    float* el = (float*)malloc(sizeof(float) * (size_t)(nE > 0 ? nE : 1));
    for (int e = 0; e < nE; ++e) {
        float d[3];
        sub3(verts + 3L * edges[2 * e], verts + 3L * edges[2 * e + 1], d);
        float t = norm3(d) - target;
        el[e] = t * t;
    }
    float r = pairwise_sum_f32(el, nE) / (float)nE;
    free(el);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) — the counter RNG the product uses when draws are not
 * injected.  The reference draws from Julia's global RNG + Distributions' alias sampler
 * (src/transforms/mesh_func.jl:46-47,76-77), which cannot be reproduced; seeded parity is
 * therefore oracle-vs-kernel only, reference parity is via injected draws + statistics.
 * ---------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* Draws for sample s of mesh i: counter = (s, i, offset_lo, offset_hi), key = seed.
 * (out0,out1) → u_face, a 53-bit uniform integer (u_face / 2^53 is the Float64 uniform in [0,1));
 * r1,r2 Float32 in [0,1) from the top 24 bits of out2,out3 (Julia's rand(Float32) also yields multiples of 2^-24... of 2^-23 in
 * older versions; either way a uniform grid on [0,1)). */
ORC_API void orc_philox_draws(uint64_t seed, uint64_t offset, int mesh, int s, uint64_t* u_face,
                              float* r1, float* r2) {
    uint32_t c[4] = {(uint32_t)s, (uint32_t)mesh, (uint32_t)offset, (uint32_t)(offset >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *u_face = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    *r1 = (float)(c[2] >> 8) * (1.0f / 16777216.0f);
    *r2 = (float)(c[3] >> 8) * (1.0f / 16777216.0f);
}

/* sample_points — src/transforms/mesh_func.jl:21-82.
 * verts_padded: [Nmesh][Vmax][3]; faces_padded: [Nmesh][Fmax][3] local 0-based (pad = anything);
 * Face probabilities: Float64 area / max(Σarea, eps) (:32-39; Σ sequential in Float64 as Julia's
 * sum(...; dims=2) does).  The reference then draws from Distributions.Categorical (alias tables)
 * with Julia's global RNG — irreproducible — so the draw is DEFINED here as inverse-CDF with the CDF
 * held in 53-bit fixed point: C_f = Σ_{g<=f} floor(p_g * 2^53) (integer sums: order-independent),
 * face = smallest f with C_f > m for a 53-bit uniform integer m, clamped to the last valid face
 * (which thereby absorbs the rounding residual, the analogue of :36-37).
 * If inj_face != NULL the face ids / r1 / r2 are taken from the injected arrays ([Nmesh][S]) instead
 * of Philox — this is the mode in which parity with the reference's arithmetic (:60-82) is bit-exact:
 *   u = sqrt(r1); w1 = 1-u; w2 = u*(1-v); w3 = u*v;  p = ((w1*v1)+(w2*v2))+(w3*v3)
 * samples: [Nmesh][S][3]; face_idx_out (opt): [Nmesh][S]. */
ORC_API void orc_sample_points(const float* verts_padded, const int32_t* faces_padded,
                               const int32_t* verts_len, const int32_t* faces_len, int Nmesh,
                               int Vmax, int Fmax, int S, double eps, uint64_t seed,
                               uint64_t offset, const int32_t* inj_face, const float* inj_r1,
                               const float* inj_r2, float* samples, int32_t* face_idx_out) {
    (void)verts_len;
    uint64_t* cdf = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(Fmax > 0 ? Fmax : 1));
    float* areas = (float*)malloc(sizeof(float) * (size_t)(Fmax > 0 ? Fmax : 1));
    for (int i = 0; i < Nmesh; ++i) {
        const float* V = verts_padded + (long)i * Vmax * 3;
        const int32_t* Fc = faces_padded + (long)i * Fmax * 3;
        int nF = faces_len[i];
        orc_faces_areas_normals(V, Fc, Vmax, nF, areas, NULL);
        double tot = 0.0;
        for (int f = 0; f < nF; ++f) tot = tot + (double)areas[f];
        double den = tot > eps ? tot : eps;
        uint64_t run = 0;
        for (int f = 0; f < nF; ++f) { run += (uint64_t)(((double)areas[f] / den) * 9007199254740992.0); cdf[f] = run; }
        for (int s = 0; s < S; ++s) {
            int face;
            float r1, r2;
            if (inj_face) {
                face = inj_face[(long)i * S + s];
                r1 = inj_r1[(long)i * S + s];
                r2 = inj_r2[(long)i * S + s];
            } else {
                uint64_t u;
                orc_philox_draws(seed, offset, i, s, &u, &r1, &r2);
                int lo = 0, hi = nF - 1; /* smallest f with cdf[f] > u, clamped to nF-1 */
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
                }
                face = lo;
            }
            const float* v1 = V + 3L * Fc[3L * face + 0];
            const float* v2 = V + 3L * Fc[3L * face + 1];
            const float* v3 = V + 3L * Fc[3L * face + 2];
            float u_ = sqrtf(r1), v_ = r2;
            float w1 = 1.0f - u_;
            float w2 = u_ * (1.0f - v_);
            float w3 = u_ * v_;
            float* o = samples + ((long)i * S + s) * 3;
            for (int d = 0; d < 3; ++d) o[d] = ((w1 * v1[d]) + (w2 * v2[d])) + (w3 * v3[d]);
            if (face_idx_out) face_idx_out[(long)i * S + s] = face;
        }
    }
    free(cdf);
    free(areas);
}
