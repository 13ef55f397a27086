/*
 * flux3d_b200.h — C ABI of libflux3d_b200.so: the Blackwell (sm_100a) implementation of the
 * Flux3D.jl batched 3D-metric hot path.  This is the drop-in boundary: a Julia `ccall`
 * (flux3d.jl_b200/julia/Flux3DB200.jl), Python ctypes (flux3d.jl_b200/_lib.py) or any other FFI
 * binds exactly these symbols.  No torch / CUDA.jl types appear: plain device pointers, sizes and a
 * CUDA stream handle.
 *
 * Conventions
 *   - Every function returns an int32 status: 0 OK, 1 invalid argument, 2 misaligned buffer,
 *     3 workspace too small, 4 CUDA error, 5 NCCL error.  f3d_last_error() gives the message of the
 *     last failure on the calling thread.  No exceptions cross the boundary.
 *   - All array pointers are DEVICE pointers unless the name ends in `_host`.  The caller owns every
 *     buffer, workspace included; the library allocates nothing on the hot path and launches
 *     asynchronously on `stream` (no host synchronisation).
 *   - Layout: a Julia (3,N,B) Float32 array == C [B][N][3] (xyz interleaved).  Julia (F,K,N,B) ==
 *     C [B][N][K][F].  Indices are 0-based int32 here; the Julia shim adds 1.
 *   - Arithmetic: every distance / cross product / normalisation is evaluated in the operation
 *     order of the Julia reference with separately rounded IEEE binary32 operations (Julia does not
 *     contract a*b+c), unless F3D_FLAG_FMA is passed.
 *
 * Each entry point cites the reference code (FluxML/Flux3D.jl v0.1.6, file:line) it replaces.
 */
#ifndef FLUX3D_B200_H
#define FLUX3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* f3d_stream_t; /* cudaStream_t / CUstream */

#if defined(__GNUC__)
#define F3D_API __attribute__((visibility("default")))
#else
#define F3D_API
#endif

enum {
    F3D_OK = 0,
    F3D_ERR_INVALID = 1,
    F3D_ERR_MISALIGNED = 2,
    F3D_ERR_WORKSPACE = 3,
    F3D_ERR_CUDA = 4,
    F3D_ERR_NCCL = 5
};

enum {
    F3D_FLAG_NONE = 0,
    /* Evaluate squared distances as fma(dz,dz,fma(dy,dy,dx*dx)) instead of the reference's
       ((dx*dx)+(dy*dy))+(dz*dz).  Faster (6 instead of 8 FP32 operations per pair) but NOT the
       reference arithmetic: near-ties may resolve differently.  Off by default. */
    F3D_FLAG_FMA = 1,
    /* Measurement aid for f3d_chamfer_fwd: launch only the pairwise sweep kernel (the dominant kernel)
       and skip the finalize pass, so a caller can bracket exactly that kernel with events.  The outputs
       are NOT written in this mode. */
    F3D_FLAG_SWEEP_ONLY = 2,
    /* f3d_chamfer_fwd: evaluate EVERY pair in the reference arithmetic (the original exact sweep) instead
       of the default filter + certified exact re-evaluation.  Both produce bit-identical results; this
       one does not depend on the filter's error bound and is kept as the cross-check. */
    F3D_FLAG_EXACT_SWEEP = 4,
    /* f3d_knn_graph: take a tensor-core (tcgen05) filter path whenever the shape allows it.  With the workspace of
       f3d_knn_graph_workspace_bytes that is the default anyway (knn_gram.cu: F <= 64, K <= 31, N <= 2048 and at least
       1.5 (K+1) chunks of 16 / 32 candidates); the flag also selects the older cp.async-staged filter (knn_tc.cu, N <= 1024)
       for narrow features (F < 16) when no workspace is given.  Results are identical either way. */
    F3D_FLAG_TENSOR = 8,
    /* f3d_chamfer_fwd: keep the filter sweep on the CUDA cores (packed-FP32 FFMA2 expanded form, chamfer.cu) also for
       problems large enough for the tensor-core sweep (tcgen05 split-TF32 filter, chamfer_tc.cu), which is the default
       when both clouds have at least 512 points.  Results are bit-identical either way. */
    F3D_FLAG_CUDA_CORES = 16,
    /* f3d_knn_graph: write edge_feat in the layout EdgeConv's 1x1-convolution MLP consumes — C [B][2F][N][K], i.e. the Julia
       (K*N, 2F, B) array of src/models/dgcnn.jl:46-52 — instead of [B][N][K][2F] == Julia (2F, K, N, B) (:45): the reference's
       PermutedDimsArray + reshape copy of the whole edge tensor disappears. */
    F3D_FLAG_EDGE_MLP_LAYOUT = 32
};

enum {
    F3D_NORMALS_REFERENCE_CPU = 0, /* last face per corner slot wins (what rep/mesh.jl:604-615 does on CPU) */
    F3D_NORMALS_ACCUMULATE = 1     /* sum over all incident corners (what its docstring says) */
};

F3D_API int32_t f3d_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf; returns its length. */
F3D_API int32_t f3d_last_error(char* buf, size_t n);

/* ------------------------------------------------------------------------------------------------
 * chamfer_distance — replaces _chamfer_distance + _nearest_neighbors(::CuArray, ::CuArray)
 * (src/metrics/pcloud.jl:39-52 and :72-86; semantics of the CPU method :54-70).
 *
 *   A [B][N][3], Bp [B][M][3]  →  loss_dev[0] = w1*ΣΣ‖a-b_nn(a)‖²/(N*B_total) + w2*ΣΣ‖b-a_nn(b)‖²/(M*B_total)
 *
 * B_total is the GLOBAL batch size of the mean (pass B, or the un-sharded batch when this call
 * handles one shard of a batch split across GPUs; the shard results then add up to the reference
 * value — see f3d_allreduce_sum_f32).  terms_dev (optional, 2 floats) receives the two un-weighted
 * means.  nnA_dev [B][N] / nnB_dev [B][M] (optional) receive the argmin indices (ties → lowest).
 * Brute force, O(N+M) memory: the N×M matrix of :75-78 is never materialised.
 * ---------------------------------------------------------------------------------------------- */
F3D_API size_t f3d_chamfer_workspace_bytes(int32_t B, int32_t N, int32_t M);
F3D_API int32_t f3d_chamfer_fwd(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1,
                        float w2, int32_t B_total, float* loss_dev, float* terms_dev,
                        int32_t* nnA_dev, int32_t* nnB_dev, void* ws, size_t ws_bytes,
                        int32_t flags, f3d_stream_t stream);

/* chamfer_distance on HOST arrays — the array entry points chamfer_distance(A::AbstractArray, B::AbstractArray; w1, w2)
 * (src/metrics/pcloud.jl:28-37) called with `Array`s: upload, sweep and loss read-back as ONE call, with the upload
 * running inside the sweep grid.  When both arrays are page-locked (cudaHostAlloc / cudaHostRegister: the device can
 * address them) and 16-byte aligned, the grid's first `uploaders` CTAs pull the clouds over PCIe batch element by batch
 * element while the other CTAs sweep, each waiting only for its own element; otherwise the arrays are copied with
 * cudaMemcpyAsync on `stream` first.  Either way the result is bit-identical to f3d_chamfer_fwd on resident inputs.
 *   f3d_chamfer_pipe_create: uploaders = CTAs (x128 threads) that upload, 0 = default (32).  The handle owns 64 bytes of
 *     mapped host memory (no device memory); one handle per (device, host thread).
 *   A_host [B][N][3], B_host [B][M][3]: HOST arrays; they must stay valid until the call's work on `stream` is done.
 *   loss_dev (optional, 1 float): device copy of the loss, valid in `stream` order.
 *   loss_host (optional): when given, the grid stores the loss into mapped host memory and the call returns once it
 *     has landed (no D2H copy) — the one entry point of this ABI that blocks, because a host scalar was asked for.
 *   ws: f3d_chamfer_pipe_workspace_bytes(B, N, M) device bytes (staging copies of both clouds + the sweep workspace),
 *     256-byte aligned, owned by the caller.
 *   comm (optional): a communicator with peer mailboxes (f3d_comm_enable_p2p) — the call then handles this rank's shard
 *     of a B_total batch and the loss it delivers is the whole batch's (see f3d_chamfer_fwd_allreduce). */
F3D_API int32_t f3d_chamfer_pipe_create(int32_t uploaders, void** pipe);
F3D_API size_t f3d_chamfer_pipe_workspace_bytes(int32_t B, int32_t N, int32_t M);
F3D_API int32_t f3d_chamfer_pipe_run(void* pipe, const float* A_host, const float* B_host, int32_t B, int32_t N,
                             int32_t M, float w1, float w2, int32_t B_total, float* loss_dev,
                             float* loss_host, void* ws, size_t ws_bytes, int32_t flags, void* comm,
                             f3d_stream_t stream);
F3D_API int32_t f3d_chamfer_pipe_destroy(void* pipe);

/* Pullback of src/metrics/pcloud.jl:47-50 (indices constant, :45 is @ignore):
 *   gA = gout*( 2w1/(N*B_total) (A - B[nnA])  -  scatter_add_{nnB}( 2w2/(M*B_total) (B - A[nnB]) ) ), gB symmetric.
 * gout_dev: 1 float (upstream gradient of the scalar loss).  gA/gB are fully overwritten. */
F3D_API int32_t f3d_chamfer_bwd(const float* A, const float* Bp, int32_t B, int32_t N, int32_t M, float w1,
                        float w2, int32_t B_total, const int32_t* nnA_dev, const int32_t* nnB_dev,
                        const float* gout_dev, float* gA, float* gB, f3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * kNN graph — replaces CreateSingleKNNGraph + the batch loop / gather / concat prologue of EdgeConv
 * (src/models/dgcnn.jl:3-9 and :32-45).
 *   X [B][N][F]; for every point the K nearest OTHER points = positions 2..K+1 of the (K+1)-NN list
 *   sorted ascending by (squared distance, index).  1 <= K < N, K <= 63, F <= 256.
 *   idx [B][N][K] (required); dist [B][N][K] (optional squared distances);
 *   gathered [B][N][K][F] (optional; == the Julia (F,K,N,B) KNNGraph tensor of :36);
 *   edge_feat [B][N][K][2F] (optional; == cat(X, KNNGraph - X; dims=1) of :45); with F3D_FLAG_EDGE_MLP_LAYOUT it is written
 *     as [B][2F][N][K] == the (K*N, 2F, B) input of the MLP (:46-52) instead.
 *   N <= 1024, 16 <= F <= 64, K <= 31 run the Gram matrix on the tensor cores (tcgen05, TF32) as a filter and re-evaluate the
 *   surviving candidates in the reference arithmetic — results are bit-identical to the all-exact CUDA-core kernel,
 *   which F3D_FLAG_EXACT_SWEEP (or any larger shape) selects.  ws is optional: when >= 8 bytes are given, two
 *   uint32 diagnostics are written to it — queries whose candidate set overflowed to an exact scan of the cloud,
 *   and the number of candidates re-evaluated exactly.
 * ---------------------------------------------------------------------------------------------- */
F3D_API size_t f3d_knn_graph_workspace_bytes(int32_t B, int32_t N, int32_t F, int32_t K);
F3D_API int32_t f3d_knn_graph(const float* X, int32_t B, int32_t N, int32_t F, int32_t K, int32_t* idx,
                      float* dist, float* gathered, float* edge_feat, void* ws, size_t ws_bytes,
                      int32_t flags, f3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * TriMesh kernels on packed verts [nV][3] / packed faces [nF][3] (global 0-based vertex ids).
 * ---------------------------------------------------------------------------------------------- */
/* compute_faces_areas_packed (src/rep/mesh.jl:765-780) and compute_faces_normals_packed (:689-700).
 * areas [nF] and/or normals [nF][3] may be NULL. */
F3D_API int32_t f3d_faces_areas_normals(const float* verts, const int32_t* faces, int32_t nV, int32_t nF,
                                float* areas, float* normals, f3d_stream_t stream);

/* Topology products, built ONCE per mesh topology on the HOST and cached by the caller, mirroring
 * the cached fields of TriMesh (src/rep/mesh.jl:93-97):
 *   edges_host [<=3nF][2] unique (min,max) edges sorted lexicographically (_compute_edges_packed :907-955)
 *   f2e_host [nF][3]      faces→edges, column order (e23,e31,e12)                (:943-949)   (optional)
 *   lap_rowptr_host [nV+1], lap_colidx_host [2nE+nV], lap_vals_host [2nE+nV]: CSR of the Laplacian,
 *                          L[i,i]=-1, L[i,j]=Float32(1/deg i), columns ascending (_compute_laplacian_packed :957-1002)
 *   v2c_rowptr_host [nV+1], v2c_host [3nF]: vertex → incident corners (face*3+slot), ordered by
 *                          (slot, face) — the gather form of the scatter at :604-615 in the order a
 *                          serial execution applies it (deterministic, no atomics).
 * All pointers are HOST pointers; *nE_host receives the edge count. */
F3D_API int32_t f3d_mesh_topology_build_host(const int32_t* faces_host, int32_t nV, int32_t nF,
                                     int32_t* edges_host, int32_t* nE_host, int32_t* f2e_host,
                                     int32_t* lap_rowptr_host, int32_t* lap_colidx_host,
                                     float* lap_vals_host, int32_t* v2c_rowptr_host,
                                     int32_t* v2c_host);

/* _packed_to_padded / _padded_to_packed (src/rep/utils.jl:131-185) on the device, for 4-byte elements (Float32 verts and
 * normals, Int32 faces): packed [ΣL][D] with item i = rows offsets[i] .. offsets[i+1]-1 (offsets [N+1], device)  <->
 * padded [N][W][D]; rows past an item's length are filled with the bit pattern fill_bits (e.g. 0 for 0.0f, 0xffffffff
 * for the -1 of padded faces).  delta (optional, [N], device) is subtracted from every real element on the way to
 * padded and added on the way to packed: the global <-> local vertex ids of packed / padded faces
 * (src/rep/mesh.jl:884-896).  total_rows = offsets[N]. */
F3D_API int32_t f3d_packed_to_padded(const void* packed, const int32_t* offsets, const int32_t* delta, int32_t N,
                             int32_t W, int32_t D, uint32_t fill_bits, void* padded, f3d_stream_t stream);
F3D_API int32_t f3d_padded_to_packed(const void* padded, const int32_t* offsets, const int32_t* delta, int32_t N,
                             int32_t W, int32_t D, int32_t total_rows, void* packed, f3d_stream_t stream);

/* compute_verts_normals_packed (src/rep/mesh.jl:589-618).  mode: F3D_NORMALS_*.  out [nV][3]. */
F3D_API int32_t f3d_verts_normals(const float* verts, const int32_t* faces, const int32_t* v2c_rowptr,
                          const int32_t* v2c, int32_t nV, int32_t nF, int32_t mode, float* out,
                          f3d_stream_t stream);

/* laplacian_loss (src/metrics/mesh.jl:9-15): mean_i ‖Σ_j L[i,j] v_j‖₂ over all nV packed vertices.
 * nV_total: global vertex count of the mean (nV, or the un-sharded count for a mesh-sharded call).
 * ws: f3d_laplacian_workspace_bytes(nV). */
F3D_API size_t f3d_laplacian_workspace_bytes(int32_t nV);
F3D_API int32_t f3d_laplacian_loss(const float* verts, const int32_t* lap_rowptr, const int32_t* lap_colidx,
                           const float* lap_vals, int32_t nV, int32_t nV_total, float* loss_dev,
                           void* ws, size_t ws_bytes, f3d_stream_t stream);
/* Pullback: gverts[j] = gout * Σ_i L[i,j] * n̂_i / nV_total, n̂_i = (Lv)_i/‖(Lv)_i‖ (0 where the norm is 0).
 * Uses the CSR of Lᵀ = same pattern (L's pattern is symmetric) with values 1/deg(i) per source row. */
F3D_API int32_t f3d_laplacian_loss_bwd(const float* verts, const int32_t* lap_rowptr,
                               const int32_t* lap_colidx, const float* lap_vals, int32_t nV,
                               int32_t nV_total, const float* gout_dev, float* gverts, void* ws,
                               size_t ws_bytes, f3d_stream_t stream);

/* edge_loss (src/metrics/mesh.jl:24-32): mean_e (‖v_e1 - v_e2‖ - target)².  ws: f3d_edge_loss_workspace_bytes(nE). */
F3D_API size_t f3d_edge_loss_workspace_bytes(int32_t nE);
F3D_API int32_t f3d_edge_loss(const float* verts, const int32_t* edges, int32_t nE, int32_t nE_total,
                      float target, float* loss_dev, void* ws, size_t ws_bytes, f3d_stream_t stream);
/* Pullback: gverts[i] = gout * Σ_{j ∈ N(i)} (2/nE_total) (‖v_i - v_j‖ - target) (v_i - v_j)/‖v_i - v_j‖   (0 where the
 * norm is 0), gathered over the vertex's neighbours = the off-diagonal columns of the Laplacian CSR
 * (deterministic, no atomics).  nE_total: the edge count of the mean (0 = the nE implied by the CSR). */
F3D_API int32_t f3d_edge_loss_bwd(const float* verts, const int32_t* lap_rowptr, const int32_t* lap_colidx,
                          int32_t nV, int32_t nE_total, float target, const float* gout_dev,
                          float* gverts, f3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * sample_points — replaces sample_points/_sample_points/_rand_barycentric_coords
 * (src/transforms/mesh_func.jl:21-82).
 *   verts_padded [Nmesh][Vmax][3], faces_padded [Nmesh][Fmax][3] (LOCAL 0-based ids, padding = any),
 *   verts_len/faces_len [Nmesh] (device).  samples [Nmesh][S][3]; face_idx_out [Nmesh][S] optional.
 *   Face probabilities are Float64 area/max(Σarea,eps) as at :32-39.  Draws: Philox4x32-10 keyed by
 *   (seed, offset) — counter (s, mesh) — unless inj_face/inj_r1/inj_r2 ([Nmesh][S], device) are given,
 *   in which case the face ids and the two uniforms are taken from them (bit-parity mode; the
 *   reference's own draws come from Julia's global RNG and cannot be reproduced).
 *   bary_out [Nmesh][S][3] (optional) receives the barycentric weights (w1,w2,w3) of every sample — with
 *   face_idx_out it is what the pullback needs.
 *   ws: f3d_sample_points_workspace_bytes(Nmesh, Fmax).
 * ---------------------------------------------------------------------------------------------- */
F3D_API size_t f3d_sample_points_workspace_bytes(int32_t Nmesh, int32_t Fmax);
F3D_API int32_t f3d_sample_points(const float* verts_padded, const int32_t* faces_padded,
                          const int32_t* verts_len, const int32_t* faces_len, int32_t Nmesh,
                          int32_t Vmax, int32_t Fmax, int32_t S, double eps, uint64_t seed,
                          uint64_t offset, const int32_t* inj_face, const float* inj_r1,
                          const float* inj_r2, float* samples, int32_t* face_idx_out, float* bary_out,
                          void* ws, size_t ws_bytes, f3d_stream_t stream);
/* f3d_sample_points with a run-time draw counter, for CUDA-graph capture (fit_mesh draws fresh samples every iteration,
 * examples/fit_mesh.jl:78-84; a captured launch would otherwise repeat the draws baked into its parameters):
 *   offset_dev (device, one 64-bit word; NULL = f3d_sample_points): its value is ADDED to `offset` when the kernel runs, and it is
 *   incremented by one after the draws (stream-ordered), so every replay of a graph that contains this call uses a new block of
 *   Philox counters.  Ignored in the injected-draw mode. */
F3D_API int32_t f3d_sample_points_replayable(const float* verts_padded, const int32_t* faces_padded,
                          const int32_t* verts_len, const int32_t* faces_len, int32_t Nmesh,
                          int32_t Vmax, int32_t Fmax, int32_t S, double eps, uint64_t seed,
                          uint64_t offset, uint64_t* offset_dev, const int32_t* inj_face, const float* inj_r1,
                          const float* inj_r2, float* samples, int32_t* face_idx_out, float* bary_out,
                          void* ws, size_t ws_bytes, f3d_stream_t stream);
/* Pullback of _sample_points (src/transforms/mesh_func.jl:60-73; the face draws are constants, :47 is @ignore):
 *   gverts_padded[mesh][faces[face][k]] += w_k * gsamples[mesh][s]   for the three corners k of every sample.
 * gverts_padded [Nmesh][Vmax][3] is ACCUMULATED into (zero it first); float RED.ADD, so the summation order —
 * and with it the last bit — may vary from run to run. */
F3D_API int32_t f3d_sample_points_bwd(const float* gsamples, const int32_t* face_idx, const float* bary,
                              const int32_t* faces_padded, int32_t Nmesh, int32_t Vmax, int32_t Fmax,
                              int32_t S, float* gverts_padded, f3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: the batch axis is sharded across ranks (one process per GPU); the only data-path
 * exchange is one all-reduce of the per-shard scalar loss.  NCCL is loaded lazily (dlopen), so the
 * library itself has no link-time NCCL dependency.
 *   f3d_comm_unique_id_host: rank 0 fills a 128-byte ncclUniqueId to broadcast out of band.
 *   f3d_comm_init: collective; *comm receives an opaque handle.
 *   f3d_allreduce_sum_f32: in-place ncclAllReduce(sum) of `count` floats at dev_buf, on `stream`.
 * ---------------------------------------------------------------------------------------------- */
F3D_API int32_t f3d_comm_unique_id_host(void* id128_host);
F3D_API int32_t f3d_comm_init(int32_t nranks, int32_t rank, const void* id128_host, void** comm);
F3D_API int32_t f3d_allreduce_sum_f32(void* comm, float* dev_buf, int32_t count, f3d_stream_t stream);
/* The same exchange FUSED into the chamfer kernels (no NCCL call, no extra launch on the hot path):
 *   f3d_comm_enable_p2p: collective, once per communicator — every rank allocates a mailbox (2 x nranks 8-byte words),
 *     exports it with CUDA IPC, gathers the handles over NCCL and maps its peers' mailboxes (NVLink peer access).
 *   f3d_chamfer_fwd_allreduce: f3d_chamfer_fwd on this rank's shard (B batch elements of a B_total batch); the finalize
 *     kernel's last block stores the shard loss — already divided by N*B_total / M*B_total — into its slot of every
 *     peer's mailbox as one word {step number, float bits}, waits for the nranks words in its own mailbox and adds them
 *     in rank order, so loss_dev[0] holds the WHOLE batch's loss, the same bits on every rank.  Collective: every rank
 *     must call it the same number of times.  A peer that never arrives turns the loss into NaN after 2 s.
 *   f3d_chamfer_pipe_run takes the same communicator (or NULL) as `comm`. */
F3D_API int32_t f3d_comm_enable_p2p(void* comm, f3d_stream_t stream);
F3D_API int32_t f3d_chamfer_fwd_allreduce(void* comm, const float* A, const float* Bp, int32_t B, int32_t N, int32_t M,
                                  float w1, float w2, int32_t B_total, float* loss_dev, int32_t* nnA_dev,
                                  int32_t* nnB_dev, void* ws, size_t ws_bytes, int32_t flags,
                                  f3d_stream_t stream);
F3D_API int32_t f3d_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* FLUX3D_B200_H */
