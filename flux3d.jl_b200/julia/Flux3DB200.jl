# Flux3DB200.jl — the Julia side of the drop-in: methods that out-dispatch Flux3D's generic ones for
# CuArray{Float32} storage and `ccall` libflux3d_b200.so (include/flux3d_b200.h).
#
# STATUS: written against Flux3D v0.1.6 + CUDA.jl, NOT executed — the build image has no Julia
# (SURVEY.md §0 fact 2).  The same C symbols are exercised from Python/ctypes by tests/ and bench.py, and
# tests/test_julia_shim.py checks every ccall in this file (symbol, return type, argument count and C types) against
# include/flux3d_b200.h — the only verification possible without a Julia binary.
#
# Usage (a maintainer adds ONE line to src/Flux3D.jl after the includes):
#     include(joinpath(ENV["FLUX3D_B200_HOME"], "julia", "Flux3DB200.jl"))
# Layout: a Julia (3,N,B) Float32 CuArray is byte-identical to the C [B][N][3] array the library reads,
# so no copies or permutes are made.  Indices come back 0-based Int32; +1 is applied here.
module Flux3DB200

using CUDA, Zygote, NNlib
import Flux3D
import Flux3D: TriMesh, PointCloud, get_verts_packed, get_verts_padded, get_faces_packed, get_faces_padded

const LIB = get(ENV, "FLUX3D_B200_LIB", joinpath(@__DIR__, "..", "libflux3d_b200.so"))

function check(status::Int32)
    status == 0 && return
    buf = Vector{UInt8}(undef, 512)
    ccall((:f3d_last_error, LIB), Int32, (Ptr{UInt8}, Csize_t), buf, 512)
    error("libflux3d_b200 status $status: ", unsafe_string(pointer(buf)))   # same style as rep/mesh.jl:126-128
end

devptr(x::CuArray{T}) where {T} = reinterpret(Ptr{T}, pointer(x))
cur_stream() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)

const _ws = Dict{Any,CuVector{UInt8}}()
function workspace(key, nbytes)
    w = get(_ws, key, nothing)
    if w === nothing || length(w) < nbytes
        w = CuVector{UInt8}(undef, max(nbytes, 256)); _ws[key] = w
    end
    return w
end

# ---- chamfer_distance: replaces _chamfer_distance + _nearest_neighbors(::CuArray,::CuArray) -------------
# (src/metrics/pcloud.jl:39-52, :72-86)
function chamfer_fwd(A::CuArray{Float32,3}, B::CuArray{Float32,3}, w1::Float32, w2::Float32; want_indices = true, batch_total = 0)
    (_, N, Bn) = size(A); M = size(B, 2)
    size(B, 3) == Bn || error("batch sizes differ: $Bn vs $(size(B, 3))")
    nbytes = ccall((:f3d_chamfer_workspace_bytes, LIB), Csize_t, (Int32, Int32, Int32), Bn, N, M)
    ws = workspace((:chamfer, Bn, N, M), nbytes)
    res = CUDA.zeros(Float32, 3)
    nnA = want_indices ? CuArray{Int32}(undef, N, Bn) : nothing
    nnB = want_indices ? CuArray{Int32}(undef, M, Bn) : nothing
    check(ccall((:f3d_chamfer_fwd, LIB), Int32,
        (Ptr{Float32}, Ptr{Float32}, Int32, Int32, Int32, Float32, Float32, Int32, Ptr{Float32}, Ptr{Float32},
         Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}),
        devptr(A), devptr(B), Bn, N, M, w1, w2, batch_total, devptr(res), devptr(res) + 4,
        want_indices ? devptr(nnA) : C_NULL, want_indices ? devptr(nnB) : C_NULL, devptr(ws), length(ws), 0, cur_stream()))
    return res, nnA, nnB
end

function Flux3D._chamfer_distance(A::CuArray{Float32,3}, B::CuArray{Float32,3}, w1::Float32, w2::Float32)
    res, _, _ = chamfer_fwd(A, B, w1, w2; want_indices = false)
    return CUDA.@allowscalar res[1]
end

Zygote.@adjoint function Flux3D._chamfer_distance(A::CuArray{Float32,3}, B::CuArray{Float32,3}, w1::Float32, w2::Float32)
    res, nnA, nnB = chamfer_fwd(A, B, w1, w2)
    (_, N, Bn) = size(A); M = size(B, 2)
    function back(g)
        gout = CuArray(Float32[g]); gA = similar(A); gB = similar(B)
        check(ccall((:f3d_chamfer_bwd, LIB), Int32,
            (Ptr{Float32}, Ptr{Float32}, Int32, Int32, Int32, Float32, Float32, Int32, Ptr{Int32}, Ptr{Int32},
             Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Cvoid}),
            devptr(A), devptr(B), Bn, N, M, w1, w2, 0, devptr(nnA), devptr(nnB), devptr(gout), devptr(gA), devptr(gB), cur_stream()))
        return (gA, gB, nothing, nothing)
    end
    return CUDA.@allowscalar(res[1]), back
end

# ---- chamfer_distance on host Arrays (src/metrics/pcloud.jl:28-37 called with `Array`s) -----------------------
# OPT-IN: Flux3D.chamfer_distance(::Array, ::Array) itself is NOT overridden — the reference's CPU method keeps working
# (and stays differentiable on the CPU, test/metrics.jl:112-114) on machines without a B200.  A caller who wants host
# arrays swept on the GPU calls Flux3DB200.chamfer_distance_host: one C call, the sweep grid pulls the page-locked arrays
# over PCIe itself and stores the loss into mapped host memory.  Page-lock ONCE, outside the hot loop, with pin_host!.
const _pipe = Ref{Ptr{Cvoid}}(C_NULL)
function pipe_handle()
    if _pipe[] == C_NULL
        check(ccall((:f3d_chamfer_pipe_create, LIB), Int32, (Int32, Ptr{Ptr{Cvoid}}), 0, _pipe))
    end
    return _pipe[]
end

pin_host!(A::Array{Float32,3}) = (CUDA.pin(A); A)   # cudaHostRegister: lets the device read the array in place

function chamfer_distance_host(A::Array{Float32,3}, B::Array{Float32,3}; w1::Number = 1.0, w2::Number = 1.0)
    (_, N, Bn) = size(A); M = size(B, 2)
    size(B, 3) == Bn || error("batch sizes differ: $Bn vs $(size(B, 3))")
    # pageable arrays also work — the library then copies them with cudaMemcpyAsync first
    nbytes = ccall((:f3d_chamfer_pipe_workspace_bytes, LIB), Csize_t, (Int32, Int32, Int32), Bn, N, M)
    ws = workspace((:chamfer_pipe, Bn, N, M), nbytes)
    out = Ref{Float32}(0f0)
    GC.@preserve A B begin
        check(ccall((:f3d_chamfer_pipe_run, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Int32, Int32, Float32, Float32, Int32, Ptr{Float32}, Ptr{Float32},
             Ptr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}, Ptr{Cvoid}),
            pipe_handle(), pointer(A), pointer(B), Bn, N, M, Float32(w1), Float32(w2), 0, C_NULL, out,
            devptr(ws), length(ws), 0, #=comm=# C_NULL, cur_stream()))
    end
    return out[]   # a host Float32, like the reference on Arrays
end

# ---- batch sharded over GPUs (one Julia process per GPU): the shard losses are summed INSIDE the finalize kernel ----
# comm = comm_init(nranks, rank, id) once (id: the 128 bytes of f3d_comm_unique_id_host from rank 0, broadcast e.g. with MPI)
function comm_init(nranks::Integer, rank::Integer, id::Vector{UInt8})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:f3d_comm_init, LIB), Int32, (Int32, Int32, Ptr{UInt8}, Ptr{Ptr{Cvoid}}), nranks, rank, id, h))
    check(ccall((:f3d_comm_enable_p2p, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h[], cur_stream()))   # peer mailboxes over NVLink
    return h[]
end

# this rank's shard (3,N,b_local) / (3,M,b_local) of a batch of batch_total elements -> the loss of the WHOLE batch
function chamfer_distance_sharded(comm::Ptr{Cvoid}, A::CuArray{Float32,3}, B::CuArray{Float32,3}, batch_total::Integer;
                                  w1::Float32 = 1f0, w2::Float32 = 1f0)
    (_, N, Bn) = size(A); M = size(B, 2)
    nbytes = ccall((:f3d_chamfer_workspace_bytes, LIB), Csize_t, (Int32, Int32, Int32), Bn, N, M)
    ws = workspace((:chamfer, Bn, N, M), nbytes)
    loss = CUDA.zeros(Float32, 1)
    check(ccall((:f3d_chamfer_fwd_allreduce, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Int32, Int32, Int32, Float32, Float32, Int32, Ptr{Float32}, Ptr{Int32}, Ptr{Int32},
         Ptr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}),
        comm, devptr(A), devptr(B), Bn, N, M, w1, w2, batch_total, devptr(loss), C_NULL, C_NULL, devptr(ws), length(ws), 0, cur_stream()))
    return CUDA.@allowscalar loss[1]
end

# _nearest_neighbors(::CuArray, ::CuArray) — src/metrics/pcloud.jl:72-86: CartesianIndex matrices (N,B),(M,B)
function Flux3D._nearest_neighbors(x::CuArray{Float32,3}, y::CuArray{Float32,3})
    _, nnx, nny = chamfer_fwd(x, y, 1f0, 1f0)
    hx, hy = Array(nnx), Array(nny)
    return ([CartesianIndex(Int(hx[i, b]) + 1, b) for i in axes(hx, 1), b in axes(hx, 2)],
            [CartesianIndex(Int(hy[j, b]) + 1, b) for j in axes(hy, 1), b in axes(hy, 2)])
end

# ---- kNN graph: replaces CreateSingleKNNGraph / the EdgeConv prologue (src/models/dgcnn.jl:3-9, 32-45) ----
const FLAG_EDGE_MLP_LAYOUT = Int32(32)
function knn_graph(X::CuArray{Float32,3}, K::Int; gathered = false, edge = false, mlp_layout = false)
    (F, N, Bn) = size(X)
    idx = CuArray{Int32}(undef, K, N, Bn)
    G = gathered ? CuArray{Float32}(undef, F, K, N, Bn) : nothing
    # edge features: (2F, K, N, B) == cat(X, KNNGraph - X; dims=1) (:45), or — mlp_layout — already permuted and reshaped to
    # the (K*N, 2F, B) array the Conv1x1 MLP consumes (:46-52): C [B][2F][N][K]
    E = edge ? (mlp_layout ? CuArray{Float32}(undef, K * N, 2F, Bn) : CuArray{Float32}(undef, 2F, K, N, Bn)) : nothing
    # the workspace takes the cloud's tensor-core operand image (knn_gram.cu); without it the library stays on slower kernels
    nbytes = ccall((:f3d_knn_graph_workspace_bytes, LIB), Csize_t, (Int32, Int32, Int32, Int32), Bn, N, F, K)
    ws = workspace((:knn, Bn, N, F, K), nbytes)
    check(ccall((:f3d_knn_graph, LIB), Int32,
        (Ptr{Float32}, Int32, Int32, Int32, Int32, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}),
        devptr(X), Bn, N, F, K, devptr(idx), C_NULL, gathered ? devptr(G) : C_NULL, edge ? devptr(E) : C_NULL, devptr(ws), nbytes,
        (edge && mlp_layout) ? FLAG_EDGE_MLP_LAYOUT : Int32(0), cur_stream()))
    return idx, G, E
end

Flux3D.CreateSingleKNNGraph(X::CuArray{Float32,2}, K::Int) =
    dropdims(knn_graph(reshape(X, size(X, 1), size(X, 2), 1), K; gathered = true)[2]; dims = 4)
Zygote.@nograd knn_graph

# (K*N, 2F, B) edge features from X and the (constant) neighbour indices.  Differentiable in X like the reference, where
# only CreateSingleKNNGraph is @nograd (:9) and X flows through cat(X, KNNGraph - X) (:39-45):
#   E[(k,n), c, b]     = X[c, n, b]                      (c <= F)
#   E[(k,n), F + c, b] = X[c, idx[k,n,b], b] - X[c, n, b]
edge_features_mlp(X::CuArray{Float32,3}, K::Int) = knn_graph(X, K; edge = true, mlp_layout = true)[3]
Zygote.@adjoint function edge_features_mlp(X::CuArray{Float32,3}, K::Int)
    (F, N, Bn) = size(X)
    idx, _, E = knn_graph(X, K; edge = true, mlp_layout = true)
    function back(g)
        g4 = reshape(CuArray{Float32}(g), K, N, 2F, Bn)                    # (k, n, c, b)
        gc = permutedims(dropdims(sum(g4[:, :, 1:F, :]; dims = 1); dims = 1), (2, 1, 3))          # centre halves: (F, N, B)
        gd = g4[:, :, F+1:2F, :]                                            # (k, n, c, b): gradient of x_j - x_i
        gX = gc .- permutedims(dropdims(sum(gd; dims = 1); dims = 1), (2, 1, 3))
        # scatter-add of the neighbour halves: gX[:, idx[k,n,b], b] += gd[k, n, :, b]
        lin = vec(Int.(idx) .+ 1 .+ reshape((0:Bn-1) .* N, 1, 1, Bn))                              # (K*N*B,) columns of reshape(gX, F, N*B)
        vals = reshape(permutedims(gd, (3, 1, 2, 4)), F, K * N * Bn)
        gXf = reshape(gX, F, N * Bn)
        NNlib.scatter!(+, gXf, vals, lin)
        return (reshape(gXf, F, N, Bn), nothing)
    end
    return E, back
end

# EdgeConv on device arrays: the prologue AND the permute/reshape (:36-52) are one kernel — the (K*N, 2F, B) array is
# written once, in the layout the MLP reads; the MLP/MaxPool tail (:54-68) is unchanged.
function (m::Flux3D.EdgeConv)(X::CuArray{Float32,3})
    F, N, B = size(X)
    Xe = edge_features_mlp(X, m.K)
    Xe = m.mlp(Xe)
    an = size(Xe, 2)
    Xe = reshape(Xe, m.K, an * N, B)
    Xe = m.maxpool_K(Xe)
    Xe = reshape(Xe, N, an, B)
    return permutedims(Xe, (2, 1, 3))
end

# ---- TriMesh: device-resident faces/topology cached per mesh (the reference caches the same products on the
# host, src/rep/mesh.jl:93-97) ---------------------------------------------------------------------------
struct Topology
    faces::CuArray{Int32,2}; edges::CuArray{Int32,2}; nE::Int
    rowptr::CuVector{Int32}; colidx::CuVector{Int32}; vals::CuVector{Float32}
    v2c_rowptr::CuVector{Int32}; v2c::CuVector{Int32}
end
const _topo = WeakKeyDict{Any,Topology}()

function topology(m::TriMesh)
    get!(_topo, m._faces_list) do
        faces = Int32.(get_faces_packed(m)) .- Int32(1)                  # (3, ΣF) == C [ΣF][3], 0-based
        nV = size(get_verts_packed(m), 2); nF = size(faces, 2)
        edges = Matrix{Int32}(undef, 2, max(3nF, 1)); nE = Ref{Int32}(0)
        rowptr = Vector{Int32}(undef, nV + 1); colidx = Vector{Int32}(undef, 6nF + nV); vals = Vector{Float32}(undef, 6nF + nV)
        v2c_rowptr = Vector{Int32}(undef, nV + 1); v2c = Vector{Int32}(undef, max(3nF, 1))
        check(ccall((:f3d_mesh_topology_build_host, LIB), Int32,
            (Ptr{Int32}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Ptr{Int32}, Ptr{Int32}),
            faces, nV, nF, edges, nE, C_NULL, rowptr, colidx, vals, v2c_rowptr, v2c))
        nnz = 2 * nE[] + nV
        Topology(CuArray(faces), CuArray(edges[:, 1:nE[]]), nE[], CuArray(rowptr), CuArray(colidx[1:nnz]), CuArray(vals[1:nnz]),
                 CuArray(v2c_rowptr), CuArray(v2c))
    end
end

# laplacian_loss — src/metrics/mesh.jl:9-15 (the reference copies verts to the host and runs a CPU SpMM).  The kernel works on
# the packed verts; get_verts_packed stays on the Zygote tape (examples/fit_mesh.jl:78-84 differentiates through it), the
# array-level function below carries the adjoint.
function _laplacian_loss_dev(verts::CuArray{Float32,2}, t::Topology)
    nV = size(verts, 2)
    ws = workspace((:lap, nV), ccall((:f3d_laplacian_workspace_bytes, LIB), Csize_t, (Int32,), nV))
    loss = CUDA.zeros(Float32, 1)
    check(ccall((:f3d_laplacian_loss, LIB), Int32,
        (Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Int32, Int32, Ptr{Float32}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
        devptr(verts), devptr(t.rowptr), devptr(t.colidx), devptr(t.vals), nV, 0, devptr(loss), devptr(ws), length(ws), cur_stream()))
    return CUDA.@allowscalar loss[1]
end
Zygote.@adjoint function _laplacian_loss_dev(verts::CuArray{Float32,2}, t::Topology)
    function back(g)
        nV = size(verts, 2)
        ws = workspace((:lap, nV), ccall((:f3d_laplacian_workspace_bytes, LIB), Csize_t, (Int32,), nV))
        gout = CuArray(Float32[g]); gv = similar(verts)
        check(ccall((:f3d_laplacian_loss_bwd, LIB), Int32,
            (Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float32}, Int32, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
            devptr(verts), devptr(t.rowptr), devptr(t.colidx), devptr(t.vals), nV, 0, devptr(gout), devptr(gv), devptr(ws), length(ws), cur_stream()))
        return (gv, nothing)
    end
    return _laplacian_loss_dev(verts, t), back
end
Flux3D.laplacian_loss(m::TriMesh{Float32,R,CuArray}) where {R} =
    _laplacian_loss_dev(get_verts_packed(m), Zygote.ignore(() -> topology(m)))

# edge_loss — src/metrics/mesh.jl:24-32: mean_e (|v_e1 - v_e2| - target)^2 over the unique edges
function _edge_loss_dev(verts::CuArray{Float32,2}, t::Topology, target::Float32)
    ws = workspace((:edge, t.nE), ccall((:f3d_edge_loss_workspace_bytes, LIB), Csize_t, (Int32,), t.nE))
    loss = CUDA.zeros(Float32, 1)
    check(ccall((:f3d_edge_loss, LIB), Int32,
        (Ptr{Float32}, Ptr{Int32}, Int32, Int32, Float32, Ptr{Float32}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
        devptr(verts), devptr(t.edges), t.nE, 0, target, devptr(loss), devptr(ws), length(ws), cur_stream()))
    return CUDA.@allowscalar loss[1]
end
Zygote.@adjoint function _edge_loss_dev(verts::CuArray{Float32,2}, t::Topology, target::Float32)
    function back(g)
        gout = CuArray(Float32[g]); gv = similar(verts)
        check(ccall((:f3d_edge_loss_bwd, LIB), Int32,
            (Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Float32, Ptr{Float32}, Ptr{Float32}, Ptr{Cvoid}),
            devptr(verts), devptr(t.rowptr), devptr(t.colidx), size(verts, 2), t.nE, target, devptr(gout), devptr(gv), cur_stream()))
        return (gv, nothing, nothing)
    end
    return _edge_loss_dev(verts, t, target), back
end
Flux3D.edge_loss(m::TriMesh{Float32,R,CuArray}, target_length::Number = 0.0) where {R} =
    _edge_loss_dev(get_verts_packed(m), Zygote.ignore(() -> topology(m)), Float32(target_length))

# compute_faces_normals_packed / compute_faces_areas_packed — src/rep/mesh.jl:689-700, 765-780: one launch for both
function faces_areas_normals(m::TriMesh{Float32,R,CuArray}; areas::Bool, normals::Bool) where {R}
    t = topology(m); verts = get_verts_packed(m); nF = size(t.faces, 2)
    a = areas ? CuArray{Float32}(undef, nF) : nothing
    n = normals ? CuArray{Float32}(undef, 3, nF) : nothing
    check(ccall((:f3d_faces_areas_normals, LIB), Int32,
        (Ptr{Float32}, Ptr{Int32}, Int32, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Cvoid}),
        devptr(verts), devptr(t.faces), size(verts, 2), nF, areas ? devptr(a) : C_NULL, normals ? devptr(n) : C_NULL, cur_stream()))
    return a, n
end
Flux3D.compute_faces_normals_packed(m::TriMesh{Float32,R,CuArray}) where {R} = faces_areas_normals(m; areas = false, normals = true)[2]
Flux3D.compute_faces_areas_packed(m::TriMesh{Float32,R,CuArray}; eps::Number = 1e-6) where {R} = faces_areas_normals(m; areas = true, normals = false)[1]

# compute_verts_normals_packed — src/rep/mesh.jl:589-618.  mode 0 = what the reference computes on the CPU
# (last face per corner slot), mode 1 = what its docstring says (sum over incident faces).
function Flux3D.compute_verts_normals_packed(m::TriMesh{Float32,R,CuArray}; mode::Integer = 0) where {R}
    t = topology(m); verts = get_verts_packed(m); out = similar(verts)
    check(ccall((:f3d_verts_normals, LIB), Int32,
        (Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Int32, Ptr{Float32}, Ptr{Cvoid}),
        devptr(verts), devptr(t.faces), devptr(t.v2c_rowptr), devptr(t.v2c), size(verts, 2), size(t.faces, 2), mode, devptr(out), cur_stream()))
    return out
end

# sample_points — src/transforms/mesh_func.jl:21-58: one launch for the whole batch, device RNG (Philox).  The padded verts
# stay on the Zygote tape (get_verts_padded); the array-level function carries the adjoint (the face draws are constants,
# :47 is @ignore): gverts[:, faces[k, face_s], mesh] += w_k * gsamples[:, s, mesh].
# `counter` (a 1-element CuVector{UInt64}): the draw counter lives on the device — its value is added to `offset` when the kernel
# runs and it is bumped after the draws, so a step recorded with CUDA.capture draws fresh samples at every replay (the reference's
# fit_mesh samples anew in every iteration, examples/fit_mesh.jl:78-84).
function _sample_points_dev(verts::CuArray{Float32,3}, faces::CuArray{Int32,3}, vlen::CuVector{Int32}, flen::CuVector{Int32},
                            num_samples::Int, eps::Float64, seed::UInt64, offset::UInt64; want_aux::Bool = false,
                            counter::Union{Nothing,CuVector{UInt64}} = nothing)
    (_, V, Nm) = size(verts); F = size(faces, 2)
    samples = similar(verts, 3, num_samples, Nm)
    fidx = want_aux ? CuArray{Int32}(undef, num_samples, Nm) : nothing
    bary = want_aux ? CuArray{Float32}(undef, 3, num_samples, Nm) : nothing
    nws = ccall((:f3d_sample_points_workspace_bytes, LIB), Csize_t, (Int32, Int32), Nm, F)
    ws = workspace((:sample, Nm, F), nws)
    if counter === nothing
        check(ccall((:f3d_sample_points, LIB), Int32,
            (Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Int32, Int32, Float64, UInt64, UInt64,
             Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Float32}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
            devptr(verts), devptr(faces), devptr(vlen), devptr(flen), Nm, V, F, num_samples, eps, seed, offset,
            C_NULL, C_NULL, C_NULL, devptr(samples), want_aux ? devptr(fidx) : C_NULL, want_aux ? devptr(bary) : C_NULL,
            devptr(ws), length(ws), cur_stream()))
    else
        check(ccall((:f3d_sample_points_replayable, LIB), Int32,
            (Ptr{Float32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Int32, Int32, Float64, UInt64, UInt64, Ptr{UInt64},
             Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Int32}, Ptr{Float32}, Ptr{Cvoid}, Csize_t, Ptr{Cvoid}),
            devptr(verts), devptr(faces), devptr(vlen), devptr(flen), Nm, V, F, num_samples, eps, seed, offset, devptr(counter),
            C_NULL, C_NULL, C_NULL, devptr(samples), want_aux ? devptr(fidx) : C_NULL, want_aux ? devptr(bary) : C_NULL,
            devptr(ws), length(ws), cur_stream()))
    end
    return samples, fidx, bary
end
Zygote.@adjoint function _sample_points_dev(verts::CuArray{Float32,3}, faces::CuArray{Int32,3}, vlen::CuVector{Int32}, flen::CuVector{Int32},
                                           num_samples::Int, eps::Float64, seed::UInt64, offset::UInt64)
    samples, fidx, bary = _sample_points_dev(verts, faces, vlen, flen, num_samples, eps, seed, offset; want_aux = true)
    (_, V, Nm) = size(verts); F = size(faces, 2)
    function back(g)
        gs = CuArray{Float32}(g[1]); gv = CUDA.zeros(Float32, 3, V, Nm)      # accumulated into: zeroed first
        check(ccall((:f3d_sample_points_bwd, LIB), Int32,
            (Ptr{Float32}, Ptr{Int32}, Ptr{Float32}, Ptr{Int32}, Int32, Int32, Int32, Int32, Ptr{Float32}, Ptr{Cvoid}),
            devptr(gs), devptr(fidx), devptr(bary), devptr(faces), Nm, V, F, num_samples, devptr(gv), cur_stream()))
        return (gv, nothing, nothing, nothing, nothing, nothing, nothing, nothing)
    end
    return (samples, nothing, nothing), back
end
function Flux3D.sample_points(m::TriMesh{Float32,R,CuArray}, num_samples::Int = 5000; eps::Number = 1e-6,
                              seed::UInt64 = rand(UInt64), offset::UInt64 = UInt64(0)) where {R}
    verts = get_verts_padded(m)                                           # (3, V, N), differentiable
    faces, vlen, flen = Zygote.ignore() do
        (CuArray(Int32.(get_faces_padded(m)) .- Int32(1)),                # (3, F, N) local ids, pad = -1
         CuArray(Int32.(m._verts_len)), CuArray(Int32.(m._faces_len)))
    end
    return _sample_points_dev(verts, faces, vlen, flen, num_samples, Float64(eps), seed, offset)[1]
end

# _packed_to_padded / _padded_to_packed — src/rep/utils.jl:131-185 — on the device (one launch each, no host loop).
# Julia (D, ΣL) / (D, W, N) arrays are the C [ΣL][D] / [N][W][D] arrays the kernels read.
item_offsets(items_len) = CuArray(Int32.(vcat(0, cumsum(collect(items_len)))))

function Flux3D._packed_to_padded(packed::CuArray{Float32,2}, items_len::AbstractArray{<:Number,1}, pad_value::Number)
    D = size(packed, 1); N = length(items_len); W = Int(maximum(items_len))
    padded = similar(packed, D, W, N); offs = item_offsets(items_len)
    check(ccall((:f3d_packed_to_padded, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Int32, UInt32, Ptr{Cvoid}, Ptr{Cvoid}),
        devptr(packed), devptr(offs), C_NULL, N, W, D, reinterpret(UInt32, Float32(pad_value)), devptr(padded), cur_stream()))
    return padded
end

function Flux3D._padded_to_packed(padded::CuArray{Float32,3}, items_len::AbstractArray{<:Number,1}, pad_value::Nothing = nothing)
    (D, W, N) = size(padded)
    N == length(items_len) || error("items_len length should match the last dimension of padded array")   # utils.jl:177-178
    total = Int(sum(items_len)); packed = similar(padded, D, total); offs = item_offsets(items_len)
    check(ccall((:f3d_padded_to_packed, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}),
        devptr(padded), devptr(offs), C_NULL, N, W, D, total, devptr(packed), cur_stream()))
    return packed
end

# each is the other's pullback (pads receive / contribute zero), as in test/rep.jl:455-458, 484-488
Zygote.@adjoint Flux3D._packed_to_padded(packed::CuArray{Float32,2}, items_len::AbstractArray{<:Number,1}, pad_value::Number) =
    Flux3D._packed_to_padded(packed, items_len, pad_value), g -> (Flux3D._padded_to_packed(CuArray{Float32,3}(g), items_len), nothing, nothing)
Zygote.@adjoint Flux3D._padded_to_packed(padded::CuArray{Float32,3}, items_len::AbstractArray{<:Number,1}, pad_value::Nothing) =
    Flux3D._padded_to_packed(padded, items_len), g -> (Flux3D._packed_to_padded(CuArray{Float32,2}(g), items_len, 0), nothing, nothing)

end # module
