"""kNN graph of DGCNN's EdgeConv — host-side mirror of src/models/dgcnn.jl:3-9 (CreateSingleKNNGraph) and
:32-45 (the batch loop / gather / tile / concat prologue of EdgeConv).  The Conv/BN MLP and MaxPool that
follow (:46-61) are stock Flux layers (cuDNN/cuBLAS) and stay outside this library.

Layout: the reference's X (F, N, B) is the torch tensor X[b, n, f]; KNNGraph (F, K, N, B) is
gathered[b, n, k, f]; the (2F, K, N, B) edge tensor is edge[b, n, k, 0:2F]."""
from __future__ import annotations

import torch

from . import _lib
from .pcloud import as_f32_tensor


FLAG_EDGE_MLP_LAYOUT = 32  # edge features as (B, 2F, N, K) == Julia (K*N, 2F, B): what EdgeConv's MLP consumes (dgcnn.jl:46-52)


def knn_graph(X, K: int, *, want_dist: bool = False, want_gathered: bool = False, want_edge: bool = False, flags: int = 0,
              want_stats: bool = False, mlp_layout: bool = False):
    """For every point of every cloud the K nearest OTHER points, sorted ascending by (distance, index);
    the first hit of the (K+1)-list is dropped by position exactly as dgcnn.jl:6 does.
    X: (B, N, F) or (N, F).  Returns dict(idx (B,N,K) int32 [, dist (B,N,K)] [, gathered (B,N,K,F)]
    [, edge (B,N,K,2F) — or, with mlp_layout, (B,2F,N,K): the permuted + reshaped (K*N, 2F, B) array of dgcnn.jl:46-52])."""
    L = _lib.lib()
    X = as_f32_tensor(X, None if (isinstance(X, torch.Tensor) and X.is_cuda) else "cuda")
    if X.dim() == 2:
        X = X.unsqueeze(0)
    if X.dim() != 3:
        raise ValueError("X must be (N, F) or (B, N, F)")
    B, N, F = X.shape
    dev = X.device
    idx = torch.empty((B, N, K), dtype=torch.int32, device=dev)
    dist = torch.empty((B, N, K), dtype=torch.float32, device=dev) if want_dist else None
    gat = torch.empty((B, N, K, F), dtype=torch.float32, device=dev) if want_gathered else None
    edge = torch.empty((B, 2 * F, N, K) if mlp_layout else (B, N, K, 2 * F), dtype=torch.float32, device=dev) if want_edge else None
    if want_edge and mlp_layout:
        flags = int(flags) | FLAG_EDGE_MLP_LAYOUT
    with torch.cuda.device(dev):
        # the workspace holds the cloud's tensor-core operand image (knn_gram.cu); its first words are diagnostics
        nws = int(L.f3d_knn_graph_workspace_bytes(B, N, F, K))
        ws = _lib.workspace(("knn", B, N, F, K), nws, dev)
        if want_stats:
            ws[:64].zero_()
        _lib.check(L.f3d_knn_graph(_lib.ptr(X), B, N, F, K, _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(gat), _lib.ptr(edge),
                                   _lib.ptr(ws), nws, int(flags), _lib.stream_ptr(dev)))
    out = {"idx": idx}
    if want_stats:
        out["stats"] = ws[:64].view(torch.int32).clone()  # [queries that fell back to an exact scan, candidates re-evaluated exactly]
    if want_dist:
        out["dist"] = dist
    if want_gathered:
        out["gathered"] = gat
    if want_edge:
        out["edge"] = edge
    return out


def create_single_knn_graph(X, K: int) -> torch.Tensor:
    """CreateSingleKNNGraph(X, K) — dgcnn.jl:3-7.  X: (N, F) (== Julia (F, N)) → (N, K, F) (== Julia (F, K, N))."""
    return knn_graph(X, K, want_gathered=True)["gathered"][0]


class _EdgeFeatFn(torch.autograd.Function):
    """cat(X, KNNGraph - X; dims=1) with the neighbour indices held constant.  In the reference only CreateSingleKNNGraph is
    @nograd (dgcnn.jl:9): X itself stays differentiable through both halves (dgcnn.jl:39-45) —
    gX[i] += sum_k (g_centre[i,k] - g_diff[i,k]),  gX[idx[i,k]] += g_diff[i,k]."""

    @staticmethod
    def forward(ctx, X, K, mlp_layout):
        out = knn_graph(X, K, want_edge=True, mlp_layout=mlp_layout)
        ctx.save_for_backward(out["idx"])
        ctx.cfg = (X.shape, mlp_layout)
        ctx.mark_non_differentiable(out["idx"])
        return out["edge"], out["idx"]

    @staticmethod
    def backward(ctx, g, _gidx):
        (idx,) = ctx.saved_tensors
        (B, N, F), mlp_layout = ctx.cfg
        K = idx.shape[2]
        g = g.permute(0, 2, 3, 1) if mlp_layout else g            # -> (B, N, K, 2F)
        gc, gd = g[..., :F], g[..., F:]
        gX = (gc - gd).sum(dim=2)                                  # the x_i halves
        flat = (idx.long() + (torch.arange(B, device=idx.device) * N).view(B, 1, 1)).reshape(-1)
        gX = gX.reshape(B * N, F).index_add(0, flat, gd.reshape(-1, F))   # the neighbours' x_j
        return gX.reshape(B, N, F), None, None


def edgeconv_features(X, K: int, *, mlp_layout: bool = False) -> torch.Tensor:
    """The EdgeConv prologue — dgcnn.jl:32-45: cat(X_tiled, KNNGraph - X_tiled; dims=1) as (B, N, K, 2F)
    (== Julia (2F, K, N, B)); with mlp_layout=True as (B, 2F, N, K) == the (K*N, 2F, B) array the 1x1-conv MLP reads
    (dgcnn.jl:46-52), written once by the kernel instead of permuted afterwards.  The neighbour indices carry no gradient
    (@nograd, :9); X does, through both halves, exactly as in the reference."""
    Xt = as_f32_tensor(X, None if (isinstance(X, torch.Tensor) and X.is_cuda) else "cuda")
    if Xt.dim() == 2:
        Xt = Xt.unsqueeze(0)
    return _EdgeFeatFn.apply(Xt, K, mlp_layout)[0]
