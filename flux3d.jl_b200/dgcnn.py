"""kNN graph of DGCNN's EdgeConv — host-side mirror of src/models/dgcnn.jl:3-9 (CreateSingleKNNGraph) and
:32-45 (the batch loop / gather / tile / concat prologue of EdgeConv).  The Conv/BN MLP and MaxPool that
follow (:46-61) are stock Flux layers (cuDNN/cuBLAS) and stay outside this library.

Layout: the reference's X (F, N, B) is the torch tensor X[b, n, f]; KNNGraph (F, K, N, B) is
gathered[b, n, k, f]; the (2F, K, N, B) edge tensor is edge[b, n, k, 0:2F]."""
from __future__ import annotations

import torch

from . import _lib
from .pcloud import as_f32_tensor


def knn_graph(X, K: int, *, want_dist: bool = False, want_gathered: bool = False, want_edge: bool = False, flags: int = 0,
              want_stats: bool = False):
    """For every point of every cloud the K nearest OTHER points, sorted ascending by (distance, index);
    the first hit of the (K+1)-list is dropped by position exactly as dgcnn.jl:6 does.
    X: (B, N, F) or (N, F).  Returns dict(idx (B,N,K) int32 [, dist (B,N,K)] [, gathered (B,N,K,F)]
    [, edge (B,N,K,2F)])."""
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
    edge = torch.empty((B, N, K, 2 * F), dtype=torch.float32, device=dev) if want_edge else None
    stats = torch.zeros(16, dtype=torch.int32, device=dev) if want_stats else None
    with torch.cuda.device(dev):
        _lib.check(L.f3d_knn_graph(_lib.ptr(X), B, N, F, K, _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(gat), _lib.ptr(edge),
                                   _lib.ptr(stats), 64 if want_stats else 0, int(flags), _lib.stream_ptr(dev)))
    out = {"idx": idx}
    if want_stats:
        out["stats"] = stats  # [queries that fell back to an exact scan, candidates re-evaluated exactly]
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


def edgeconv_features(X, K: int) -> torch.Tensor:
    """The EdgeConv prologue — dgcnn.jl:32-45: cat(X_tiled, KNNGraph - X_tiled; dims=1) as (B, N, K, 2F)
    (== Julia (2F, K, N, B)).  No gradient flows through it in the reference's KNN branch (@nograd, :9);
    the x_i half is a plain copy."""
    return knn_graph(X, K, want_edge=True)["edge"]
