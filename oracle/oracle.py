"""ctypes binding + numpy twin of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package never does (tests/test_boundary.py greps for it).

The C functions (oracle/f3d_oracle.c) are the primary oracle; the ``np_*`` functions are an
independent numpy/scipy restatement of the same reference lines, used to cross-check the C code
(two restatements written separately agreeing bit-for-bit is the best available substitute for
running Julia, which this image does not have).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libf3d_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "f3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orc_nearest_neighbors.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.orc_nearest_neighbors.restype = None
        L.orc_chamfer_distance.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, i32p, i32p, f32p]
        L.orc_chamfer_distance.restype = C.c_float
        L.orc_chamfer_backward.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, i32p, i32p, C.c_float, f32p, f32p]
        L.orc_chamfer_backward.restype = None
        L.orc_knn_graph.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, f32p, f32p]
        L.orc_knn_graph.restype = C.c_int
        L.orc_edge_features.argtypes = [f32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_edge_features.restype = None
        L.orc_faces_areas_normals.argtypes = [f32p, i32p, C.c_int, C.c_int, f32p, f32p]
        L.orc_faces_areas_normals.restype = None
        L.orc_verts_normals.argtypes = [f32p, i32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_verts_normals.restype = None
        L.orc_edges_packed.argtypes = [i32p, C.c_int, C.c_int, i32p, i32p]
        L.orc_edges_packed.restype = C.c_int
        L.orc_laplacian_csr.argtypes = [i32p, C.c_int, C.c_int, i32p, i32p, f32p]
        L.orc_laplacian_csr.restype = None
        L.orc_laplacian_loss.argtypes = [f32p, i32p, i32p, f32p, C.c_int]
        L.orc_laplacian_loss.restype = C.c_float
        L.orc_edge_loss.argtypes = [f32p, i32p, C.c_int, C.c_float]
        L.orc_edge_loss.restype = C.c_float
        L.orc_sample_points.argtypes = [f32p, i32p, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_uint64, C.c_uint64, i32p, f32p, f32p, f32p, i32p]
        L.orc_sample_points.restype = None
        L.orc_philox_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), f32p, f32p]
        L.orc_philox_draws.restype = None
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads() -> int:
    return lib().orc_num_threads()


# ----------------------------------------------------------------------------- point clouds
def nearest_neighbors(x, y):
    """src/metrics/pcloud.jl:54-70.  x: (B,N,3), y: (B,M,3) → (nn_x (B,N), nn_y (B,M)) 0-based."""
    x, y = _f32(x), _f32(y)
    B, N, _ = x.shape
    M = y.shape[1]
    nx = np.empty((B, N), np.int32)
    ny = np.empty((B, M), np.int32)
    lib().orc_nearest_neighbors(_f(x), _f(y), B, N, M, _i(nx), _i(ny))
    return nx, ny


def chamfer_distance(A, Bc, w1=1.0, w2=1.0, return_all=False):
    """src/metrics/pcloud.jl:39-52."""
    A, Bc = _f32(A), _f32(Bc)
    if A.ndim == 2:
        A, Bc = A[None], Bc[None]
    B, N, _ = A.shape
    M = Bc.shape[1]
    assert Bc.shape[0] == B
    nA = np.empty((B, N), np.int32)
    nB = np.empty((B, M), np.int32)
    terms = np.empty(2, np.float32)
    loss = lib().orc_chamfer_distance(_f(A), _f(Bc), B, N, M, w1, w2, _i(nA), _i(nB), _f(terms))
    loss = np.float32(loss)
    return (loss, nA, nB, terms) if return_all else loss


def chamfer_backward(A, Bc, nA, nB, w1=1.0, w2=1.0, gout=1.0):
    A, Bc = _f32(A), _f32(Bc)
    B, N, _ = A.shape
    M = Bc.shape[1]
    gA = np.empty_like(A)
    gB = np.empty_like(Bc)
    lib().orc_chamfer_backward(_f(A), _f(Bc), B, N, M, w1, w2, _i(_i32(nA)), _i(_i32(nB)), gout, _f(gA), _f(gB))
    return gA, gB


def knn_graph(X, K, want_dist=False, want_gathered=False):
    """src/models/dgcnn.jl:3-7,36.  X: (B,N,F) → idx (B,N,K) [, dist (B,N,K)] [, gathered (B,N,K,F)]."""
    X = _f32(X)
    B, N, F = X.shape
    idx = np.empty((B, N, K), np.int32)
    dist = np.empty((B, N, K), np.float32) if want_dist else None
    gat = np.empty((B, N, K, F), np.float32) if want_gathered else None
    rc = lib().orc_knn_graph(_f(X), B, N, F, K, _i(idx), _f(dist), _f(gat))
    if rc:
        raise ValueError("orc_knn_graph: need 1 <= K < N")
    out = [idx]
    if want_dist:
        out.append(dist)
    if want_gathered:
        out.append(gat)
    return out[0] if len(out) == 1 else tuple(out)


def edge_features(X, idx):
    """src/models/dgcnn.jl:39-45 → (B,N,K,2F)."""
    X, idx = _f32(X), _i32(idx)
    B, N, F = X.shape
    K = idx.shape[2]
    out = np.empty((B, N, K, 2 * F), np.float32)
    lib().orc_edge_features(_f(X), _i(idx), B, N, F, K, _f(out))
    return out


# ----------------------------------------------------------------------------- meshes
def faces_areas_normals(verts, faces):
    """src/rep/mesh.jl:765-780, :689-700.  verts (V,3) f32, faces (F,3) 0-based."""
    verts, faces = _f32(verts), _i32(faces)
    nF = faces.shape[0]
    areas = np.empty(nF, np.float32)
    normals = np.empty((nF, 3), np.float32)
    lib().orc_faces_areas_normals(_f(verts), _i(faces), verts.shape[0], nF, _f(areas), _f(normals))
    return areas, normals


def verts_normals(verts, faces, mode=0):
    """src/rep/mesh.jl:589-618.  mode 0 = REFERENCE_CPU (last face per slot), 1 = ACCUMULATE."""
    verts, faces = _f32(verts), _i32(faces)
    out = np.empty_like(verts)
    lib().orc_verts_normals(_f(verts), _i(faces), verts.shape[0], faces.shape[0], mode, _f(out))
    return out


def edges_packed(faces, nV):
    """src/rep/mesh.jl:907-955 → (edges (E,2), faces_to_edges (F,3)) 0-based."""
    faces = _i32(faces)
    nF = faces.shape[0]
    edges = np.empty((max(3 * nF, 1), 2), np.int32)
    f2e = np.empty((nF, 3), np.int32)
    nE = lib().orc_edges_packed(_i(faces), nV, nF, _i(edges), _i(f2e))
    return edges[:nE].copy(), f2e


def laplacian_csr(edges, nV):
    """src/rep/mesh.jl:957-1002 as CSR (rowptr, colidx, vals)."""
    edges = _i32(edges)
    nE = edges.shape[0]
    rowptr = np.empty(nV + 1, np.int32)
    colidx = np.empty(2 * nE + nV, np.int32)
    vals = np.empty(2 * nE + nV, np.float32)
    lib().orc_laplacian_csr(_i(edges), nE, nV, _i(rowptr), _i(colidx), _f(vals))
    return rowptr, colidx, vals


def laplacian_loss(verts, faces):
    """src/metrics/mesh.jl:9-15 on packed verts/faces."""
    verts = _f32(verts)
    nV = verts.shape[0]
    edges, _ = edges_packed(faces, nV)
    rowptr, colidx, vals = laplacian_csr(edges, nV)
    return np.float32(lib().orc_laplacian_loss(_f(verts), _i(rowptr), _i(colidx), _f(vals), nV))


def edge_loss(verts, faces, target=0.0):
    """src/metrics/mesh.jl:24-32."""
    verts = _f32(verts)
    edges, _ = edges_packed(faces, verts.shape[0])
    return np.float32(lib().orc_edge_loss(_f(verts), _i(edges), edges.shape[0], target))


def sample_points(verts_padded, faces_padded, verts_len, faces_len, S, eps=1e-6, seed=0, offset=0,
                  inj_face=None, inj_r1=None, inj_r2=None):
    """src/transforms/mesh_func.jl:21-82.  Returns (samples (Nmesh,S,3), face_idx (Nmesh,S))."""
    vp, fp = _f32(verts_padded), _i32(faces_padded)
    vl, fl = _i32(verts_len), _i32(faces_len)
    Nm, Vmax, _ = vp.shape
    Fmax = fp.shape[1]
    out = np.empty((Nm, S, 3), np.float32)
    fidx = np.empty((Nm, S), np.int32)
    if inj_face is not None:
        inj_face, inj_r1, inj_r2 = _i32(inj_face), _f32(inj_r1), _f32(inj_r2)
    lib().orc_sample_points(_f(vp), _i(fp), _i(vl), _i(fl), Nm, Vmax, Fmax, S, eps, seed, offset,
                            _i(inj_face), _f(inj_r1), _f(inj_r2), _f(out), _i(fidx))
    return out, fidx


def philox_draws(seed, offset, mesh, s):
    u = C.c_uint64()
    r1 = C.c_float()
    r2 = C.c_float()
    lib().orc_philox_draws(seed, offset, mesh, s, C.byref(u), C.byref(r1), C.byref(r2))
    return u.value, r1.value, r2.value


# ----------------------------------------------------------------------------- numpy twins
def np_sqdist_matrix(x, y):
    """Direct-difference Float32 form, ((dx²)+(dy²))+(dz²) (generic F: sequential)."""
    x, y = _f32(x), _f32(y)
    d = None
    for k in range(x.shape[-1]):
        t = x[:, None, k] - y[None, :, k]
        t = t * t
        d = t if d is None else d + t
    return d


def np_nearest_neighbors(x, y):
    B = x.shape[0]
    nx, ny = [], []
    for b in range(B):
        d = np_sqdist_matrix(x[b], y[b])
        nx.append(np.argmin(d, axis=1))  # numpy argmin → first (lowest-index) minimum
        ny.append(np.argmin(d, axis=0))
    return np.stack(nx).astype(np.int32), np.stack(ny).astype(np.int32)


def np_chamfer_distance(A, Bc, w1=1.0, w2=1.0):
    """Float64-accumulated twin of src/metrics/pcloud.jl:39-52 (for tolerance checks)."""
    A, Bc = _f32(A), _f32(Bc)
    nA, nB = np_nearest_neighbors(A, Bc)
    gB = np.take_along_axis(Bc, nA[..., None].astype(np.int64), axis=1)
    gA = np.take_along_axis(A, nB[..., None].astype(np.int64), axis=1)
    dAB = np.mean(((A - gB) ** 2).astype(np.float64)) * 3.0
    dBA = np.mean(((Bc - gA) ** 2).astype(np.float64)) * 3.0
    return w1 * dAB + w2 * dBA, nA, nB


def np_naive_chamfer(x, y):
    """The reference test's own checker — test/metrics.jl:94-107 (expanded form, min not argmin).
    x: (B,N,3), y: (B,M,3); float64 to serve as a tolerance anchor."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    xx = (x ** 2).sum(-1)[:, :, None]
    yy = (y ** 2).sum(-1)[:, None, :]
    zz = np.einsum("bnd,bmd->bnm", x, y)
    P = xx + yy - 2 * zz
    return P.min(2).mean(1).mean() + P.min(1).mean(1).mean()


def np_knn_graph(X, K):
    X = _f32(X)
    B, N, F = X.shape
    out = np.empty((B, N, K), np.int32)
    for b in range(B):
        d = np_sqdist_matrix(X[b], X[b])
        order = np.lexsort((np.broadcast_to(np.arange(N), (N, N)), d), axis=1)  # by d then index
        out[b] = order[:, 1:K + 1]
    return out


def np_cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                     a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def np_norm3(c):
    return np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2])


def np_faces_areas_normals(verts, faces):
    verts, faces = _f32(verts), np.asarray(faces, np.int64)
    v1, v2, v3 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    c = np_cross(v2 - v1, v3 - v1)
    n = np_norm3(c)
    return n / np.float32(2), c / np.maximum(n, np.float32(1e-6))[:, None]


def np_verts_normals(verts, faces, mode=0):
    verts, faces = _f32(verts), np.asarray(faces, np.int64)
    vn = np.zeros_like(verts)
    for k in range(3):
        vk, va, vb = verts[faces[:, k]], verts[faces[:, (k + 1) % 3]], verts[faces[:, (k + 2) % 3]]
        c = np_cross(va - vk, vb - vk)
        if mode == 0:
            vn[faces[:, k]] = vn[faces[:, k]] + c  # numpy fancy assignment: last write wins, like Julia
        else:
            for f in range(faces.shape[0]):
                vn[faces[f, k]] = vn[faces[f, k]] + c[f]
    n = np.maximum(np_norm3(vn), np.float32(1e-6))
    return vn / n[:, None]


def np_edges_packed(faces):
    faces = np.asarray(faces, np.int64)
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    e = np.sort(e, axis=1)
    return np.unique(e, axis=0).astype(np.int32)


def np_laplacian_loss(verts, faces, dtype=np.float32):
    """scipy.sparse twin of src/rep/mesh.jl:957-1002 + src/metrics/mesh.jl:9-15."""
    import scipy.sparse as sp
    verts = np.asarray(verts, dtype)
    V = verts.shape[0]
    e = np_edges_packed(faces).astype(np.int64)
    A = sp.coo_matrix((np.ones(2 * len(e)), (np.r_[e[:, 0], e[:, 1]], np.r_[e[:, 1], e[:, 0]])), shape=(V, V)).tocsr()
    deg = np.asarray(A.sum(1)).ravel()
    inv = np.where(deg > 0, 1.0 / np.maximum(deg, 1), 0.0).astype(dtype)
    Lm = sp.diags(inv) @ A.astype(dtype) - sp.identity(V, dtype=dtype)
    Lv = Lm.astype(dtype) @ verts
    return np.mean(np.sqrt((Lv ** 2).sum(1)), dtype=dtype)


# ----------------------------------------------------------------------------- fixtures
def load_obj(path):
    """Minimal OBJ reader (v / f lines, first index of each v/vt/vn triple, fan-triangulated).
    Returns verts (V,3) f32 and faces (F,3) int32 0-based — what load_trimesh yields for the
    reference's test assets (src/rep/mesh.jl:297-325; teapot V=1202 F=2256, sphere V=2562 F=5120)."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                ids = [int(t.split("/")[0]) for t in line.split()[1:]]
                ids = [i - 1 if i > 0 else len(vs) + i for i in ids]
                for k in range(1, len(ids) - 1):
                    fs.append([ids[0], ids[k], ids[k + 1]])
    return np.asarray(vs, np.float32), np.asarray(fs, np.int32)


def kdtree_chamfer(A, Bc, w1=1.0, w2=1.0, workers=1):
    """The reference's CPU ALGORITHM (src/metrics/pcloud.jl:54-70: one KD-tree build + N 1-NN queries
    per batch element and direction, serial over the batch) with scipy's cKDTree standing in for
    NearestNeighbors.jl (leafsize 10 = NearestNeighbors' default).  Used only as the timed CPU
    baseline; scipy searches in float64, so this is not the parity oracle.
    workers=1 is the reference's behaviour (single thread); workers=-1 spreads the batch elements
    over all host cores (a thread pool; cKDTree releases the GIL) — more than the reference does."""
    from scipy.spatial import cKDTree
    A, Bc = _f32(A), _f32(Bc)
    B = A.shape[0]

    def one(b):
        _, ia = cKDTree(Bc[b], leafsize=10).query(A[b], k=1)
        _, ib = cKDTree(A[b], leafsize=10).query(Bc[b], k=1)
        return (float(np.sum((A[b] - Bc[b][ia]) ** 2, dtype=np.float64)),
                float(np.sum((Bc[b] - A[b][ib]) ** 2, dtype=np.float64)))

    if workers == 1:
        parts = [one(b) for b in range(B)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        n = os.cpu_count() if workers in (-1, None) else workers
        with ThreadPoolExecutor(max_workers=n) as ex:
            parts = list(ex.map(one, range(B)))
    sA = sum(p[0] for p in parts)
    sB = sum(p[1] for p in parts)
    return np.float32(w1 * sA / (B * A.shape[1]) + w2 * sB / (B * Bc.shape[1]))


# ---- packed / padded / list converters (src/rep/utils.jl:51-206), numpy restatement --------------------------------
def np_packed_to_padded(packed, items_len, pad_value=0):
    """_packed_to_padded — src/rep/utils.jl:131-152.  packed (ΣL, D) (== Julia (D, ΣL)) -> (N, max len, D)."""
    packed = np.asarray(packed)
    n, m = len(items_len), int(max(items_len))
    padded = np.full((n, m) + packed.shape[1:], pad_value, packed.dtype)
    cur = 0
    for i, ln in enumerate(items_len):
        padded[i, :ln] = packed[cur:cur + ln]
        cur += ln
    return padded


def np_padded_to_packed(padded, items_len):
    """_padded_to_packed — src/rep/utils.jl:168-185.  (N, W, D) -> (ΣL, D): the first items_len[i] rows of every item."""
    padded = np.asarray(padded)
    assert len(items_len) == padded.shape[0]  # utils.jl:177-178
    return np.concatenate([padded[i, :ln] for i, ln in enumerate(items_len)], axis=0)


def np_list_to_padded(lst, pad_value=0):
    """_list_to_padded — src/rep/utils.jl:51-93."""
    return np_packed_to_padded(np.concatenate(lst, axis=0), [len(x) for x in lst], pad_value)
