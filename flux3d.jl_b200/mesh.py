"""TriMesh — mirror of the reference struct (src/rep/mesh.jl:70-98) and of the mesh functions on the hot
path: compute_verts_normals_* (:589-670), compute_faces_normals_* (:689-745), compute_faces_areas_*
(:765-836), get_edges_packed / get_faces_to_edges_packed / get_laplacian_packed (:907-1002),
laplacian_loss / edge_loss (src/metrics/mesh.jl:9-32).

Layout.  The reference keeps verts as Julia (3, V) arrays (list), (3, ΣV) (packed) and (3, V, N)
(padded); the same bytes row-major are (V, 3), (ΣV, 3) and (N, V, 3), which are the torch shapes here.
Faces are integer arrays that live on the HOST in the reference (rep/mesh.jl:87-89) and are uploaded
on every use; here they are uploaded ONCE and cached on the device next to the topology products
(edges, Laplacian CSR, vertex→corner CSR), mirroring the cached fields :93-97.  Indices are 0-based
(the reference is 1-based); padded faces hold LOCAL ids and are padded with -1 (reference: 0).

All arithmetic happens in libflux3d_b200.so; torch is used for memory, slicing and autograd plumbing."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .pcloud import as_f32_tensor

NORMALS_REFERENCE_CPU = 0  # last face per corner slot wins: what rep/mesh.jl:604-615 computes on the CPU
NORMALS_ACCUMULATE = 1     # sum over all incident corners: what its docstring says


def _packed_to_padded_raw(packed: torch.Tensor, offsets: torch.Tensor, N: int, W: int, fill_bits: int = 0, delta=None) -> torch.Tensor:
    """f3d_packed_to_padded on a contiguous 4-byte-element tensor (rows, D...) -> (N, W, D...)."""
    L = _lib.lib()
    packed = packed.contiguous()
    D = int(np.prod(packed.shape[1:])) if packed.dim() > 1 else 1
    out = torch.empty((N, W) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    with torch.cuda.device(packed.device):
        _lib.check(L.f3d_packed_to_padded(_lib.ptr(packed), _lib.ptr(offsets), _lib.ptr(delta), N, W, D, fill_bits,
                                          _lib.ptr(out), _lib.stream_ptr(packed.device)))
    return out


def _padded_to_packed_raw(padded: torch.Tensor, offsets: torch.Tensor, total_rows: int, delta=None) -> torch.Tensor:
    """f3d_padded_to_packed on a contiguous 4-byte-element tensor (N, W, D...) -> (total_rows, D...)."""
    L = _lib.lib()
    padded = padded.contiguous()
    N, W = padded.shape[0], padded.shape[1]
    D = int(np.prod(padded.shape[2:])) if padded.dim() > 2 else 1
    out = torch.empty((total_rows,) + tuple(padded.shape[2:]), dtype=padded.dtype, device=padded.device)
    with torch.cuda.device(padded.device):
        _lib.check(L.f3d_padded_to_packed(_lib.ptr(padded), _lib.ptr(offsets), _lib.ptr(delta), N, W, D, total_rows,
                                          _lib.ptr(out), _lib.stream_ptr(padded.device)))
    return out


class _PackedToPadded(torch.autograd.Function):
    """_packed_to_padded (src/rep/utils.jl:131-152) with zero fill; the pullback is _padded_to_packed of the gradient."""

    @staticmethod
    def forward(ctx, packed, offsets, N, W):
        ctx.offsets, ctx.rows = offsets, packed.shape[0]
        return _packed_to_padded_raw(packed, offsets, N, W, 0)

    @staticmethod
    def backward(ctx, g):
        return _padded_to_packed_raw(g, ctx.offsets, ctx.rows), None, None, None


class _PaddedToPacked(torch.autograd.Function):
    """_padded_to_packed (src/rep/utils.jl:168-185); the pullback scatters the gradient back, pads get zero."""

    @staticmethod
    def forward(ctx, padded, offsets, total_rows):
        ctx.offsets, ctx.N, ctx.W = offsets, padded.shape[0], padded.shape[1]
        return _padded_to_packed_raw(padded, offsets, total_rows)

    @staticmethod
    def backward(ctx, g):
        return _packed_to_padded_raw(g, ctx.offsets, ctx.N, ctx.W, 0), None, None


def packed_to_padded(packed: torch.Tensor, items_len, pad_value: float = 0.0) -> torch.Tensor:
    """_packed_to_padded(packed, items_len, pad_value) — src/rep/utils.jl:131-152 — on the device: packed (ΣL, D) float32
    CUDA tensor (== Julia (D, ΣL)), items_len a list of lengths -> (N, max len, D) (== Julia (D, max len, N))."""
    lens = [int(x) for x in items_len]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32), device=packed.device)
    if pad_value == 0.0:
        return _PackedToPadded.apply(packed, offs, len(lens), max(lens))
    bits = int(np.float32(pad_value).view(np.uint32))
    return _packed_to_padded_raw(packed, offs, len(lens), max(lens), bits)


def padded_to_packed(padded: torch.Tensor, items_len) -> torch.Tensor:
    """_padded_to_packed(padded, items_len) — src/rep/utils.jl:168-185 — on the device: (N, W, D) -> (ΣL, D)."""
    lens = [int(x) for x in items_len]
    if len(lens) != padded.shape[0]:
        raise ValueError("items_len length should match the first dimension of the padded array")  # utils.jl:177-178
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]).astype(np.int32), device=padded.device)
    return _PaddedToPacked.apply(padded, offs, int(sum(lens)))


def _as_faces(f) -> np.ndarray:
    if isinstance(f, torch.Tensor):
        f = f.detach().cpu().numpy()
    f = np.ascontiguousarray(np.asarray(f), dtype=np.int64)
    if f.ndim != 2 or f.shape[1] != 3:
        raise ValueError("faces must be (F, 3)")
    return f.astype(np.int32)


class TriMesh:
    """TriMesh(verts_list, faces_list; offset=-1) — src/rep/mesh.jl:119-172.

    verts_list: list of (V_i, 3) float arrays/tensors; faces_list: list of (F_i, 3) integer arrays with
    0-based LOCAL vertex ids.  A single (V,3)/(F,3) pair is a batch of one (:186)."""

    def __init__(self, verts_list, faces_list, offset: int = -1, device="cuda"):
        if not isinstance(verts_list, (list, tuple)):
            verts_list, faces_list = [verts_list], [faces_list]
        if len(verts_list) != len(faces_list):  # rep/mesh.jl:126-128
            raise ValueError(f"batch size of verts and faces should match, {len(verts_list)} != {len(faces_list)}")
        if len(verts_list) == 0:
            raise ValueError("empty mesh batch")
        self.device = torch.device(device)
        verts = [as_f32_tensor(v, self.device) for v in verts_list]
        for v in verts:
            if v.dim() != 2 or v.shape[1] != 3:
                raise ValueError("verts must be (V, 3)")
        self._faces_list = [_as_faces(f) for f in faces_list]
        self._verts_len = [int(v.shape[0]) for v in verts]
        self._faces_len = [int(f.shape[0]) for f in self._faces_list]
        for f, nv in zip(self._faces_list, self._verts_len):
            if f.size and (f.min() < 0 or f.max() >= nv):
                raise ValueError("face index out of range")
        self.N = len(verts)
        self.V = max(self._verts_len)
        self.F = max(self._faces_len)
        self.equalised = all(v == self.V for v in self._verts_len) and all(f == self.F for f in self._faces_len)
        self.valid = [f > 0 for f in self._faces_len]
        self.offset = int(offset)
        # verts: the packed tensor is the primary (differentiable) storage
        self._verts_packed = verts[0] if self.N == 1 else torch.cat(verts, dim=0)
        self._verts_padded = None
        self._vert_offsets = np.concatenate([[0], np.cumsum(self._verts_len)]).astype(np.int64)
        self._face_offsets = np.concatenate([[0], np.cumsum(self._faces_len)]).astype(np.int64)
        # faces / topology caches (host numpy + device tensors), filled lazily
        self._faces_packed = None
        self._faces_padded = None
        self._dev: dict = {}
        self._topo = None

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def _from_packed(cls, other: "TriMesh", verts_packed: torch.Tensor) -> "TriMesh":
        """Same topology, new packed verts (the analogue of functor/TriMesh(xs._verts_list, x._faces_list),
        rep/mesh.jl:189-190) — shares every cached topology product."""
        m = cls.__new__(cls)
        m.__dict__.update(other.__dict__)
        m._verts_packed = verts_packed
        m._verts_padded = None
        return m

    def to(self, device):
        return TriMesh(self.get_verts_list(), self._faces_list, offset=self.offset, device=device)

    def __len__(self):
        return self.N

    def __repr__(self):
        return (f"TriMesh{{Float32, Int32, {self.device}}} Structure:\n    Batch size: {self.N}\n"
                f"    Max verts: {self.V}\n    Max faces: {self.F}\n    offset: {self.offset}")

    # ------------------------------------------------------------------ verts getters (rep/mesh.jl:327-380)
    def get_verts_packed(self) -> torch.Tensor:
        return self._verts_packed

    def get_verts_list(self):
        o = self._vert_offsets
        return [self._verts_packed[o[i]:o[i + 1]] for i in range(self.N)]

    def get_verts_padded(self) -> torch.Tensor:
        if self._verts_padded is None:
            if self.equalised:
                self._verts_padded = self._verts_packed.reshape(self.N, self.V, 3)
            elif self._verts_packed.is_cuda:
                # _packed_to_padded (rep/utils.jl:131-152) as one kernel; differentiable (pullback: _padded_to_packed)
                self._verts_padded = _PackedToPadded.apply(self._verts_packed.contiguous(), self.vert_offsets_device(), self.N, self.V)
            else:
                # a host-resident container (only the layout getters work there — every kernel needs CUDA storage):
                # plain indexing, the way the reference fills its padded Array (rep/utils.jl:144-149)
                pad = self._verts_packed.new_zeros((self.N, self.V, 3))
                for i in range(self.N):
                    pad[i, :self._verts_len[i]] = self._verts_packed[self._vert_offsets[i]:self._vert_offsets[i + 1]]
                self._verts_padded = pad
        return self._verts_padded

    def vert_offsets_device(self):
        return self._device_tensor("vert_offsets", lambda: np.asarray(self._vert_offsets, np.int32))

    def face_offsets_device(self):
        return self._device_tensor("face_offsets", lambda: np.asarray(self._face_offsets, np.int32))

    # ------------------------------------------------------------------ faces getters (rep/mesh.jl:382-450, 884-905)
    def get_faces_list(self):
        return self._faces_list

    def get_faces_packed(self) -> np.ndarray:
        """(ΣF, 3) int32, GLOBAL (packed) vertex ids — rep/mesh.jl:884-896."""
        if self._faces_packed is None:
            self._faces_packed = np.ascontiguousarray(np.concatenate(
                [f + np.int32(self._vert_offsets[i]) for i, f in enumerate(self._faces_list)], axis=0), np.int32)
        return self._faces_packed

    def get_faces_padded(self) -> np.ndarray:
        """(N, F, 3) int32, LOCAL ids, padded with -1 — rep/mesh.jl:898-905."""
        if self._faces_padded is None:
            fp = np.full((self.N, self.F, 3), -1, np.int32)
            for i, f in enumerate(self._faces_list):
                fp[i, :f.shape[0]] = f
            self._faces_padded = fp
        return self._faces_padded

    def _device_tensor(self, key, make):
        t = self._dev.get(key)
        if t is None:
            t = torch.from_numpy(np.ascontiguousarray(make())).to(self.device)
            self._dev[key] = t
        return t

    def faces_packed_device(self):
        return self._device_tensor("faces_packed", self.get_faces_packed)

    def faces_padded_device(self):
        """Padded faces with LOCAL vertex ids, pads -1, built ON the device from the packed faces (global ids): the
        inverse of the packed-face offsets of rep/mesh.jl:884-896."""
        t = self._dev.get("faces_padded")
        if t is None:
            if self.equalised and self.N == 1:
                t = self.faces_packed_device().reshape(1, self.F, 3)
            else:
                delta = self._device_tensor("vert_base", lambda: np.asarray(self._vert_offsets[:-1], np.int32))
                t = _packed_to_padded_raw(self.faces_packed_device(), self.face_offsets_device(), self.N, self.F, 0xFFFFFFFF, delta)
            self._dev["faces_padded"] = t
        return t

    def verts_len_device(self):
        return self._device_tensor("verts_len", lambda: np.asarray(self._verts_len, np.int32))

    def faces_len_device(self):
        return self._device_tensor("faces_len", lambda: np.asarray(self._faces_len, np.int32))

    # ------------------------------------------------------------------ topology (built once, cached)
    def _topology(self):
        if self._topo is None:
            L = _lib.lib()
            faces = self.get_faces_packed()
            nV, nF = int(self._vert_offsets[-1]), int(faces.shape[0])
            edges = np.empty((max(3 * nF, 1), 2), np.int32)
            f2e = np.empty((nF, 3), np.int32)
            nE = ctypes.c_int32(0)
            # upper bound 3nF edges → 6nF + nV Laplacian entries
            rowptr = np.empty(nV + 1, np.int32)
            colidx = np.empty(6 * nF + nV, np.int32)
            vals = np.empty(6 * nF + nV, np.float32)
            v2c_rowptr = np.empty(nV + 1, np.int32)
            v2c = np.empty(max(3 * nF, 1), np.int32)
            _lib.check(L.f3d_mesh_topology_build_host(_lib.ptr(faces), nV, nF, _lib.ptr(edges), ctypes.byref(nE),
                                                      _lib.ptr(f2e), _lib.ptr(rowptr), _lib.ptr(colidx),
                                                      _lib.ptr(vals), _lib.ptr(v2c_rowptr), _lib.ptr(v2c)))
            nE = nE.value
            nnz = 2 * nE + nV
            self._topo = dict(edges=edges[:nE].copy(), f2e=f2e, rowptr=rowptr, colidx=colidx[:nnz].copy(),
                              vals=vals[:nnz].copy(), v2c_rowptr=v2c_rowptr, v2c=v2c)
        return self._topo

    def get_edges_packed(self) -> np.ndarray:
        """(E, 2) unique (min, max) edges in lexicographic order — rep/mesh.jl:907-955."""
        return self._topology()["edges"]

    def get_faces_to_edges_packed(self) -> np.ndarray:
        """(ΣF, 3): edge ids of (v2,v3), (v3,v1), (v1,v2) — rep/mesh.jl:943-949."""
        return self._topology()["f2e"]

    def get_laplacian_packed(self):
        """CSR (rowptr, colidx, vals) of the (ΣV, ΣV) Laplacian — rep/mesh.jl:957-1002
        (the reference returns the same matrix as a SparseMatrixCSC)."""
        t = self._topology()
        return t["rowptr"], t["colidx"], t["vals"]

    def _topo_device(self, name):
        return self._device_tensor("topo_" + name, lambda: self._topology()[name])

    # ------------------------------------------------------------------ normals / areas
    def _split(self, packed, offsets):
        return [packed[offsets[i]:offsets[i + 1]] for i in range(self.N)]

    def _pad(self, packed, offsets, width):
        """packed (ΣL, ...) -> zero-padded (N, width, ...): one f3d_packed_to_padded launch instead of a host loop."""
        offs = self.vert_offsets_device() if offsets is self._vert_offsets else self.face_offsets_device()
        return _packed_to_padded_raw(packed, offs, self.N, width, 0)

    def compute_verts_normals_packed(self, mode: int = NORMALS_REFERENCE_CPU) -> torch.Tensor:
        """(ΣV, 3) — rep/mesh.jl:589-618.  mode: NORMALS_REFERENCE_CPU (bit-matches the reference's CPU result)
        or NORMALS_ACCUMULATE (the documented area-weighted sum over all incident faces)."""
        L = _lib.lib()
        verts = self._verts_packed.detach()
        out = torch.empty_like(verts)
        with torch.cuda.device(self.device):
            _lib.check(L.f3d_verts_normals(_lib.ptr(verts), _lib.ptr(self.faces_packed_device()),
                                           _lib.ptr(self._topo_device("v2c_rowptr")), _lib.ptr(self._topo_device("v2c")),
                                           verts.shape[0], self.get_faces_packed().shape[0], mode, _lib.ptr(out),
                                           _lib.stream_ptr(self.device)))
        return out

    def compute_verts_normals_padded(self, mode: int = NORMALS_REFERENCE_CPU):
        return self._pad(self.compute_verts_normals_packed(mode), self._vert_offsets, self.V)  # :640-644

    def compute_verts_normals_list(self, mode: int = NORMALS_REFERENCE_CPU):
        return self._split(self.compute_verts_normals_packed(mode), self._vert_offsets)  # :666-670

    def _faces_areas_normals(self, want_areas, want_normals):
        L = _lib.lib()
        verts = self._verts_packed.detach()
        nF = self.get_faces_packed().shape[0]
        areas = torch.empty(nF, dtype=torch.float32, device=self.device) if want_areas else None
        normals = torch.empty((nF, 3), dtype=torch.float32, device=self.device) if want_normals else None
        with torch.cuda.device(self.device):
            _lib.check(L.f3d_faces_areas_normals(_lib.ptr(verts), _lib.ptr(self.faces_packed_device()), verts.shape[0], nF,
                                                 _lib.ptr(areas), _lib.ptr(normals), _lib.stream_ptr(self.device)))
        return areas, normals

    def compute_faces_normals_packed(self):
        return self._faces_areas_normals(False, True)[1]  # :689-700

    def compute_faces_normals_padded(self):
        return self._pad(self.compute_faces_normals_packed(), self._face_offsets, self.F)

    def compute_faces_normals_list(self):
        return self._split(self.compute_faces_normals_packed(), self._face_offsets)

    def compute_faces_areas_packed(self):
        return self._faces_areas_normals(True, False)[0]  # :765-780

    def compute_faces_areas_padded(self):
        return self._pad(self.compute_faces_areas_packed(), self._face_offsets, self.F)  # :799-808 (zero fill)

    def compute_faces_areas_list(self):
        return self._split(self.compute_faces_areas_packed(), self._face_offsets)


# module-level functions with the reference's names -------------------------------------------------------
def compute_verts_normals_packed(m: TriMesh, mode: int = NORMALS_REFERENCE_CPU):
    return m.compute_verts_normals_packed(mode)


def compute_faces_normals_packed(m: TriMesh):
    return m.compute_faces_normals_packed()


def compute_faces_areas_packed(m: TriMesh):
    return m.compute_faces_areas_packed()


def get_verts_packed(m: TriMesh):
    return m.get_verts_packed()


def get_edges_packed(m: TriMesh):
    return m.get_edges_packed()


def get_laplacian_packed(m: TriMesh):
    return m.get_laplacian_packed()


def offset(m: TriMesh, offset_verts_packed: torch.Tensor) -> TriMesh:
    """Flux3D.offset(m, offset_verts_packed) — src/transforms/mesh_func.jl:435-438: new mesh, verts + offset
    (the reference deep-copies the mesh; here the cached topology is shared, it cannot change)."""
    return TriMesh._from_packed(m, m.get_verts_packed() + offset_verts_packed)


def load_trimesh(path, device="cuda") -> TriMesh:
    """load_trimesh(fn) — src/rep/mesh.jl:297-325, for Wavefront .obj files (v / f lines; polygons are
    fan-triangulated as MeshIO does).  Accepts one path or a list of paths."""
    paths = path if isinstance(path, (list, tuple)) else [path]
    vl, fl = [], []
    for p in paths:
        vs, fs = [], []
        with open(p) as fh:
            for line in fh:
                if line.startswith("v "):
                    vs.append([float(t) for t in line.split()[1:4]])
                elif line.startswith("f "):
                    ids = [int(t.split("/")[0]) for t in line.split()[1:]]
                    ids = [i - 1 if i > 0 else len(vs) + i for i in ids]
                    for k in range(1, len(ids) - 1):
                        fs.append([ids[0], ids[k], ids[k + 1]])
        vl.append(np.asarray(vs, np.float32))
        fl.append(np.asarray(fs, np.int32))
    return TriMesh(vl, fl, device=device)


# ---- losses (src/metrics/mesh.jl) -----------------------------------------------------------------------
class _LaplacianLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, mesh, nV_total):
        L = _lib.lib()
        nV = verts.shape[0]
        dev = verts.device
        rowptr, colidx, vals = (mesh._topo_device(k) for k in ("rowptr", "colidx", "vals"))
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _lib.workspace(("lap", nV), L.f3d_laplacian_workspace_bytes(nV), dev)
            _lib.check(L.f3d_laplacian_loss(_lib.ptr(verts), _lib.ptr(rowptr), _lib.ptr(colidx), _lib.ptr(vals), nV,
                                            nV_total, _lib.ptr(loss), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.save_for_backward(verts)
        ctx.mesh, ctx.nV_total = mesh, nV_total
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        (verts,) = ctx.saved_tensors
        L = _lib.lib()
        mesh, nV = ctx.mesh, verts.shape[0]
        dev = verts.device
        rowptr, colidx, vals = (mesh._topo_device(k) for k in ("rowptr", "colidx", "vals"))
        g = gout.to(torch.float32).reshape(1).contiguous()
        gverts = torch.empty_like(verts)
        with torch.cuda.device(dev):
            ws = _lib.workspace(("lap", nV), L.f3d_laplacian_workspace_bytes(nV), dev)
            _lib.check(L.f3d_laplacian_loss_bwd(_lib.ptr(verts), _lib.ptr(rowptr), _lib.ptr(colidx), _lib.ptr(vals), nV,
                                                ctx.nV_total, _lib.ptr(g), _lib.ptr(gverts), _lib.ptr(ws), ws.numel(),
                                                _lib.stream_ptr(dev)))
        return gverts, None, None


def laplacian_loss(m: TriMesh, *, verts_total: int = 0) -> torch.Tensor:
    """laplacian_loss(m) — src/metrics/mesh.jl:9-15: mean over ALL packed vertices of ‖(L v)_i‖₂.
    verts_total: global vertex count when this process holds one shard of a mesh batch (the shard
    results then sum to the reference value)."""
    return _LaplacianLossFn.apply(m.get_verts_packed().contiguous(), m, int(verts_total))


class _EdgeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, mesh, target, edges_total):
        L = _lib.lib()
        edges = mesh._topo_device("edges")
        nE = int(edges.shape[0])
        dev = verts.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _lib.workspace(("edge", nE), L.f3d_edge_loss_workspace_bytes(nE), dev)
            _lib.check(L.f3d_edge_loss(_lib.ptr(verts), _lib.ptr(edges), nE, int(edges_total), float(target),
                                       _lib.ptr(loss), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        ctx.save_for_backward(verts)
        ctx.mesh, ctx.target, ctx.total = mesh, float(target), int(edges_total) or nE
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        (verts,) = ctx.saved_tensors
        L, mesh = _lib.lib(), ctx.mesh
        g = gout.to(torch.float32).reshape(1).contiguous()
        gverts = torch.empty_like(verts)
        with torch.cuda.device(verts.device):
            _lib.check(L.f3d_edge_loss_bwd(_lib.ptr(verts), _lib.ptr(mesh._topo_device("rowptr")),
                                           _lib.ptr(mesh._topo_device("colidx")), verts.shape[0], ctx.total, ctx.target,
                                           _lib.ptr(g), _lib.ptr(gverts), _lib.stream_ptr(verts.device)))
        return gverts, None, None, None


def edge_loss(m: TriMesh, target_length: float = 0.0, *, edges_total: int = 0) -> torch.Tensor:
    """edge_loss(m, target_length=0.0) — src/metrics/mesh.jl:24-32: mean over the unique edges of
    (‖v1 - v2‖ - target)², differentiable w.r.t. the vertices."""
    return _EdgeLossFn.apply(m.get_verts_packed().contiguous(), m, float(target_length), int(edges_total))
