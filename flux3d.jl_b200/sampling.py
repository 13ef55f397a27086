"""sample_points — host-side mirror of src/transforms/mesh_func.jl:21-82 (TriMesh → PointCloud points).

One launch of f3d_sample_points for the whole mesh batch (the reference loops over meshes on the host
with CPU sampling and ~17 copies/launches per mesh).  Draws are Philox4x32-10 keyed by (seed, offset);
``seed=None`` takes a fresh seed from torch's CPU generator, so ``torch.manual_seed`` makes runs
reproducible the way ``Random.seed!`` does for the reference."""
from __future__ import annotations

import torch

from . import _lib
from .mesh import TriMesh

EPS = 1e-6  # src/transforms/utils.jl:4


def _launch(m, verts, num_samples, eps, seed, offset, inj, want_faces, want_bary, counter=None):
    L = _lib.lib()
    dev = m.device
    faces = m.faces_padded_device()
    out = torch.empty((m.N, num_samples, 3), dtype=torch.float32, device=dev)
    fidx = torch.empty((m.N, num_samples), dtype=torch.int32, device=dev) if want_faces else None
    bary = torch.empty((m.N, num_samples, 3), dtype=torch.float32, device=dev) if want_bary else None
    inj_face, inj_r1, inj_r2 = inj
    with torch.cuda.device(dev):
        nws = L.f3d_sample_points_workspace_bytes(m.N, m.F)
        ws = _lib.workspace(("sample", m.N, m.F), nws, dev) if nws else None
        _lib.check(L.f3d_sample_points_replayable(_lib.ptr(verts), _lib.ptr(faces), _lib.ptr(m.verts_len_device()),
                                                  _lib.ptr(m.faces_len_device()), m.N, m.V, m.F, num_samples, float(eps),
                                                  int(seed), int(offset), _lib.ptr(counter), _lib.ptr(inj_face), _lib.ptr(inj_r1),
                                                  _lib.ptr(inj_r2), _lib.ptr(out), _lib.ptr(fidx), _lib.ptr(bary), _lib.ptr(ws),
                                                  ws.numel() if ws is not None else 0, _lib.stream_ptr(dev)))
    return out, fidx, bary


class _SampleFn(torch.autograd.Function):
    """Differentiable w.r.t. the (padded) vertices; the face draws are constants (mesh_func.jl:47 is @ignore)."""

    @staticmethod
    def forward(ctx, verts_padded, m, num_samples, eps, seed, offset, inj, counter=None):
        out, fidx, bary = _launch(m, verts_padded, num_samples, eps, seed, offset, inj, True, True, counter)
        ctx.save_for_backward(fidx, bary)
        ctx.m, ctx.S = m, num_samples
        ctx.mark_non_differentiable(fidx)
        return out, fidx

    @staticmethod
    def backward(ctx, gout, _gf):
        fidx, bary = ctx.saved_tensors
        m, L = ctx.m, _lib.lib()
        g = gout.to(torch.float32).contiguous()
        gv = torch.zeros((m.N, m.V, 3), dtype=torch.float32, device=m.device)
        with torch.cuda.device(m.device):
            _lib.check(L.f3d_sample_points_bwd(_lib.ptr(g), _lib.ptr(fidx), _lib.ptr(bary), _lib.ptr(m.faces_padded_device()),
                                               m.N, m.V, m.F, ctx.S, _lib.ptr(gv), _lib.stream_ptr(m.device)))
        return gv, None, None, None, None, None, None, None


def sample_points(m: TriMesh, num_samples: int = 5000, *, eps: float = EPS, seed=None, offset: int = 0,
                  inj_face=None, inj_r1=None, inj_r2=None, return_faces: bool = False, counter=None):
    """sample_points(m, num_samples=5000; eps=1e-6) → (N, num_samples, 3) float32 device tensor
    (== Julia (3, num_samples, N)), differentiable w.r.t. the mesh vertices.  inj_face/inj_r1/inj_r2 ((N, S)
    int32 / float32 / float32 device tensors) inject the draws (bit-parity mode); return_faces also returns the
    sampled face ids (N, S).  counter: a 1-element int64 device tensor added to ``offset`` on the device and incremented after the
    draws — inside a captured CUDA graph (graph.capture_step) every replay then draws fresh samples."""
    if num_samples <= 0:
        raise ValueError("num_samples must be positive")
    dev = m.device
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    if inj_face is not None:
        inj_face = inj_face.to(dev, torch.int32).contiguous()
        inj_r1 = inj_r1.to(dev, torch.float32).contiguous()
        inj_r2 = inj_r2.to(dev, torch.float32).contiguous()
        if inj_face.shape != (m.N, num_samples) or inj_r1.shape != inj_face.shape or inj_r2.shape != inj_face.shape:
            raise ValueError("injected draws must have shape (N, num_samples)")
    inj = (inj_face, inj_r1, inj_r2)
    verts = m.get_verts_padded()
    if torch.is_grad_enabled() and verts.requires_grad:
        out, fidx = _SampleFn.apply(verts.contiguous(), m, num_samples, eps, seed, offset, inj, counter)
    else:
        out, fidx, _ = _launch(m, verts.detach().contiguous(), num_samples, eps, seed, offset, inj, return_faces, False, counter)
    return (out, fidx) if return_faces else out
