"""chamfer_distance / laplacian_loss / edge_loss — host-side mirror of src/metrics/{pcloud,mesh}.jl.

Same names, argument meaning and error behaviour as the reference; all arithmetic happens in
libflux3d_b200.so (hand-written sm_100a kernels) through the C ABI.  No CPU / PyTorch fallback."""
from __future__ import annotations

import torch

from . import _lib
from .pcloud import PointCloud, as_f32_tensor

FLAG_NONE = 0
FLAG_FMA = 1  # non-reference arithmetic (fused multiply-add distances); see include/flux3d_b200.h
FLAG_SWEEP_ONLY = 2  # measurement aid: launch only the sweep kernel
FLAG_EXACT_SWEEP = 4  # every pair in the reference arithmetic (cross-check of the default filtered sweep)
FLAG_TENSOR = 8  # force the tensor-core (tcgen05) filter path for any shape (kNN graph, chamfer sweep)
FLAG_CUDA_CORES = 16  # chamfer: keep the filter sweep on the CUDA cores also for large problems (A/B aid, cross-check)

_workspace = _lib.workspace
_stream_ptr = _lib.stream_ptr


def chamfer_forward_raw(A: torch.Tensor, B: torch.Tensor, w1: float, w2: float, *, batch_total: int = 0,
                        want_indices: bool = True, flags: int = FLAG_NONE, out=None):
    """One call of f3d_chamfer_fwd on device tensors A (B,N,3), B (B,M,3).
    Returns (loss[1], terms[2], nnA (B,N) | None, nnB (B,M) | None) — all device tensors, no sync."""
    L = _lib.lib()
    if not (A.is_cuda and B.is_cuda):
        raise _lib.Flux3DB200Error("chamfer_forward_raw needs CUDA tensors (there is no CPU path)")
    if A.dim() != 3 or B.dim() != 3 or A.shape[2] != 3 or B.shape[2] != 3:
        raise ValueError("expected (B, N, 3) and (B, M, 3) point arrays")
    if A.shape[0] != B.shape[0]:
        raise ValueError(f"batch sizes differ: {A.shape[0]} vs {B.shape[0]}")
    Bn, N, M = A.shape[0], A.shape[1], B.shape[1]
    dev = A.device
    with torch.cuda.device(dev):
        ws = _workspace(("chamfer", Bn, N, M), L.f3d_chamfer_workspace_bytes(Bn, N, M), dev)
        if out is None:
            res = torch.empty(3, dtype=torch.float32, device=dev)
            nnA = torch.empty((Bn, N), dtype=torch.int32, device=dev) if want_indices else None
            nnB = torch.empty((Bn, M), dtype=torch.int32, device=dev) if want_indices else None
        else:
            res, nnA, nnB = out
        loss, terms = res[0:1], res[1:3]
        _lib.check(L.f3d_chamfer_fwd(_lib.ptr(A), _lib.ptr(B), Bn, N, M, w1, w2, batch_total,
                                     _lib.ptr(loss), _lib.ptr(terms), _lib.ptr(nnA), _lib.ptr(nnB),
                                     _lib.ptr(ws), ws.numel(), flags, _stream_ptr(dev)))
    return loss, terms, nnA, nnB


_pipes: dict = {}
_inflight: list = []  # (event, A, B) of host-array calls whose uploads may still be running


def _is_host(x) -> bool:
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def _host_f32(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor) and x.dtype == torch.float32 and not x.is_cuda and x.is_contiguous():
        return x  # the common case: no conversion, no copy
    return as_f32_tensor(x)


_host_plans: dict = {}  # (device index, stream, uploaders, B, N, M) -> (pipe handle, workspace, its address, its size)


def chamfer_forward_host(A, B, w1: float = 1.0, w2: float = 1.0, *, batch_total: int = 0, uploaders: int = 0,
                         flags: int = FLAG_NONE, device="cuda", to_host: bool = False, comm=None) -> torch.Tensor:
    """One call of f3d_chamfer_pipe_run: HOST arrays A (B,N,3), B (B,M,3) → the loss.  With page-locked inputs the grid
    pulls the batch over PCIe itself (its first ``uploaders`` CTAs; 0 = default) while the other CTAs sweep what has
    landed; pageable inputs are copied first.  ``to_host=False``: loss[1] on ``device``, no host synchronisation.
    ``to_host=True``: a 0-dim CPU tensor — the grid stores the loss into mapped host memory and the call returns when
    it has landed (no D2H copy).  ``comm``: the handle of a Communicator with peer mailboxes — A and B are then this
    rank's shard of a ``batch_total`` batch and the loss is the whole batch's, summed inside the finalize kernel."""
    L = _lib.lib()
    A = _host_f32(A)
    B = _host_f32(B)
    if A.is_cuda or B.is_cuda:
        raise ValueError("chamfer_forward_host takes host arrays; use chamfer_forward_raw for device tensors")
    if A.dim() != 3 or B.dim() != 3 or A.shape[2] != 3 or B.shape[2] != 3:
        raise ValueError("expected (B, N, 3) and (B, M, 3) point arrays")
    if A.shape[0] != B.shape[0]:
        raise ValueError(f"batch sizes differ: {A.shape[0]} vs {B.shape[0]}")
    Bn, N, M = A.shape[0], A.shape[1], B.shape[1]
    dev = device if isinstance(device, torch.device) else torch.device(device)
    cur = torch.cuda.current_device()
    idx = cur if dev.index is None else dev.index
    uploaders = max(0, min(int(uploaders), 1024))
    guard = torch.cuda.device(idx) if idx != cur else None  # the library works on the current device
    if guard is not None:
        guard.__enter__()
    try:
        stream = torch.cuda.current_stream(idx)
        sptr = stream.cuda_stream
        key = (idx, sptr, uploaders, Bn, N, M)
        plan = _host_plans.get(key)
        if plan is None:
            h = _pipes.get((idx, uploaders))
            if h is None:
                h = _lib.C.c_void_p()
                _lib.check(L.f3d_chamfer_pipe_create(uploaders, _lib.C.byref(h)))
                _pipes[(idx, uploaders)] = h
            ws = _workspace(("chamfer_pipe", Bn, N, M), L.f3d_chamfer_pipe_workspace_bytes(Bn, N, M), torch.device("cuda", idx))
            plan = _host_plans[key] = (h, ws, ws.data_ptr(), ws.numel())
        h, ws, wptr, wsize = plan
        if to_host:
            out = _lib.C.c_float()
            _lib.check(L.f3d_chamfer_pipe_run(h, A.data_ptr(), B.data_ptr(), Bn, N, M, w1, w2, batch_total, None,
                                              _lib.C.byref(out), wptr, wsize, flags, comm, sptr))
            return torch.tensor(out.value, dtype=torch.float32)  # the call returned after the loss landed
        loss = torch.empty(1, dtype=torch.float32, device=torch.device("cuda", idx))
        _lib.check(L.f3d_chamfer_pipe_run(h, A.data_ptr(), B.data_ptr(), Bn, N, M, w1, w2, batch_total, loss.data_ptr(),
                                          None, wptr, wsize, flags, comm, sptr))
        # the grid reads A and B asynchronously: keep them alive until the stream has consumed them
        ev = torch.cuda.Event()
        ev.record(stream)
        _inflight[:] = [e for e in _inflight if not e[0].query()]
        _inflight.append((ev, A, B))
    finally:
        if guard is not None:
            guard.__exit__(None, None, None)
    return loss


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, w1, w2, batch_total, flags):
        loss, _, nnA, nnB = chamfer_forward_raw(A, B, w1, w2, batch_total=batch_total, flags=flags)
        ctx.save_for_backward(A, B, nnA, nnB)
        ctx.cfg = (w1, w2, batch_total)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        A, B, nnA, nnB = ctx.saved_tensors
        w1, w2, batch_total = ctx.cfg
        L = _lib.lib()
        gA = torch.empty_like(A)
        gB = torch.empty_like(B)
        g = gout.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(A.device):
            _lib.check(L.f3d_chamfer_bwd(_lib.ptr(A), _lib.ptr(B), A.shape[0], A.shape[1], B.shape[1], w1, w2,
                                         batch_total, _lib.ptr(nnA), _lib.ptr(nnB), _lib.ptr(g), _lib.ptr(gA),
                                         _lib.ptr(gB), _stream_ptr(A.device)))
        return gA, gB, None, None, None, None


def chamfer_distance(A, B, num_samples: int = 5000, *, w1: float = 1.0, w2: float = 1.0,
                     flags: int = FLAG_NONE, batch_total: int = 0, device="cuda"):
    """chamfer_distance(A, B; w1=1.0, w2=1.0) — src/metrics/pcloud.jl:11-37; for two TriMesh arguments,
    chamfer_distance(m1, m2, num_samples=5000; w1, w2) — src/metrics/mesh.jl:34-44.

    A, B: PointCloud, or arrays/tensors of shape (N,3) / (B,N,3) (== Julia (3,N) / (3,N,B)).  Host inputs
    are copied to ``device``.  Returns a 0-dim float32 CUDA tensor (differentiable w.r.t. A and B) — except when both
    clouds are host arrays and nothing needs a gradient: then, like the reference on ``Array``s, the scalar comes back
    on the host (0-dim CPU tensor; f3d_chamfer_pipe_run: uploads overlapped with the sweep, no D2H copy)."""
    from .mesh import TriMesh  # local import: mesh.py imports this module's helpers
    if isinstance(A, TriMesh) and isinstance(B, TriMesh):
        from .sampling import sample_points
        A = sample_points(A, num_samples)
        B = sample_points(B, num_samples)
    if isinstance(A, PointCloud):
        A = A.points
    if isinstance(B, PointCloud):
        B = B.points
    needs_grad = torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in (A, B))
    if _is_host(A) and _is_host(B) and not needs_grad:
        # both clouds on the host, nothing to differentiate: upload pipelined against the sweep, one C call
        A, B = as_f32_tensor(A), as_f32_tensor(B)
        if A.dim() == 2:
            A = A.unsqueeze(0)
        if B.dim() == 2:
            B = B.unsqueeze(0)
        return chamfer_forward_host(A, B, float(w1), float(w2), batch_total=int(batch_total), flags=int(flags),
                                    device=device, to_host=True)
    A = as_f32_tensor(A, None if (isinstance(A, torch.Tensor) and A.is_cuda) else device)
    B = as_f32_tensor(B, A.device)
    if A.dim() == 2:
        A = A.unsqueeze(0)
    if B.dim() == 2:
        B = B.unsqueeze(0)
    if not (torch.is_grad_enabled() and (A.requires_grad or B.requires_grad)):
        # nothing to differentiate: the argmin indices (only the pullback needs them) are not materialised
        loss, _, _, _ = chamfer_forward_raw(A, B, float(w1), float(w2), batch_total=int(batch_total),
                                            want_indices=False, flags=int(flags))
        return loss.reshape(())
    return _ChamferFn.apply(A, B, float(w1), float(w2), int(batch_total), int(flags))


def nearest_neighbors(A, B, flags: int = FLAG_NONE):
    """_nearest_neighbors(x, y) — src/metrics/pcloud.jl:54-86: (nn_for_x (B,N), nn_for_y (B,M)), 0-based int32."""
    A = as_f32_tensor(A.points if isinstance(A, PointCloud) else A, "cuda")
    B = as_f32_tensor(B.points if isinstance(B, PointCloud) else B, "cuda")
    if A.dim() == 2:
        A, B = A.unsqueeze(0), B.unsqueeze(0)
    _, _, nnA, nnB = chamfer_forward_raw(A, B, 1.0, 1.0, flags=flags)
    return nnA, nnB
