"""Batch-sharded multi-GPU path: one process per GPU, the batch axis split contiguously across ranks, and
ONE all-reduce of the per-shard scalar loss.  The reference has no distributed code.

``shard_range`` is pure host logic (tested on CPU).  The all-reduce runs either through the library's own
NCCL binding (f3d_comm_init / f3d_allreduce_sum_f32 — what a Julia caller would use) or through an
existing torch.distributed process group (backend nccl on GPUs; gloo in the CPU tests of the sharding
logic)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def shard_range(total: int, rank: int, world: int):
    """Contiguous split of ``total`` batch elements over ``world`` ranks, remainder to the low ranks.
    Returns (start, stop); empty when total < world for the high ranks (those ranks then contribute 0)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class Communicator:
    """Thin owner of the library's NCCL handle.  The 128-byte unique id is created on rank 0 and broadcast
    out of band — here through torch.distributed's store/broadcast_object_list, in Julia through MPI or a
    shared file."""

    def __init__(self, rank: int, world: int, device):
        import torch.distributed as dist
        self.rank, self.world, self.device = rank, world, torch.device(device)
        L = _lib.lib()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(L.f3d_comm_unique_id_host(buf))
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        idbuf = ctypes.create_string_buffer(obj[0], 128)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.f3d_comm_init(world, rank, idbuf, ctypes.byref(h)))
        self._h = h
        self.p2p = False

    def enable_p2p(self):
        """Collective: map every rank's mailbox over NVLink (CUDA IPC) so that chamfer_distance_sharded can sum the
        shard losses INSIDE the finalize kernel instead of calling NCCL after it (f3d_comm_enable_p2p)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().f3d_comm_enable_p2p(self._h, _lib.stream_ptr(self.device)))
        self.p2p = True
        return self

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("allreduce_sum_ needs a contiguous float32 CUDA tensor")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().f3d_allreduce_sum_f32(self._h, _lib.ptr(t), t.numel(), _lib.stream_ptr(self.device)))
        return t

    def close(self):
        if self._h:
            _lib.check(_lib.lib().f3d_comm_destroy(self._h))
            self._h = None


class _NcclOnly:
    """View of a Communicator that hides its peer mailboxes (measurement aid: forces the NCCL all-reduce path)."""

    def __init__(self, comm):
        self._c, self.p2p = comm, False

    def allreduce_sum_(self, t):
        return self._c.allreduce_sum_(t)


def allreduce_loss_(loss: torch.Tensor, comm=None) -> torch.Tensor:
    """In-place sum of the per-shard loss over all ranks: ``comm`` a Communicator, or None to use the default
    torch.distributed group."""
    if comm is not None:
        return comm.allreduce_sum_(loss)
    import torch.distributed as dist
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    return loss


class _ShardedChamferFn(torch.autograd.Function):
    """chamfer_distance of a batch sharded over ranks, differentiable with respect to this rank's shard: the loss every rank
    holds is the whole batch's, so d loss / d shard is the single-GPU pullback (src/metrics/pcloud.jl:47-50) evaluated with
    the global N*B_total / M*B_total — f3d_chamfer_bwd with batch_total; no further exchange is needed."""

    @staticmethod
    def forward(ctx, A, B, w1, w2, batch_total, comm, flags):
        L = _lib.lib()
        Bn, N, M = A.shape[0], A.shape[1], B.shape[1]
        dev = A.device
        nnA = torch.empty((Bn, N), dtype=torch.int32, device=dev)
        nnB = torch.empty((Bn, M), dtype=torch.int32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        Ad, Bd = A.detach().contiguous(), B.detach().contiguous()
        with torch.cuda.device(dev):
            ws = _lib.workspace(("chamfer", Bn, N, M), L.f3d_chamfer_workspace_bytes(Bn, N, M), dev)
            if comm is not None and getattr(comm, "p2p", False) and not flags:
                _lib.check(L.f3d_chamfer_fwd_allreduce(comm._h, _lib.ptr(Ad), _lib.ptr(Bd), Bn, N, M, w1, w2, batch_total, _lib.ptr(loss),
                                                       _lib.ptr(nnA), _lib.ptr(nnB), _lib.ptr(ws), ws.numel(), 0, _lib.stream_ptr(dev)))
            else:
                _lib.check(L.f3d_chamfer_fwd(_lib.ptr(Ad), _lib.ptr(Bd), Bn, N, M, w1, w2, batch_total, _lib.ptr(loss), None,
                                             _lib.ptr(nnA), _lib.ptr(nnB), _lib.ptr(ws), ws.numel(), flags, _lib.stream_ptr(dev)))
                allreduce_loss_(loss, comm)
        ctx.save_for_backward(Ad, Bd, nnA, nnB)
        ctx.cfg = (w1, w2, batch_total)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        A, B, nnA, nnB = ctx.saved_tensors
        w1, w2, batch_total = ctx.cfg
        L = _lib.lib()
        gA, gB = torch.empty_like(A), torch.empty_like(B)
        g = gout.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(A.device):
            _lib.check(L.f3d_chamfer_bwd(_lib.ptr(A), _lib.ptr(B), A.shape[0], A.shape[1], B.shape[1], w1, w2, batch_total,
                                         _lib.ptr(nnA), _lib.ptr(nnB), _lib.ptr(g), _lib.ptr(gA), _lib.ptr(gB), _lib.stream_ptr(A.device)))
        return gA, gB, None, None, None, None, None


def chamfer_distance_sharded(A_shard, B_shard, batch_total: int, *, w1: float = 1.0, w2: float = 1.0, comm=None,
                             flags: int = 0, to_host: bool = False) -> torch.Tensor:
    """chamfer_distance over a batch split across ranks: every rank passes its shard (b_local, N, 3) /
    (b_local, M, 3) and the GLOBAL batch size; the per-shard partial losses (already divided by the global
    N*B_total / M*B_total) are summed with one all-reduce, so every rank returns the reference's value for
    the whole batch.  A rank with an empty shard contributes 0."""
    from .metrics import chamfer_forward_host, chamfer_forward_raw
    if isinstance(A_shard, torch.Tensor) and isinstance(B_shard, torch.Tensor) and A_shard.is_cuda and (A_shard.requires_grad or B_shard.requires_grad) \
            and torch.is_grad_enabled():
        # differentiable: the forward keeps the argmin indices, the pullback is f3d_chamfer_bwd with the GLOBAL denominators
        return _ShardedChamferFn.apply(A_shard, B_shard, float(w1), float(w2), int(batch_total), comm, int(flags))
    if comm is not None and getattr(comm, "p2p", False) and not flags:
        # fused exchange: the finalize kernel's last block trades the shard losses with its peers through mailboxes
        # mapped over NVLink — no NCCL call, no extra launch.  Every rank must take part, so shards may not be empty.
        if len(A_shard) == 0:
            raise ValueError("the fused cross-rank sum needs a non-empty shard on every rank (batch_total >= world size)")
        if not (isinstance(A_shard, torch.Tensor) and A_shard.is_cuda):
            return chamfer_forward_host(A_shard, B_shard, w1, w2, batch_total=batch_total, comm=comm._h,
                                        to_host=to_host, device=comm.device).reshape(())
        L = _lib.lib()
        Bn, N, M = A_shard.shape[0], A_shard.shape[1], B_shard.shape[1]
        dev = A_shard.device
        with torch.cuda.device(dev):
            ws = _lib.workspace(("chamfer", Bn, N, M), L.f3d_chamfer_workspace_bytes(Bn, N, M), dev)
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.check(L.f3d_chamfer_fwd_allreduce(comm._h, _lib.ptr(A_shard), _lib.ptr(B_shard), Bn, N, M, w1, w2, batch_total,
                                                   _lib.ptr(loss), None, None, _lib.ptr(ws), ws.numel(), 0, _lib.stream_ptr(dev)))
        return loss.reshape(())
    if not (isinstance(A_shard, torch.Tensor) and A_shard.is_cuda) and len(A_shard) > 0:
        # host shards: upload pipelined against the sweep (f3d_chamfer_pipe_run), then the same one all-reduce
        loss = chamfer_forward_host(A_shard, B_shard, w1, w2, batch_total=batch_total, flags=flags)
        return allreduce_loss_(loss, comm).reshape(())
    if A_shard.shape[0] == 0:
        loss = torch.zeros(1, dtype=torch.float32, device=A_shard.device)
    else:
        loss, _, _, _ = chamfer_forward_raw(A_shard, B_shard, w1, w2, batch_total=batch_total, want_indices=False,
                                            flags=flags)
        loss = loss.clone()
    return allreduce_loss_(loss, comm).reshape(())
