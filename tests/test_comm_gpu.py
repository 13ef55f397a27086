"""The library's own NCCL binding on one GPU (world size 1) — the N>1 logic is covered on CPU with gloo
(tests/test_distributed_cpu.py) and measured by bench.py --gpus N."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_comm_world1(f3d):
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29621")
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        comm = f3d.Communicator(0, 1, "cuda:0")
        t = torch.tensor([1.5, -2.0], device="cuda")
        comm.allreduce_sum_(t)
        torch.cuda.synchronize()
        assert t.tolist() == [1.5, -2.0]
        A = torch.rand((3, 500, 3), device="cuda")
        B = torch.rand((3, 400, 3), device="cuda")
        full = f3d.chamfer_distance(A, B)
        lo, hi = f3d.shard_range(3, 0, 1)
        sharded = f3d.chamfer_distance_sharded(A[lo:hi], B[lo:hi], 3, comm=comm)
        assert torch.equal(full, sharded)
        # two "virtual shards" on one GPU add up to the full-batch loss (global denominators)
        parts = [f3d.chamfer_forward_raw(A[a:b].contiguous(), B[a:b].contiguous(), 1.0, 1.0, batch_total=3)[0].clone()
                 for a, b in ((0, 2), (2, 3))]
        assert abs((parts[0] + parts[1]).item() - full.item()) <= 1e-6 * full.item()
        # the fused cross-rank sum (peer mailboxes): with one rank it must set everything up (CUDA IPC export, handle
        # gather over NCCL) and reproduce the plain call bit for bit, for device and for host shards
        comm.enable_p2p()
        fused = f3d.chamfer_distance_sharded(A, B, 3, comm=comm)
        fused_host = f3d.chamfer_distance_sharded(A.cpu().pin_memory(), B.cpu().pin_memory(), 3, comm=comm, to_host=True)
        assert torch.equal(full, fused) and fused_host.item() == full.item()
        comm.close()
    finally:
        if own:
            dist.destroy_process_group()
