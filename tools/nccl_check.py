"""Multi-GPU check (run under torchrun, one rank per GPU): the batch-sharded chamfer path with (a) the library's own
NCCL binding (f3d_comm_* / f3d_allreduce_sum_f32) and (b) torch.distributed's group gives, on every rank, the same
loss as the un-sharded call on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import flux3d_b200 as f3d

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, N, M = 13, 1500, 1100   # B not divisible by the world size: remainder goes to the low ranks
A = torch.from_numpy(np.random.default_rng(1).random((B, N, 3), dtype=np.float32)).to(dev)
Bc = torch.from_numpy(np.random.default_rng(2).random((B, M, 3), dtype=np.float32)).to(dev)
full = f3d.chamfer_distance(A, Bc, w1=0.7, w2=1.3)
lo, hi = f3d.shard_range(B, rank, world)
via_torch = f3d.chamfer_distance_sharded(A[lo:hi].contiguous(), Bc[lo:hi].contiguous(), B, w1=0.7, w2=1.3)
comm = f3d.Communicator(rank, world, dev)
via_lib = f3d.chamfer_distance_sharded(A[lo:hi].contiguous(), Bc[lo:hi].contiguous(), B, w1=0.7, w2=1.3, comm=comm)
torch.cuda.synchronize()
ok = abs(via_torch.item() - full.item()) <= 1e-6 * full.item() and abs(via_lib.item() - full.item()) <= 1e-6 * full.item()
print(f"rank {rank}/{world}: shard [{lo},{hi}) full {full.item():.9f} torch-nccl {via_torch.item():.9f} lib-nccl {via_lib.item():.9f} {'OK' if ok else 'MISMATCH'}", flush=True)
comm.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
