"""Multi-GPU check of the fused cross-rank loss sum (run under torchrun, one rank per GPU):
fused (peer mailboxes inside the finalize kernel) == NCCL all-reduce of the shard losses == full-batch loss, and what
each costs per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import flux3d_b200 as f3d

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Bt, N, M = 8 * world, 2048, 1536
A = np.random.default_rng(11).random((Bt, N, 3), dtype=np.float32)
B = np.random.default_rng(12).random((Bt, M, 3), dtype=np.float32)
lo, hi = f3d.shard_range(Bt, rank, world)
dA, dB = torch.from_numpy(A[lo:hi]).to(dev), torch.from_numpy(B[lo:hi]).to(dev)
comm = f3d.Communicator(rank, world, dev).enable_p2p()
ref = f3d.chamfer_distance_sharded(dA, dB, Bt)                      # torch.distributed NCCL all-reduce
for it in range(5):                                                  # several steps: both mailbox parities, reuse
    fused = f3d.chamfer_distance_sharded(dA, dB, Bt, comm=comm)
    host = f3d.chamfer_distance_sharded(torch.from_numpy(A[lo:hi]).pin_memory(), torch.from_numpy(B[lo:hi]).pin_memory(), Bt, comm=comm, to_host=True)
    torch.cuda.synchronize()
    assert abs(fused.item() - ref.item()) <= 1e-6 * ref.item(), (rank, it, fused.item(), ref.item())
    assert host.item() == fused.item(), (rank, it, host.item(), fused.item())
# every rank holds the same bits
t = torch.tensor([fused.item()], dtype=torch.float64, device=dev)
g = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(g, t)
assert all(x.item() == g[0].item() for x in g)
if rank == 0:
    full = float(f3d.chamfer_distance(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)).item())
    assert abs(full - fused.item()) <= 1e-6 * full
    print(f"fused sum ok on {world} ranks: {fused.item():.9f} (single-GPU full batch {full:.9f}, NCCL path {ref.item():.9f})", flush=True)

# the tensor-core sweep (its cleanup kernel carries the exchange) on a shard large enough to take it by default, and the
# differentiable sharded loss: gradient == that of the whole batch on one GPU
Bt2 = 40 * world
A2 = np.random.default_rng(21).random((Bt2, 2048, 3), dtype=np.float32)
B2 = np.random.default_rng(22).random((Bt2, 2048, 3), dtype=np.float32)
lo2, hi2 = f3d.shard_range(Bt2, rank, world)
sA = torch.from_numpy(A2[lo2:hi2]).to(dev).requires_grad_(True)
sB = torch.from_numpy(B2[lo2:hi2]).to(dev).requires_grad_(True)
for it in range(3):
    l_tc = f3d.chamfer_distance_sharded(sA.detach(), sB.detach(), Bt2, comm=comm)
    l_nc = f3d.chamfer_distance_sharded(sA.detach(), sB.detach(), Bt2)
    torch.cuda.synchronize()
    assert abs(l_tc.item() - l_nc.item()) <= 1e-6 * l_nc.item(), (rank, it, l_tc.item(), l_nc.item())
lg = f3d.chamfer_distance_sharded(sA, sB, Bt2, comm=comm)
lg.backward()
assert abs(lg.item() - l_nc.item()) <= 1e-6 * l_nc.item()
if rank == 0:
    fA = torch.from_numpy(A2).to(dev).requires_grad_(True)
    fB = torch.from_numpy(B2).to(dev).requires_grad_(True)
    lf = f3d.chamfer_distance(fA, fB)
    lf.backward()
    assert abs(lf.item() - lg.item()) <= 1e-6 * lf.item()
    assert torch.allclose(fA.grad[lo2:hi2], sA.grad, rtol=1e-5, atol=1e-12) and torch.allclose(fB.grad[lo2:hi2], sB.grad, rtol=1e-5, atol=1e-12)
    print(f"tensor-core sweep + fused sum + sharded gradient ok on {world} ranks: {lg.item():.9f}", flush=True)


def timed(fn, steps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

wl = dict(B=32, N=4096, M=4096)
a = torch.rand((32, 4096, 3), device=dev); b = torch.rand((32, 4096, 3), device=dev)
t_local = timed(lambda: f3d.chamfer_forward_raw(a, b, 1.0, 1.0, batch_total=32 * world, want_indices=False))
t_nccl = timed(lambda: f3d.chamfer_distance_sharded(a, b, 32 * world))
t_lib = timed(lambda: f3d.chamfer_distance_sharded(a, b, 32 * world, comm=f3d.distributed._NcclOnly(comm)))
t_fused = timed(lambda: f3d.chamfer_distance_sharded(a, b, 32 * world, comm=comm))
if rank == 0:
    print(f"cfg2 shard per rank, us/step (max over ranks, back to back): no exchange {t_local:.1f} | torch NCCL all-reduce {t_nccl:.1f} | "
          f"library NCCL all-reduce {t_lib:.1f} | fused peer-mailbox sum {t_fused:.1f}", flush=True)
comm.close()
dist.destroy_process_group()
