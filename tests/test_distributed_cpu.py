"""N>1 host logic on CPU: two gloo ranks shard a batch with shard_range, compute their partial Chamfer losses
with the GLOBAL denominators (here via the oracle — the CUDA path needs a GPU) and sum them with the same
all-reduce wrapper the GPU path uses.  The result must equal the un-sharded loss."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import flux3d_b200 as f3d
    from oracle import oracle as O
    A = np.random.default_rng(501).random((B, 200, 3), dtype=np.float32)
    Bc = np.random.default_rng(502).random((B, 150, 3), dtype=np.float32)
    lo, hi = f3d.shard_range(B, rank, world)
    if hi > lo:
        _, _, _, terms = O.chamfer_distance(A[lo:hi], Bc[lo:hi], return_all=True)
        # shard means → partial sums over the global denominators (what f3d_chamfer_fwd does with B_total)
        part = (0.5 * terms[0] + 2.0 * terms[1]) * (hi - lo) / B
    else:
        part = 0.0
    t = torch.tensor([part], dtype=torch.float32)
    f3d.allreduce_loss_(t)
    q.put((rank, float(t.item()), (lo, hi)))
    dist.destroy_process_group()


def _run(world, B, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_two_rank_sharded_loss_matches_unsharded(oracle):
    for B, port in ((5, 29611), (1, 29612)):  # B=1 < world: rank 1 holds an empty shard
        res = _run(2, B, port)
        A = np.random.default_rng(501).random((B, 200, 3), dtype=np.float32)
        Bc = np.random.default_rng(502).random((B, 150, 3), dtype=np.float32)
        full = float(oracle.chamfer_distance(A, Bc, 0.5, 2.0))
        assert res[0][1] == res[1][1]
        assert abs(res[0][1] - full) <= 1e-5 * full
        assert res[0][2][1] == res[1][2][0] and res[1][2][1] == B
