"""Two ranks on two GPUs of one box (skipped where fewer are visible — the driver's single-GPU test box; bench.py --gpus N checks
the loss of every multi-rank run against the CPU port itself): tools/p2p_check.py under torchrun verifies that the fused
cross-rank sum inside the chamfer kernels == the NCCL all-reduce of the shard losses == the single-GPU loss of the whole batch,
for device and host shards, on both sweeps, several steps in a row, and that the sharded loss is differentiable."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_fused_sum():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", os.path.join(ROOT, "tools", "p2p_check.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "fused sum ok on 2 ranks" in r.stdout
