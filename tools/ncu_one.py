import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d
import numpy as np
B, N, M = (int(v) for v in sys.argv[1].split("x"))
seeds = (201, 202) if N == 4096 else (501, 502)
A = torch.from_numpy(np.random.default_rng(seeds[0]).random((B, N, 3), dtype=np.float32)).cuda()
Bc = torch.from_numpy(np.random.default_rng(seeds[1]).random((B, M, 3), dtype=np.float32)).cuda()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, want_indices=False)
torch.cuda.synchronize()
