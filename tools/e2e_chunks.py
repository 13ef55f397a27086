"""e2e time of chamfer_distance on pinned host arrays vs the number of upload chunks (f3d_chamfer_pipe_run)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flux3d_b200 as f3d

B, N, M = 32, 4096, 4096
A = torch.from_numpy(np.random.default_rng(201).random((B, N, 3), dtype=np.float32)).pin_memory()
Bc = torch.from_numpy(np.random.default_rng(202).random((B, M, 3), dtype=np.float32)).pin_memory()
for chunks in (4, 16, 32, 64):
    for _ in range(5):
        f3d.chamfer_forward_host(A, Bc, uploaders=chunks).item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        v = f3d.chamfer_forward_host(A, Bc, uploaders=chunks).item()
    dt = (time.perf_counter() - t0) / 200
    print(f"chunks={chunks:2d}  {dt*1e6:7.1f} us/step  {B*N*M/dt:.3e} pairs/s  loss={v:.10f}")
dA, dB = A.cuda(), Bc.cuda()
out = (torch.empty(3, device="cuda"), None, None)
for _ in range(5):
    f3d.chamfer_forward_raw(dA, dB, 1.0, 1.0, want_indices=False, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    f3d.chamfer_forward_raw(dA, dB, 1.0, 1.0, want_indices=False, out=out)
e1.record(); torch.cuda.synchronize()
print(f"device-resident back-to-back (warm L2): {e0.elapsed_time(e1)/200*1e3:.1f} us/step")
