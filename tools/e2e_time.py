"""Development aid: wall-clock step time of chamfer_distance(pinned host arrays) -> host scalar (f3d_chamfer_pipe_run), per sweep."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d
B, N, M = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x4096x4096").split("x"))
A = torch.from_numpy(np.random.default_rng(201).random((B, N, 3), dtype=np.float32)).pin_memory()
Bc = torch.from_numpy(np.random.default_rng(202).random((B, M, 3), dtype=np.float32)).pin_memory()
for name, fl in (("tensor-core sweep", f3d.FLAG_TENSOR), ("CUDA-core sweep", f3d.FLAG_CUDA_CORES)):
    for ups in (0, 4, 8, 16, 32):
        for _ in range(5):
            l = f3d.chamfer_forward_host(A, Bc, to_host=True, flags=fl, uploaders=ups)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            l = f3d.chamfer_forward_host(A, Bc, to_host=True, flags=fl, uploaders=ups)
        dt = (time.perf_counter() - t0) / 50
        print(f"{name:18s} uploaders {ups:3d}: {dt * 1e6:7.1f} us/step  loss {l.item():.9g}", flush=True)
