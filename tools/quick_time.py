"""Quick device-timed throughput of the chamfer path (development aid; bench.py is the contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flux3d_b200 as f3d

def run(B, N, M, flags, iters=20):
    A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
    for _ in range(5):
        f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=flags)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=flags)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"B={B} N={N} M={M} flags={flags}: {ms*1e3:.1f} us/call  {B*N*M/ms/1e9*1e3:.1f} Gpairs/s", flush=True)

if __name__ == "__main__":
    for flags in (0, 1):
        run(32, 4096, 4096, flags)
        run(32, 8192, 8192, flags)
        run(2, 1024, 1024, flags)
