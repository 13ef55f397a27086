"""Development aid: time f3d_chamfer_fwd (sweep-only and full step) for every library under build/variants/
(each in its own process via FLUX3D_B200_LIB).  bench.py is the contract; this only ranks kernel variants."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import torch, flux3d_b200 as f3d
def t(B, N, M, flags, iters=30):
    A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
    out = (torch.empty(3, device="cuda"), None, None)
    for _ in range(5): f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=flags, want_indices=False, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=flags, want_indices=False, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
print(os.path.basename(os.environ.get("FLUX3D_B200_LIB", "default")),
      "cfg2 sweep %%.1f us full %%.1f us | 8192 sweep %%.1f full %%.1f | 1024x2 full %%.1f" %% (t(32,4096,4096,2), t(32,4096,4096,0), t(32,8192,8192,2), t(32,8192,8192,0), t(2,1024,1024,0)), flush=True)
''' % ROOT
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib: env["FLUX3D_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", CHILD], env=env)
