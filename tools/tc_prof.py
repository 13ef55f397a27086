"""Development aid: per-role wait / work cycles of chamfer_tc_sweep_kernel (a -DF3D_TC_PROF build, FLUX3D_B200_LIB)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d
B, N, M = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x4096x4096").split("x"))
A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
for _ in range(3):
    f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, want_indices=False, flags=f3d.FLAG_TENSOR | f3d.FLAG_SWEEP_ONLY)
torch.cuda.synchronize()
L = f3d._lib.lib()
buf = np.zeros(148 * 16, np.int64)
L.f3d_debug_read_tc.argtypes = [C.c_void_p, C.c_size_t]
assert L.f3d_debug_read_tc(buf.ctypes.data, buf.nbytes) == 0
p = buf.reshape(148, 16).astype(np.float64)
names = ["mma: wait a_full", "mma: wait full_b (converter)", "mma: wait tempty[0] (read-out r0)", "mma: wait tempty[1] (read-out r1)", "mma: issue+commit",
         "conv: wait cfull (TMA)", "conv: wait empty_b (MMA)", "conv: convert+store+fence", "read-out warp 0: wait tfull (MMA)", "read-out warp 0: read-out", "read-out warp 0: between tiles",
         "read-out warp 8: wait tfull (MMA)", "read-out warp 8: read-out", "read-out warp 8: between tiles"]
tiles = (B * (N + M) // 256) * (M // 256) / 148.0
print(f"{B}x{N}x{M}: ~{tiles:.0f} candidate tiles per CTA; cycles per tile (mean over CTAs):")
for i, n in enumerate(names):
    print(f"  {n:40s} {p[:, i].mean() / tiles:9.1f}")
