"""Development aid: per-block phase timeline of chamfer_filter_finalize_kernel from a -DF3D_EXP_CLOCK build
(build/variants/<name>.so given as argv[1]) relative to the sweep kernel's CTAs."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["FLUX3D_B200_LIB"] = os.path.join(ROOT, "build", "variants", sys.argv[1])
sys.path.insert(0, ROOT)
import torch, flux3d_b200 as f3d
L = ctypes.CDLL(os.environ["FLUX3D_B200_LIB"])
B, N, M = 32, 4096, 4096
A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
out = (torch.empty(3, device="cuda"), None, None)
for _ in range(3):
    f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, want_indices=False, out=out)
torch.cuda.synchronize()
sw = np.zeros((8192, 8), np.int64); fi = np.zeros((2048, 8), np.int64)
assert L.f3d_debug_read(sw.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(sw.nbytes)) == 0
assert L.f3d_debug_read_fin(fi.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(fi.nbytes)) == 0
sw = sw[:2048]; fi = fi[:1024]
T0 = sw[:, 4].min()
print("sweep: first CTA start 0, last CTA start %.1f us, last CTA end %.1f us" % ((sw[:, 4].max() - T0) / 1e3, (sw[:, 5].max() - T0) / 1e3))
f = (fi[:, :6] - T0) / 1e3
print("finalize blocks: first start %.1f, median start %.1f, last start %.1f; last end %.1f us" % (f[:, 0].min(), np.median(f[:, 0]), f[:, 0].max(), f[:, 5].max()))
names = ["flag wait", "phase 1", "phase 2", "phase 3", "block reduce"]
for k in range(5):
    dt = f[:, k + 1] - f[:, k]
    print("  %-12s mean %6.2f  p50 %6.2f  p95 %6.2f  max %6.2f us" % (names[k], dt.mean(), np.median(dt), np.percentile(dt, 95), dt.max()))
life = f[:, 5] - f[:, 0]
print("  block lifetime mean %.2f max %.2f us" % (life.mean(), life.max()))
order = np.argsort(f[:, 5])
print("  last 8 blocks to finish (id, start, flag-done, p1, p2, p3, end, sm):")
for i in order[-8:]:
    print("   ", i, " ".join("%.1f" % v for v in f[i]), fi[i, 7])
# how many blocks are resident over time
ts = np.linspace(0, f[:, 5].max(), 30)
print("  resident finalize blocks at t:", " ".join("%d" % ((f[:, 0] <= t) & (f[:, 5] > t)).sum() for t in ts))
