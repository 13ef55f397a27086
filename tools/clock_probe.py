"""Development aid: per-CTA phase timing of chamfer_filter_sweep_kernel from a -DF3D_EXP_CLOCK build
(build/variants/lib_clock.so): prologue / main loop / epilogue cycles, CTA start/end times, CTAs per SM."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["FLUX3D_B200_LIB"] = os.path.join(ROOT, "build", "variants", "lib_clock.so")
sys.path.insert(0, ROOT)
import torch, flux3d_b200 as f3d
L = ctypes.CDLL(os.environ["FLUX3D_B200_LIB"])
B, N, M = 32, 4096, 4096
A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
out = (torch.empty(3, device="cuda"), None, None)
for _ in range(3):
    f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=2, want_indices=False, out=out)
torch.cuda.synchronize()
n = 2048
buf = np.zeros((8192, 8), np.int64)
assert L.f3d_debug_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes)) == 0
d = buf[:n]
pro, loop, epi = d[:, 1] - d[:, 0], d[:, 2] - d[:, 1], d[:, 3] - d[:, 2]
t0, t1, sm = d[:, 4] - d[:, 4].min(), d[:, 5] - d[:, 4].min(), d[:, 6]
print("cycles  prologue mean %.0f (p95 %.0f)  loop mean %.0f (min %.0f p95 %.0f)  epilogue mean %.0f" %
      (pro.mean(), np.percentile(pro, 95), loop.mean(), loop.min(), np.percentile(loop, 95), epi.mean()))
print("kernel span %.1f us; CTA lifetime mean %.1f us; first-wave starts within %.1f us" %
      (t1.max() / 1e3, (t1 - t0).mean() / 1e3, np.sort(t0)[591] / 1e3))
order = np.argsort(t0)
for k in (0, 591, 592, 1183, 1184, 1775, 1776, 2047):
    i = order[k]; print("  start-rank %4d: start %.1f us end %.1f us  loop %d cyc  sm %d" % (k, t0[i] / 1e3, t1[i] / 1e3, loop[i], sm[i]))
cnt = np.bincount(sm, minlength=148); print("CTAs per SM: min %d max %d" % (cnt.min(), cnt.max()))
# loop speed vs time-of-start (are later CTAs, running with fewer neighbours, faster?)
for lo, hi in ((0, 592), (592, 1184), (1184, 1776), (1776, 2048)):
    idx = order[lo:hi]; print("  wave %d-%d: loop mean %.0f cyc, lifetime %.1f us" % (lo, hi, loop[idx].mean(), (t1[idx] - t0[idx]).mean() / 1e3))
