"""Development aid: timeline of the in-grid finalize from a -DF3D_EXP_CLOCK build (build/variants/<argv[1]>):
when tiles end, how many finalize slices each CTA takes and how long it stays, what is left after the last tile."""
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
sw = np.zeros((8192, 8), np.int64); fx0 = np.zeros((8192, 8), np.int64); fx = np.zeros((8192, 8), np.int64)
assert L.f3d_debug_read_fin(fx0.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(fx0.nbytes)) == 0
f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, want_indices=False, out=out)
torch.cuda.synchronize()
assert L.f3d_debug_read(sw.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(sw.nbytes)) == 0
assert L.f3d_debug_read_fin(fx.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(fx.nbytes)) == 0
sw = sw[:2048]; fx = fx[:2048]; ph = (fx[:, 2:7] - fx0[:2048, 2:7])
T0 = sw[:, 4].min()
start, tile_end, exit_, nfin = (sw[:, 4] - T0) / 1e3, (sw[:, 5] - T0) / 1e3, (fx[:, 0] - T0) / 1e3, fx[:, 1]
print("last CTA start %.1f us, last tile end %.1f us, last CTA exit %.1f us" % (start.max(), tile_end.max(), exit_.max()))
print("tile time mean %.1f us; CTAs that finalize: %d of 2048; slices per finalizing CTA mean %.2f max %d" %
      ((tile_end - start).mean(), (nfin > 0).sum(), nfin[nfin > 0].mean(), nfin.max()))
stay = exit_ - tile_end
print("time after own tile: mean %.2f us, per slice %.2f us (finalizing CTAs), max stay %.1f us" %
      (stay.mean(), (stay[nfin > 0] / nfin[nfin > 0]).mean(), stay.max()))
lw = tile_end > np.sort(tile_end)[-272]
print("last wave (272 latest tiles): slices taken %d, stay mean %.1f us" % (nfin[lw].sum(), stay[lw].mean()))
for t in (60, 80, 100, 110, 120, 125, 130, 135, 140, 150, 160):
    print("  t=%3d us: tiles done %4d, slices done by exited CTAs %4d" % (t, (tile_end <= t).sum(), nfin[exit_ <= t].sum()))
tot = nfin.sum()
print("cycles per slice by phase (thread 0): claim.begin+phase1 %.0f, probe+phase2 %.0f, claim.finish %.0f, vote %.0f, phase3 %.0f  (sum %.0f cyc = %.2f us at 1.9 GHz)" %
      (*(ph.sum(0) / tot), ph.sum() / tot, ph.sum() / tot / 1900))
