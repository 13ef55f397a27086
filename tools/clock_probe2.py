"""Development aid: phase totals per CTA of the persistent sweep (from a -DF3D_EXP_CLOCK build)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["FLUX3D_B200_LIB"] = os.path.join(ROOT, "build", "variants", "lib_clock.so")
sys.path.insert(0, ROOT)
import torch, flux3d_b200 as f3d
L = ctypes.CDLL(os.environ["FLUX3D_B200_LIB"])
for (B, N, M) in ((32, 4096, 4096), (32, 8192, 8192)):
    A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
    out = (torch.empty(3, device="cuda"), None, None)
    for _ in range(3):
        f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=2, want_indices=False, out=out)
    torch.cuda.synchronize()
    buf = np.zeros((8192, 8), np.int64)
    assert L.f3d_debug_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes)) == 0
    d = buf[:592]
    print(B, N, M, "per-CTA cycles: wait %.0f transform %.0f loop %.0f epilogue %.0f total %.0f (max total %.0f)" %
          (d[:, 0].mean(), d[:, 1].mean(), d[:, 2].mean(), d[:, 3].mean(), d[:, 4].mean(), d[:, 4].max()))
