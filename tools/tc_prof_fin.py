"""Development aid: the tail of the tensor-core chamfer step (-DF3D_TC_PROF build) — when the read-out warps, the certifier warps
(CTA end) and the cleanup kernel finish, on the device's global timer."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d
B, N, M = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32x4096x4096").split("x"))
A = torch.rand((B, N, 3), device="cuda"); Bc = torch.rand((B, M, 3), device="cuda")
for _ in range(3):
    f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, want_indices=False, flags=f3d.FLAG_TENSOR)
torch.cuda.synchronize()
L = f3d._lib.lib()
sw = np.zeros(148 * 16, np.int64); cl = np.zeros(4096 * 4, np.uint64)
for fn, buf in ((L.f3d_debug_read_tc, sw), (L.f3d_debug_read_clean, cl)):
    fn.argtypes = [C.c_void_p, C.c_size_t]
    assert fn(buf.ctypes.data, buf.nbytes) == 0
s = sw.reshape(148, 16).astype(np.float64)
rend, cend = s[:, 14], s[:, 15]
c = cl.reshape(4096, 4).astype(np.float64)
c = c[c[:, 0] > 0]
t0 = rend.max()
print(f"{B}x{N}x{M}: times in us relative to the LAST read-out warp's end")
print(f"  read-out ends   {(rend.min() - t0) / 1e3:7.1f} .. 0.0")
lag = (cend - rend) / 1e3
print(f"  CTA end - its read-out end (certifier lag): mean {lag.mean():.2f}, max {lag.max():.2f}; last CTA ends at {(cend.max() - t0) / 1e3:.1f}")
print(f"  cleanup: {len(c)} blocks; past griddepcontrol.wait at {(c[:, 0].min() - t0) / 1e3:.1f} .. {(c[:, 0].max() - t0) / 1e3:.1f}; row work done by {(c[:, 1].max() - t0) / 1e3:.1f}; "
      f"last ticket {(c[:, 2].max() - t0) / 1e3:.1f}; loss written at {(c[:, 3].max() - t0) / 1e3:.1f}")
ce = np.zeros(148 * 8, np.int64)
L.f3d_debug_read_cert.argtypes = [C.c_void_p, C.c_size_t]
assert L.f3d_debug_read_cert(ce.ctypes.data, ce.nbytes) == 0
ce = ce.reshape(148, 8).astype(np.float64)
nit = B * ((N + 255) // 256 + (M + 255) // 256) / 148
names = ["wait for the read-out", "merge + phase 1", "phase 2: first chunk", "phase 2: second chunk", "check, sums, lists"]
print(f"  certifier warp 0, clocks per item (mean over CTAs, {nit:.1f} items per CTA):")
for i, n in enumerate(names):
    print(f"    {n:36s} {ce[:, i].mean() / nit:9.0f}")
st, cee = ce[:, 5], ce[:, 6]
print(f"  CTA starts {(st.min() - t0) / 1e3:.1f} .. {(st.max() - t0) / 1e3:.1f}; certifier warp 0 ends {(cee.min() - t0) / 1e3:.1f} .. {(cee.max() - t0) / 1e3:.1f}; "
      f"per CTA certifier end - read-out end: mean {((cee - rend) / 1e3).mean():.2f} max {((cee - rend) / 1e3).max():.2f}")
print("  raw CTA 0:", [(x - t0) / 1e3 for x in (st[0], rend[0], cee[0], cend[0])])
