"""Development aid: per-block timeline of chamfer_tc_finalize_kernel against the end of the sweep (-DF3D_TC_PROF build)."""
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
nblk = min(8192, B * ((N + 255) // 256 + (M + 255) // 256) * 4)
fin = np.zeros(8192 * 8, np.uint64); sw = np.zeros(148 * 16, np.int64)
for fn, buf in ((L.f3d_debug_read_fin, fin), (L.f3d_debug_read_tc, sw)):
    fn.argtypes = [C.c_void_p, C.c_size_t]
    assert fn(buf.ctypes.data, buf.nbytes) == 0
f = fin.reshape(8192, 8)[:nblk].astype(np.float64)
send = sw.reshape(148, 16)[:, 15].astype(np.float64)
t0 = f[:, 0].min()
print(f"{B}x{N}x{M}: {nblk} finalize blocks; sweep CTAs end at {(send.min() - t0) / 1e3:.1f} .. {(send.max() - t0) / 1e3:.1f} us after the first finalize block started")
names = ["wait for flag", "phase 1 (loads, certify)", "phase 2 (staged rescan)", "list + block sum"]
for i, n in enumerate(names):
    d = (f[:, i + 1] - f[:, i]) / 1e3
    print(f"  {n:28s} mean {d.mean():7.2f} us  median {np.median(d):7.2f}  p95 {np.percentile(d, 95):7.2f}  max {d.max():7.2f}")
work = (f[:, 4] - f[:, 1]) / 1e3
print(f"  work per block (after the flag)  mean {work.mean():.2f} us; last block ends {(f[:, 4].max() - t0) / 1e3:.1f} us; blocks ending after the sweep: {(f[:, 4] > send.max()).sum()}")
late = np.sort(f[:, 4])[-5:] - send.max()
print("  last five block ends relative to the sweep end (us):", np.round(late / 1e3, 1))
order = np.argsort(f[:, 0])
print("  start times of blocks (us) every 256th:", np.round((f[order[::256], 0] - t0) / 1e3, 1))

cl = np.zeros(4096 * 4, np.uint64)
L.f3d_debug_read_clean.argtypes = [C.c_void_p, C.c_size_t]
assert L.f3d_debug_read_clean(cl.ctypes.data, cl.nbytes) == 0
c = cl.reshape(4096, 4).astype(np.float64)
used = c[:, 0] > 0
c = c[used]
last = c[:, 3].max()
print(f"  cleanup: {used.sum()} blocks; first block starts {(c[:, 0].min() - t0) / 1e3:.1f} us, last starts {(c[:, 0].max() - t0) / 1e3:.1f}; row work done by {(c[:, 1].max() - t0) / 1e3:.1f}; "
      f"loss written at {(last - t0) / 1e3:.1f} us (finalize ended {(f[:, 4].max() - t0) / 1e3:.1f}, sweep {(send.max() - t0) / 1e3:.1f})")
w = (c[:, 1] - c[:, 0]) / 1e3
print(f"  cleanup row work per block: mean {w.mean():.2f} us, p95 {np.percentile(w, 95):.2f}, max {w.max():.2f}")
