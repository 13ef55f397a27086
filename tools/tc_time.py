"""Development aid: time the chamfer forward (tensor-core sweep vs CUDA-core sweep; whole step and sweep only) on a few
shapes, back to back and with the L2 flushed, and print the tensor-core path's diagnostics."""
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d  # noqa: E402

SHAPES = [(32, 4096, 4096), (32, 8192, 8192), (16, 10000, 10000), (64, 2048, 2048), (8, 4096, 4096), (1, 16384, 16384), (2, 1024, 1024)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]


def timed(fn, reps, flush=None):
    for _ in range(3):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for e0, e1 in evs:
        if flush is not None:
            flush.zero_()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in evs)
    return t[len(t) // 2], t[0]


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, N, M in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand((B, N, 3), generator=g, device="cuda")
    Bc = torch.rand((B, M, 3), generator=g, device="cuda")
    out = (torch.empty(3, device="cuda"), None, None)
    res = {}
    for name, fl in (("tc", f3d.FLAG_TENSOR), ("cuda", f3d.FLAG_CUDA_CORES)):
        for sub, extra in (("step", 0), ("sweep", f3d.FLAG_SWEEP_ONLY)):
            fn = lambda: f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, want_indices=False, flags=fl | extra, out=out)
            res[(name, sub, "cold")] = timed(fn, 20, flush)
            res[(name, sub, "hot")] = timed(fn, 20)
    l_tc = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=f3d.FLAG_TENSOR)
    torch.cuda.synchronize()
    ws = f3d._lib.workspace(("chamfer", B, N, M), 256, A.device)
    hdr = ws[:12].cpu().numpy().view(np.int32)
    l_cc = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=f3d.FLAG_CUDA_CORES)
    torch.cuda.synchronize()
    same = bool(torch.equal(l_tc[2], l_cc[2]) and torch.equal(l_tc[3], l_cc[3]))
    pairs = B * N * M
    print(f"B={B} N={N} M={M}: loss tc {l_tc[0].item():.9g} cuda {l_cc[0].item():.9g} indices equal {same}; ambiguous rows {hdr[1]} "
          f"({100.0 * hdr[1] / (B * (N + M)):.3f} %), bound violations {hdr[2]}")
    for name in ("tc", "cuda"):
        s = res[(name, "step", "cold")][0]
        print(f"   {name:5s} step cold {s:8.1f} us (min {res[(name, 'step', 'cold')][1]:.1f})  hot {res[(name, 'step', 'hot')][0]:8.1f} | "
              f"sweep cold {res[(name, 'sweep', 'cold')][0]:8.1f}  hot {res[(name, 'sweep', 'hot')][0]:8.1f} | {pairs / (s * 1e-6):.3e} pairs/s")
