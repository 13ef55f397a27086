"""Development aid: kNN graph timing and parity between the kernels of the library (knn_gram / knn_tc / CUDA-core sweep)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d  # noqa: E402

SHAPES = [(32, 1024, 3, 20), (32, 1024, 64, 20), (32, 1024, 16, 10), (8, 2048, 3, 20), (8, 2000, 64, 16), (4, 700, 32, 20)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for e0, e1 in evs:
        flush.zero_()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in evs)
    return t[len(t) // 2]


for B, N, F, K in SHAPES:
    X = torch.randn((B, N, F), device="cuda")
    ref = f3d.knn_graph(X, K, want_dist=True, flags=f3d.FLAG_EXACT_SWEEP)
    got = f3d.knn_graph(X, K, want_dist=True, want_stats=True, flags=f3d.FLAG_TENSOR)
    torch.cuda.synchronize()
    ok = torch.equal(got["idx"], ref["idx"]) and torch.equal(got["dist"], ref["dist"])
    st = got["stats"].cpu().numpy()
    t_new = timed(lambda: f3d.knn_graph(X, K, flags=f3d.FLAG_TENSOR))
    t_def = timed(lambda: f3d.knn_graph(X, K))
    t_old = timed(lambda: f3d.knn_graph(X, K, flags=f3d.FLAG_EXACT_SWEEP))
    print(f"B={B} N={N} F={F} K={K}: equal {ok}; exact-scan rows {st[0]}, candidates/query {st[1] / (B * N):.1f}; "
          f"tensor {t_new:.1f} us, default {t_def:.1f} us, exact sweep {t_old:.1f} us", flush=True)
