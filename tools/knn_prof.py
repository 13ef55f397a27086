"""Development aid: phase clocks of CTA 0 of knn_gram_kernel (-DF3D_KNN_PROF build)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flux3d_b200 as f3d
for F in (3, 64):
    X = torch.randn((32, 1024, F), device="cuda")
    for _ in range(3):
        st = f3d.knn_graph(X, 20, want_stats=True, flags=f3d.FLAG_TENSOR)["stats"].cpu().numpy()
    print(f"F={F}: overflow rows {st[0]} (cap {st[3]}, segment/threshold {st[4]}, short {st[5]}); clocks since CTA start: prologue {st[8]}, pass 1 {st[9]}, "
          f"threshold {st[10]}, pass 2 {st[12]}, ranking {st[13]}")
