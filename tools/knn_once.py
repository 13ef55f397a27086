import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, flux3d_b200 as f3d
F = int(sys.argv[1]) if len(sys.argv) > 1 else 3
X = torch.randn((32, 1024, F), device="cuda")
for _ in range(3):
    out = f3d.knn_graph(X, 20, want_stats=True)
torch.cuda.synchronize()
print(out["stats"].cpu().numpy())
