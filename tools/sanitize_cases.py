"""Small invocations of EVERY kernel of libflux3d_b200.so, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
tools/sanitize.sh runs this file under each tool.  Shapes are small (the tools slow kernels down 10-100x) but cover every
protocol: the CUDA-core filter sweep with its PDL finalize and completion counters, its prepared-operand (TMA) path, upload
mode (start tickets, arrival flags), the tensor-core sweep (TMA ring, mbarrier pipelines, TMEM, publisher, done flags), its
finalize / cleanup pair, its upload + prepare grid, the exact sweep, chamfer backward (sorted gather and RED.ADD), kNN (both
kernels, emit kernels, MLP layout), mesh kernels with pullbacks, converters, sample_points with its pullback."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import flux3d_b200 as f3d  # noqa: E402

rng = np.random.default_rng(0)


def cloud(*shape):
    return torch.from_numpy(rng.random(shape, dtype=np.float32)).cuda()


# chamfer: CUDA-core filter sweep (+ PDL finalize), exact sweep, FMA mode, tensor-core sweep forced on ragged shapes
for (B, N, M) in ((2, 300, 517), (1, 1, 1), (3, 1025, 260)):
    A, Bc = cloud(B, N, 3), cloud(B, M, 3)
    ref = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=f3d.FLAG_EXACT_SWEEP)
    for fl in (f3d.FLAG_CUDA_CORES, f3d.FLAG_TENSOR, f3d.FLAG_FMA):
        out = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=fl)
        torch.cuda.synchronize()
        if fl != f3d.FLAG_FMA:
            assert torch.equal(out[2], ref[2]) and torch.equal(out[3], ref[3]), (B, N, M, fl)
# ties: every row ambiguous -> cleanup kernel's supertile scans, list overflow path
P = torch.full((1, 700, 3), 0.25, device="cuda")
out = f3d.chamfer_forward_raw(P, P[:, :300].contiguous(), 1.0, 1.0, flags=f3d.FLAG_TENSOR)
torch.cuda.synchronize()
assert out[0].item() == 0.0
# prepared-operand path of the CUDA-core sweep (>= 32 row blocks), one batch element
A, Bc = cloud(1, 8200, 3), cloud(1, 300, 3)
f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0, flags=f3d.FLAG_CUDA_CORES)
# host arrays: in-grid upload of both sweeps, pageable fallback
hA, hB = rng.random((3, 400, 3), dtype=np.float32), rng.random((3, 333, 3), dtype=np.float32)
pA, pB = torch.from_numpy(hA).pin_memory(), torch.from_numpy(hB).pin_memory()
l0 = f3d.chamfer_forward_host(pA, pB, to_host=True, flags=f3d.FLAG_CUDA_CORES)
l1 = f3d.chamfer_forward_host(pA, pB, to_host=True, flags=f3d.FLAG_TENSOR)
l2 = f3d.chamfer_forward_host(hA, hB, flags=f3d.FLAG_TENSOR)
torch.cuda.synchronize()
assert abs(l0.item() - l1.item()) <= 1e-6 * l0.item() and l2.item() == l1.item()
# chamfer backward: counting-sort gather (also a target with > 32 sources) and (beyond 24576 points) the RED.ADD path
for (B, N, M) in ((2, 500, 70), (1, 2100, 40), (1, 24600, 64)):
    tA, tB = cloud(B, N, 3).requires_grad_(True), cloud(B, M, 3).requires_grad_(True)
    f3d.chamfer_distance(tA, tB).backward()
# kNN graph: CUDA-core kernel (narrow / wide), tensor-core kernel, gathered / edge outputs, MLP layout, gradient
# knn_gram (TMA-fed Gram filter: split rows F <= 4, plain rows; its prepare kernel and in-CTA exact scan: the lattice cloud overflows the slots)
for (B, N, F, K, fl) in ((2, 200, 3, 10, 0), (1, 300, 20, 33, 0), (2, 256, 64, 20, 0), (1, 130, 3, 5, f3d.FLAG_TENSOR), (2, 600, 3, 10, 0), (1, 520, 40, 9, 0)):
    X = torch.from_numpy(rng.standard_normal((B, N, F)).astype(np.float32)).cuda()
    f3d.knn_graph(X, K, want_dist=True, want_gathered=True, want_edge=True, flags=fl)
    f3d.knn_graph(X, K, want_edge=True, mlp_layout=True, flags=fl)
Xl = torch.from_numpy(rng.integers(0, 3, size=(1, 600, 3)).astype(np.float32)).cuda()
assert int(f3d.knn_graph(Xl, 10, want_stats=True)["stats"][7]) == 2
Xg = torch.randn(1, 150, 6, device="cuda", requires_grad=True)
f3d.edgeconv_features(Xg, 7, mlp_layout=True).sum().backward()
# mesh kernels on the reference's teapot fixture
m = f3d.load_trimesh(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "teapot.obj"))
m.compute_verts_normals_packed(0); m.compute_verts_normals_packed(1); m.compute_faces_normals_packed(); m.compute_faces_areas_packed()
m.get_verts_padded(); m.faces_padded_device()
off = torch.zeros_like(m.get_verts_packed(), requires_grad=True)
m2 = f3d.offset(m, off)
loss = f3d.laplacian_loss(m2) + f3d.edge_loss(m2) + f3d.sample_points(m2, 500, seed=3).sum()
loss.backward()
torch.cuda.synchronize()
print("sanitize_cases: all kernels ran")
