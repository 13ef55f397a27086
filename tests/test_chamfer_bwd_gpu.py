"""GPU parity of the Chamfer pullback (f3d_chamfer_bwd) and of the public differentiable chamfer_distance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_backward_vs_oracle(f3d, oracle):
    for (B, N, M, seed) in [(2, 1000, 500, 3), (1, 64, 257, 4), (4, 2048, 2048, 5)]:
        A = np.random.default_rng(seed).random((B, N, 3), dtype=np.float32)
        Bc = np.random.default_rng(seed + 1).random((B, M, 3), dtype=np.float32)
        tA = torch.from_numpy(A).cuda().requires_grad_(True)
        tB = torch.from_numpy(Bc).cuda().requires_grad_(True)
        loss = f3d.chamfer_distance(tA, tB, w1=0.7, w2=1.3)
        (loss * 2.0).backward()
        _, nA, nB, _ = oracle.chamfer_distance(A, Bc, 0.7, 1.3, return_all=True)
        gA, gB = oracle.chamfer_backward(A, Bc, nA, nB, 0.7, 1.3, gout=2.0)
        # reference bar: atol 1e-2, rtol 1e-3 (test/metrics.jl:112-114); ours is float32-roundoff tight
        assert np.allclose(tA.grad.cpu().numpy(), gA, rtol=1e-4, atol=1e-9)
        assert np.allclose(tB.grad.cpu().numpy(), gB, rtol=1e-4, atol=1e-9)


def test_backward_vs_torch_autograd(f3d):
    """The reference's gradient test: gradient of chamfer_distance ≈ gradient of naive_chamfer."""
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((2, 1000, 3), generator=g, device="cuda", requires_grad=True)
    y = torch.rand((2, 500, 3), generator=g, device="cuda", requires_grad=True)
    f3d.chamfer_distance(x, y).backward()
    x2 = x.detach().double().requires_grad_(True)
    y2 = y.detach().double().requires_grad_(True)
    P = torch.cdist(x2, y2).pow(2)
    (P.min(2).values.mean() + P.min(1).values.mean()).backward()
    assert torch.allclose(x.grad.double(), x2.grad, rtol=1e-3, atol=1e-2)
    assert torch.allclose(y.grad.double(), y2.grad, rtol=1e-3, atol=1e-2)
    assert torch.allclose(x.grad.double(), x2.grad, rtol=1e-4, atol=1e-8)


def test_pointcloud_front_end(f3d, oracle):
    """chamfer_distance(::PointCloud, ::PointCloud; w1, w2) and the 2-D array overloads (metrics/pcloud.jl:11-37)."""
    A = np.random.default_rng(1).random((300, 3), dtype=np.float32)
    B = np.random.default_rng(2).random((200, 3), dtype=np.float32)
    ref = float(oracle.chamfer_distance(A, B, 2.0, 0.5))
    for a, b in ((A, B), (f3d.PointCloud(A), f3d.PointCloud(B)), (torch.from_numpy(A), torch.from_numpy(B).cuda())):
        got = float(f3d.chamfer_distance(a, b, w1=2.0, w2=0.5).item())
        assert abs(got - ref) <= 1e-5 * ref
    nnA, nnB = f3d.nearest_neighbors(A, B)
    oA, oB = oracle.nearest_neighbors(A[None], B[None])
    assert np.array_equal(nnA.cpu().numpy(), oA) and np.array_equal(nnB.cpu().numpy(), oB)
    with pytest.raises(ValueError):
        f3d.chamfer_distance(np.zeros((2, 5, 3), np.float32), np.zeros((3, 5, 3), np.float32))


def test_backward_is_bitwise_repeatable_and_covers_both_paths(f3d, oracle):
    """Clouds of up to 24576 points take the counting-sort gather pullback (one launch, no global atomics): bitwise identical
    reruns, also with many sources pulling on one target (> 32: summed by the whole block).  Larger clouds take the RED.ADD
    fallback: still within float32 round-off of the oracle."""
    rng = np.random.default_rng(11)
    A = rng.random((3, 5000, 3), dtype=np.float32)
    Bc = rng.random((3, 40, 3), dtype=np.float32)          # every point of B is the target of ~125 points of A
    grads = []
    for _ in range(3):
        tA = torch.from_numpy(A).cuda().requires_grad_(True)
        tB = torch.from_numpy(Bc).cuda().requires_grad_(True)
        f3d.chamfer_distance(tA, tB, w1=0.5, w2=2.0).backward()
        grads.append((tA.grad.clone(), tB.grad.clone()))
    assert all(torch.equal(g[0], grads[0][0]) and torch.equal(g[1], grads[0][1]) for g in grads[1:])
    _, nA, nB, _ = oracle.chamfer_distance(A, Bc, 0.5, 2.0, return_all=True)
    gA, gB = oracle.chamfer_backward(A, Bc, nA, nB, 0.5, 2.0)
    assert np.allclose(grads[0][0].cpu().numpy(), gA, rtol=1e-4, atol=1e-10)
    assert np.allclose(grads[0][1].cpu().numpy(), gB, rtol=2e-4, atol=1e-10)
    # several target slices per element (B = 1), segments of ~30 sources (selection in the segment); then beyond 24576 points
    # per cloud: the two-launch path
    for n2, m2, repeatable in ((9000, 300, True), (25000, 300, False)):
        A2 = rng.random((1, n2, 3), dtype=np.float32)
        B2 = rng.random((1, m2, 3), dtype=np.float32)
        got = []
        for _ in range(2):
            tA = torch.from_numpy(A2).cuda().requires_grad_(True)
            tB = torch.from_numpy(B2).cuda().requires_grad_(True)
            f3d.chamfer_distance(tA, tB).backward()
            got.append((tA.grad.clone(), tB.grad.clone()))
        if repeatable:
            assert torch.equal(got[0][0], got[1][0]) and torch.equal(got[0][1], got[1][1])
        _, nA, nB, _ = oracle.chamfer_distance(A2, B2, return_all=True)
        gA, gB = oracle.chamfer_backward(A2, B2, nA, nB)
        assert np.allclose(got[0][0].cpu().numpy(), gA, rtol=1e-4, atol=1e-10)
        assert np.allclose(got[0][1].cpu().numpy(), gB, rtol=2e-4, atol=1e-10)
    # every point on one spot: one target takes every source (lowest index on ties)
    P = torch.full((2, 700, 3), 0.5, device="cuda")
    tA = (P + 0.0).requires_grad_(True)
    tB = (P[:, :300] * 1.0 + 0.25).requires_grad_(True)
    f3d.chamfer_distance(tA, tB).backward()
    assert torch.isfinite(tA.grad).all() and torch.isfinite(tB.grad).all()
    assert torch.allclose(tA.grad.sum(), -tB.grad.sum(), rtol=1e-4)   # the loss only depends on differences
