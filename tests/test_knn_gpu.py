"""GPU parity: f3d_knn_graph (through the C ABI) vs the oracle — kNN indices bit-exact, in order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _normalized(rng, B, N):
    X = rng.standard_normal((B, N, 3)).astype(np.float32)
    X = X - X.mean(axis=1, keepdims=True)
    return (X / X.std(axis=(1, 2), keepdims=True)).astype(np.float32)  # NormalizePointCloud (pcloud_func.jl:16-22)


@pytest.mark.parametrize("B,N,F,K", [
    (32, 1024, 3, 20),   # BASELINE configs[2] (cfg3): DGCNN EdgeConv1 shape
    (32, 1024, 3, 10),   # the reference's own default K (models/dgcnn.jl:99)
    (4, 1024, 64, 20),   # EdgeConv2 shape (F = 64 features)
    (2, 100, 5, 7),      # F % 4 != 0, ragged N
    (1, 33, 3, 32),      # K+1 = 33 → two-slot list, K = N-1
    (1, 70, 2, 63),      # maximum K
    (3, 2, 3, 1),        # smallest legal problem
    (1, 257, 130, 9),    # wide features
])
@pytest.mark.parametrize("flags", [0, 4, 8])  # 0: default dispatch; 4: all-exact CUDA-core kernel; 8: tensor-core filter wherever the shape allows
def test_knn_parity(f3d, oracle, B, N, F, K, flags):
    rng = np.random.default_rng(301 + N + F + K)
    X = _normalized(rng, B, N) if F == 3 else rng.standard_normal((B, N, F)).astype(np.float32)
    out = f3d.knn_graph(torch.from_numpy(X).cuda(), K, want_dist=True, want_gathered=True, want_edge=True, flags=flags)
    torch.cuda.synchronize()
    idx, dist, gat = oracle.knn_graph(X, K, want_dist=True, want_gathered=True)
    assert np.array_equal(out["idx"].cpu().numpy(), idx)
    assert np.array_equal(out["dist"].cpu().numpy(), dist)
    assert np.array_equal(out["gathered"].cpu().numpy(), gat)
    assert np.array_equal(out["edge"].cpu().numpy(), oracle.edge_features(X, idx))


def test_knn_duplicates_and_ties(f3d, oracle):
    """Exact duplicates: the first hit is dropped BY POSITION (dgcnn.jl:6), so the higher-indexed twin keeps
    itself in its list; lattice points: ties resolve by index."""
    rng = np.random.default_rng(5)
    P = rng.standard_normal((1, 60, 3)).astype(np.float32)
    D = np.concatenate([P, P], axis=1)
    out = f3d.knn_graph(torch.from_numpy(D).cuda(), 5)["idx"].cpu().numpy()
    assert np.array_equal(out, oracle.knn_graph(D, 5))
    assert np.all(out[0, 60:, 0] == np.arange(60) + 60)
    Lt = rng.integers(0, 3, size=(2, 300, 3)).astype(np.float32)
    assert np.array_equal(f3d.knn_graph(torch.from_numpy(Lt).cuda(), 20)["idx"].cpu().numpy(), oracle.knn_graph(Lt, 20))
    assert np.array_equal(f3d.knn_graph(torch.from_numpy(Lt).cuda(), 20, flags=4)["idx"].cpu().numpy(), oracle.knn_graph(Lt, 20))
    assert np.array_equal(f3d.knn_graph(torch.from_numpy(Lt).cuda(), 20, flags=8)["idx"].cpu().numpy(), oracle.knn_graph(Lt, 20))


def test_knn_filter_adversarial(f3d, oracle):
    """Inputs that stress the TF32 filter's error bound: features far from the origin (cancellation in |x|²+|y|²-2x.y),
    wildly different scales, near-duplicates, high-dimensional clustered features."""
    rng = np.random.default_rng(11)
    base = rng.standard_normal((2, 700, 3)).astype(np.float32)
    for X in (base + np.float32(100.0),                       # offset 100: d~ loses ~4 digits
              base * np.float32(1e-4), base * np.float32(1e4),
              (base + rng.standard_normal(base.shape).astype(np.float32) * np.float32(1e-6)).astype(np.float32)):
        got = f3d.knn_graph(torch.from_numpy(np.ascontiguousarray(X)).cuda(), 20, want_dist=True, flags=8)
        idx, dist = oracle.knn_graph(X, 20, want_dist=True)
        assert np.array_equal(got["idx"].cpu().numpy(), idx) and np.array_equal(got["dist"].cpu().numpy(), dist)
    C = (rng.integers(0, 5, (2, 900, 1)) * 3.0 + rng.standard_normal((2, 900, 64)) * 0.05).astype(np.float32)
    got = f3d.knn_graph(torch.from_numpy(C).cuda(), 16)["idx"].cpu().numpy()
    assert np.array_equal(got, oracle.knn_graph(C, 16))


def test_create_single_knn_graph_shapes(f3d, oracle):
    """test/models.jl:24-41 asserts shapes only; here shape AND value."""
    X = np.random.default_rng(8).standard_normal((1024, 3)).astype(np.float32)
    g = f3d.create_single_knn_graph(torch.from_numpy(X).cuda(), 10)
    assert tuple(g.shape) == (1024, 10, 3)
    assert np.array_equal(g.cpu().numpy(), oracle.knn_graph(X[None], 10, want_gathered=True)[1][0])
    e = f3d.edgeconv_features(torch.from_numpy(X).cuda(), 10)
    assert tuple(e.shape) == (1, 1024, 10, 6)


def test_knn_properties_large(f3d):
    """Size-independent properties at N = 4096: distances ascending, self excluded, indices valid and unique,
    every returned neighbour closer than every non-returned point (checked against torch.cdist top-k distance)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn((4, 4096, 3), generator=g, device="cuda")
    out = f3d.knn_graph(X, 20, want_dist=True)
    idx, dist = out["idx"].long(), out["dist"]
    assert bool((dist[..., 1:] >= dist[..., :-1]).all())
    assert bool((idx != torch.arange(4096, device="cuda")[None, :, None]).all())
    assert int(idx.min()) >= 0 and int(idx.max()) < 4096
    srt = idx.sort(dim=-1).values
    assert bool((srt[..., 1:] != srt[..., :-1]).all())
    ref = torch.cdist(X.double(), X.double()).pow(2).topk(21, largest=False).values[..., 1:]
    assert torch.allclose(dist.double(), ref, rtol=1e-5, atol=1e-6)


def test_knn_tensor_path_is_a_real_filter(f3d):
    """The tensor-core path must actually filter: on random clouds (almost) no query may fall back to the exact scan of
    the whole cloud, and the number of exactly re-evaluated candidates per query must stay close to K+1."""
    for F, K in ((3, 20), (64, 20), (16, 10)):
        X = torch.randn((8, 1024, F), device="cuda")
        st = f3d.knn_graph(X, K, want_stats=True, flags=8)["stats"].cpu().numpy()
        queries = 8 * 1024
        assert st[0] == 0, (F, K, st)                          # nobody overflows to the exact scan on random clouds
        assert (K + 1) * queries <= st[1] <= 4 * (K + 1) * queries, (F, K, st)


@pytest.mark.parametrize("B,N,F,K", [(2, 300, 3, 20), (3, 257, 64, 10), (1, 100, 5, 7), (2, 1024, 16, 31)])
def test_edge_features_mlp_layout(f3d, oracle, B, N, F, K):
    """F3D_FLAG_EDGE_MLP_LAYOUT: the edge features arrive as (B, 2F, N, K) == the Julia (K*N, 2F, B) array that
    PermutedDimsArray(X, (2,3,1,4)) + reshape produce at src/models/dgcnn.jl:46-52 — bit-identical to the oracle's
    cat(X, KNNGraph - X) permuted that way."""
    X = np.random.default_rng(300 + F).standard_normal((B, N, F)).astype(np.float32)
    out = f3d.knn_graph(torch.from_numpy(X).cuda(), K, want_edge=True, mlp_layout=True)
    idx = oracle.knn_graph(X, K)
    assert np.array_equal(out["idx"].cpu().numpy(), idx)
    want = np.ascontiguousarray(oracle.edge_features(X, idx).transpose(0, 3, 1, 2))     # (B,N,K,2F) -> (B,2F,N,K)
    assert out["edge"].shape == (B, 2 * F, N, K)
    assert np.array_equal(out["edge"].cpu().numpy(), want)


def test_edge_features_are_differentiable_in_x(f3d):
    """Only CreateSingleKNNGraph is @nograd in the reference (dgcnn.jl:9); X flows through cat(X, KNNGraph - X) (dgcnn.jl:39-45):
    the gradient must equal that of the same expression built from torch gathers with the indices held constant."""
    X = torch.randn(2, 200, 6, device="cuda", dtype=torch.float32)
    K = 8
    for mlp_layout in (False, True):
        Xa = X.clone().requires_grad_(True)
        E = f3d.edgeconv_features(Xa, K, mlp_layout=mlp_layout)
        w = torch.randn_like(E)
        (E * w).sum().backward()
        idx = f3d.knn_graph(X, K)["idx"].long()

        def build(Xv):
            nb = torch.gather(Xv.unsqueeze(1).expand(-1, 200, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, 6))   # (B,N,K,F)
            out = torch.cat([Xv.unsqueeze(2).expand(-1, -1, K, -1), nb - Xv.unsqueeze(2)], dim=-1)
            return out.permute(0, 3, 1, 2) if mlp_layout else out
        assert torch.equal(E.detach(), build(X))          # float32 subtraction, rounded once — like the kernel
        Xb = X.clone().double().requires_grad_(True)
        ref = build(Xb)
        (ref * w.double()).sum().backward()
        assert torch.allclose(Xa.grad.double(), Xb.grad, rtol=1e-5, atol=1e-6)


def test_knn_tensor_path_zero_norm_queries(f3d, oracle):
    """Zero / tiny-norm queries (zero-padded points, dead ReLU features): the TF32 filter's window must not collapse to the
    selection threshold — near-tied neighbours of such a query have to survive into the exact re-evaluation."""
    rng = np.random.default_rng(77)
    N, F, K = 512, 32, 20
    X = rng.standard_normal((2, N, F)).astype(np.float32)
    X[:, :8] = 0.0                                     # all-zero rows (exact duplicates of each other)
    X[:, 8:16] *= np.float32(1e-6)                     # tiny norms
    # a shell of near-tied neighbours around the origin: |x| = 1 up to a few ulps
    shell = rng.standard_normal((2, 64, F)).astype(np.float32)
    shell /= np.linalg.norm(shell, axis=-1, keepdims=True)
    X[:, 16:80] = shell * (1.0 + rng.integers(-3, 4, (2, 64, 1)) * np.float32(2.0 ** -23))
    for flags in (0, f3d.FLAG_TENSOR):
        out = f3d.knn_graph(torch.from_numpy(X).cuda(), K, want_dist=True, flags=flags)
        assert np.array_equal(out["idx"].cpu().numpy(), oracle.knn_graph(X, K))


@pytest.mark.parametrize("B,N,F,K", [
    (3, 1024, 3, 20), (2, 1000, 1, 8), (2, 777, 2, 12), (2, 2048, 4, 31), (2, 1300, 3, 20),     # split-TF32 rows (F <= 4), fine / coarse chunks
    (2, 1024, 64, 20), (2, 900, 33, 16), (1, 2048, 16, 31), (2, 520, 7, 9), (2, 1500, 40, 25),  # plain rows, one and two 32-feature halves
])
def test_knn_gram_path(f3d, oracle, B, N, F, K):
    """knn_gram.cu (TMA-fed Gram filter, exact re-evaluation out of the operand tiles): parity with the oracle, and the
    diagnostics prove that this path served the call, that (almost) no row fell back to the exact scan and that the filter
    passed close to K + 1 candidates per query."""
    rng = np.random.default_rng(900 + N + F + K)
    X = rng.standard_normal((B, N, F)).astype(np.float32)
    out = f3d.knn_graph(torch.from_numpy(X).cuda(), K, want_dist=True, want_stats=True, flags=8)
    idx, dist = oracle.knn_graph(X, K, want_dist=True)
    assert np.array_equal(out["idx"].cpu().numpy(), idx)
    assert np.array_equal(out["dist"].cpu().numpy(), dist)
    st = out["stats"].cpu().numpy()
    assert st[7] == 2, st
    assert st[0] <= B * N // 100, st
    assert (K + 1) * B * N <= st[1] <= 3 * (K + 1) * B * N, st


def test_knn_gram_degenerate_inputs(f3d, oracle):
    """Heavy ties and clustered clouds overflow the candidate slots: those rows are redone by the exact scan at the end of their CTA;
    zero rows and huge offsets must not be certified wrongly."""
    rng = np.random.default_rng(77)
    lattice = rng.integers(0, 4, size=(2, 1100, 3)).astype(np.float32)                    # ~17 copies of every point
    clustered = (rng.integers(0, 3, (2, 1200, 1)) * 50.0 + rng.standard_normal((2, 1200, 64)) * 0.01).astype(np.float32)
    zeros = np.zeros((1, 1024, 3), np.float32)
    far = (rng.standard_normal((2, 1024, 3)) + 1000.0).astype(np.float32)
    for X, K in ((lattice, 20), (clustered, 16), (zeros, 5), (far, 20)):
        out = f3d.knn_graph(torch.from_numpy(X).cuda(), K, want_dist=True, want_stats=True, flags=8)
        idx, dist = oracle.knn_graph(X, K, want_dist=True)
        assert out["stats"].cpu().numpy()[7] == 2
        assert np.array_equal(out["idx"].cpu().numpy(), idx)
        assert np.array_equal(out["dist"].cpu().numpy(), dist)
