"""Pins the CPU oracle (oracle/) against every golden vector / known answer / identity the reference's own
tests hold for the hot path, and cross-checks the C oracle against its independent numpy twin.
CPU only."""
import os

import numpy as np
import pytest

from fixtures import (GOLD_FAREAS, GOLD_FNORMALS, GOLD_VNORMALS, MESH3_FACES, MESH3_VERTS, NORMALS_FACES, NORMALS_VERTS,
                      TEAPOT_LAPLACIAN_LOSS, pack, pad)


def test_teapot_laplacian_known_answer(oracle, golden_dir):
    """README.md:111-112: laplacian_loss(load_trimesh("teapot.obj")) = 0.05888283f0 — every printed digit."""
    v, f = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    assert v.shape == (1202, 3) and f.shape == (2256, 3)
    assert oracle.laplacian_loss(v, f) == TEAPOT_LAPLACIAN_LOSS
    assert abs(float(oracle.np_laplacian_loss(v, f, np.float64)) - 0.0588828611) < 1e-9


def test_sphere_fixture(oracle, golden_dir):
    v, f = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    assert v.shape == (2562, 3) and f.shape == (5120, 3)
    edges, _ = oracle.edges_packed(f, v.shape[0])
    assert edges.shape == (7680, 2)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-6)
    assert abs(float(oracle.laplacian_loss(v, f)) - 0.004000934) < 1e-8


def test_normals_areas_goldens(oracle):
    """test/rep.jl:178-388: vertex normals, face normals, face areas to 1e-4 — both normals modes."""
    for verts, faces, gv, gf, ga in zip(NORMALS_VERTS, NORMALS_FACES, GOLD_VNORMALS, GOLD_FNORMALS, GOLD_FAREAS):
        areas, fn = oracle.faces_areas_normals(verts, faces)
        assert np.allclose(areas, ga, rtol=1e-4, atol=1e-4)
        assert np.allclose(fn, gf, rtol=1e-4, atol=1e-4)
        for mode in (0, 1):
            assert np.allclose(oracle.verts_normals(verts, faces, mode), gv, rtol=1e-4, atol=1e-4)
    # packed over the batch of two (rep.jl:270-278)
    v, f = pack(NORMALS_VERTS, NORMALS_FACES)
    assert np.allclose(oracle.verts_normals(v, f, 0), np.concatenate(GOLD_VNORMALS), rtol=1e-4, atol=1e-4)


def test_edges_faces_to_edges(oracle):
    """test/rep.jl:136-156: edges == sort/unique restatement; faces_to_edges columns (e23, e31, e12)."""
    v, f = pack(MESH3_VERTS, MESH3_FACES)
    edges, f2e = oracle.edges_packed(f, v.shape[0])
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    e = np.unique(np.sort(e, axis=1), axis=0)
    assert np.array_equal(edges, e)
    for i in range(f.shape[0]):
        assert np.array_equal(edges[f2e[i, 0]], np.sort(f[i, [1, 2]]))
        assert np.array_equal(edges[f2e[i, 1]], np.sort(f[i, [0, 2]]))
        assert np.array_equal(edges[f2e[i, 2]], np.sort(f[i, [0, 1]]))


def _dense_laplacian(edges, V):
    L = np.zeros((V, V))
    for a, b in edges:
        L[a, b] = 1
        L[b, a] = 1
    deg = L.sum(1)
    inv = np.where(deg > 0, 1 / np.maximum(deg, 1), deg)
    for i in range(V):
        for j in range(V):
            if i == j:
                L[i, j] = -1
            elif L[i, j] == 1:
                L[i, j] = inv[i]
    return L


def test_laplacian_matrix_and_loss_identity(oracle):
    """test/rep.jl:158-175 (entries to 1e-5) and test/metrics.jl:8-73 (loss == dense restatement)."""
    v, f = pack(MESH3_VERTS, MESH3_FACES)
    V = v.shape[0]
    edges, _ = oracle.edges_packed(f, V)
    rowptr, colidx, vals = oracle.laplacian_csr(edges, V)
    dense = np.zeros((V, V), np.float32)
    for i in range(V):
        cols = colidx[rowptr[i]:rowptr[i + 1]]
        assert np.all(np.diff(cols) > 0)
        dense[i, cols] = vals[rowptr[i]:rowptr[i + 1]]
    L = _dense_laplacian(edges, V)
    assert np.allclose(dense, L, rtol=1e-5, atol=1e-5)
    expect = np.mean(np.sqrt(((L @ v.astype(np.float64)) ** 2).sum(1)))
    got = float(oracle.laplacian_loss(v, f))
    assert abs(got - expect) <= 3.4526698e-4 * abs(expect)  # default isapprox, rtol = sqrt(eps(Float32))
    assert abs(got - expect) <= 1e-6 * abs(expect)


def test_edge_loss_identity(oracle, golden_dir):
    """test/metrics.jl:75-84: edge_loss(m) == mean(norm(v1-v2)^2) over unique edges, teapot+sphere batch."""
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vs, fs = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    v, f = pack([vt, vs], [ft, fs])
    edges, _ = oracle.edges_packed(f, v.shape[0])
    assert edges.shape[0] == 3456 + 7680
    d = v[edges[:, 0]].astype(np.float64) - v[edges[:, 1]].astype(np.float64)
    expect = np.mean((d ** 2).sum(1))
    assert abs(float(oracle.edge_loss(v, f)) - expect) <= 1e-6 * expect


def test_chamfer_naive_identity(oracle):
    """test/metrics.jl:94-111: chamfer_distance ≈ naive_chamfer on rand(3,1000,2) vs rand(3,500,2)."""
    x = np.random.default_rng(51).random((2, 1000, 3), dtype=np.float32)
    y = np.random.default_rng(52).random((2, 500, 3), dtype=np.float32)
    got = float(oracle.chamfer_distance(x, y))
    assert abs(got - oracle.np_naive_chamfer(x, y)) <= 3.4526698e-4 * got
    assert float(oracle.chamfer_distance(x, x)) == 0.0  # test/metrics.jl:88-89


def test_chamfer_c_vs_numpy_twin(oracle):
    for (B, N, M, seed) in [(2, 1024, 1024, 101), (2, 300, 77, 5), (1, 1, 5, 6)]:
        A = np.random.default_rng(seed).random((B, N, 3), dtype=np.float32)
        Bc = np.random.default_rng(seed + 1).random((B, M, 3), dtype=np.float32)
        loss, nA, nB, _ = oracle.chamfer_distance(A, Bc, return_all=True)
        l64, tA, tB = oracle.np_chamfer_distance(A, Bc)
        assert np.array_equal(nA, tA) and np.array_equal(nB, tB)
        assert abs(float(loss) - l64) <= 1e-6 * l64
    # cfg1 known value (seeds 101/102): matches the KD-tree probe of the survey
    A = np.random.default_rng(101).random((2, 1024, 3), dtype=np.float32)
    Bc = np.random.default_rng(102).random((2, 1024, 3), dtype=np.float32)
    assert abs(float(oracle.chamfer_distance(A, Bc)) - 0.0074543296) <= 1e-6 * 0.0074543296
    assert abs(float(oracle.kdtree_chamfer(A, Bc)) - 0.0074543296) <= 1e-5 * 0.0074543296


def test_chamfer_backward_vs_finite_difference(oracle):
    """Pullback restatement vs central differences of the Float64 twin (reference bar: atol 1e-2, rtol 1e-3,
    test/metrics.jl:112-114)."""
    rng = np.random.default_rng(9)
    A = rng.random((1, 40, 3), dtype=np.float32)
    Bc = rng.random((1, 30, 3), dtype=np.float32)
    _, nA, nB, _ = oracle.chamfer_distance(A, Bc, 0.7, 1.3, return_all=True)
    gA, gB = oracle.chamfer_backward(A, Bc, nA, nB, 0.7, 1.3)
    h = 1e-3
    for (arr, g) in ((A, gA), (Bc, gB)):
        for idx in [(0, 3, 1), (0, 17, 0), (0, 29, 2)]:
            p = arr.copy(); p[idx] += h
            m = arr.copy(); m[idx] -= h
            if arr is A:
                fd = (oracle.np_chamfer_distance(p, Bc, 0.7, 1.3)[0] - oracle.np_chamfer_distance(m, Bc, 0.7, 1.3)[0]) / (2 * h)
            else:
                fd = (oracle.np_chamfer_distance(A, p, 0.7, 1.3)[0] - oracle.np_chamfer_distance(A, m, 0.7, 1.3)[0]) / (2 * h)
            assert abs(fd - g[idx]) <= 1e-2 + 1e-3 * abs(fd)


def test_knn_c_vs_numpy_twin(oracle):
    rng = np.random.default_rng(301)
    X = rng.standard_normal((2, 200, 3)).astype(np.float32)
    assert np.array_equal(oracle.knn_graph(X, 10), oracle.np_knn_graph(X, 10))
    X64 = rng.standard_normal((1, 96, 64)).astype(np.float32)
    assert np.array_equal(oracle.knn_graph(X64, 20), oracle.np_knn_graph(X64, 20))
    # exact duplicates: the lower-indexed twin is the one dropped by position (dgcnn.jl:6)
    D = np.concatenate([X[:, :50], X[:, :50]], axis=1)
    idx = oracle.knn_graph(D, 3)
    assert np.all(idx[0, :50, 0] == np.arange(50) + 50)   # point i<50: rank0 = i (itself), rank1 = i+50
    assert np.all(idx[0, 50:, 0] == np.arange(50) + 50)   # point i+50: rank0 = i (the twin!), rank1 = itself
    g = oracle.knn_graph(X, 4, want_gathered=True)[1]
    ii = oracle.knn_graph(X, 4)
    assert np.array_equal(g, np.take_along_axis(X[:, None], ii[..., None].astype(np.int64), axis=2))
    e = oracle.edge_features(X, ii)
    assert np.array_equal(e[..., :3], np.broadcast_to(X[:, :, None, :], e[..., :3].shape))
    assert np.array_equal(e[..., 3:], g - X[:, :, None, :])


def test_verts_normals_modes_c_vs_numpy(oracle, golden_dir):
    v, f = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    for mode in (0, 1):
        assert np.array_equal(oracle.verts_normals(v, f, mode), oracle.np_verts_normals(v, f, mode))
    a, n = oracle.faces_areas_normals(v, f)
    a2, n2 = oracle.np_faces_areas_normals(v, f)
    assert np.array_equal(a, a2) and np.array_equal(n, n2)
    # the two modes really differ on the teapot
    d = (oracle.verts_normals(v, f, 0) * oracle.verts_normals(v, f, 1)).sum(1)
    assert np.degrees(np.arccos(np.clip(d, -1, 1))).max() > 10.0


def test_sample_points_statistics_and_injection(oracle, golden_dir):
    """test/transforms/mesh_func.jl:4-14: samples of the unit sphere have radius ≈ 1 (rtol 1e-2); plus the
    barycentric formula with injected draws and area-proportional face frequencies."""
    v, f = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    vp, fp, vl, fl = pad([v], [f])
    pts, fidx = oracle.sample_points(vp, fp, vl, fl, 1000, seed=401)
    assert np.allclose(np.linalg.norm(pts[0], axis=1), 1.0, rtol=1e-2)
    # injected draws: p = ((w1 v1) + (w2 v2)) + (w3 v3), w from sqrt(r1), r2   (mesh_func.jl:60-82)
    rng = np.random.default_rng(402)
    S = 64
    jf = rng.integers(0, f.shape[0], (1, S)).astype(np.int32)
    r1 = rng.random((1, S), dtype=np.float32)
    r2 = rng.random((1, S), dtype=np.float32)
    pts, fidx = oracle.sample_points(vp, fp, vl, fl, S, inj_face=jf, inj_r1=r1, inj_r2=r2)
    u = np.sqrt(r1[0]); w1 = np.float32(1) - u; w2 = u * (np.float32(1) - r2[0]); w3 = u * r2[0]
    tri = v[f[jf[0]]]
    expect = (w1[:, None] * tri[:, 0] + w2[:, None] * tri[:, 1]) + w3[:, None] * tri[:, 2]
    assert np.array_equal(pts[0], expect.astype(np.float32)) and np.array_equal(fidx, jf)
    # face frequencies follow the areas (teapot: uneven areas)
    v, f = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vp, fp, vl, fl = pad([v], [f])
    _, fidx = oracle.sample_points(vp, fp, vl, fl, 200000, seed=7)
    areas, _ = oracle.faces_areas_normals(v, f)
    freq = np.bincount(fidx[0], minlength=f.shape[0]) / 200000
    p = areas / areas.sum()
    big = p > 2e-3
    assert np.allclose(freq[big], p[big], rtol=0.15)
    assert abs(freq.sum() - 1) < 1e-12 and fidx.max() < f.shape[0]


def test_philox_known_answer(oracle):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): counter/key all zero and all ones."""
    import ctypes as C
    L = oracle.lib()
    # all-zero counter & key → 6627e8d5 e169c58d bc57ac4c 9b00dbd8 ; draws() packs (c0,c1)>>11, c2>>8, c3>>8
    u, r1, r2 = oracle.philox_draws(0, 0, 0, 0)
    assert u == ((0x6627e8d5 << 32) | 0xe169c58d) >> 11
    assert r1 == np.float32((0xbc57ac4c >> 8) / 16777216.0) and r2 == np.float32((0x9b00dbd8 >> 8) / 16777216.0)
    u, r1, r2 = oracle.philox_draws(0xffffffffffffffff, 0xffffffffffffffff, -1, -1)
    assert u == ((0x408f276d << 32) | 0x41c83b0e) >> 11
    assert r1 == np.float32((0xa20bc7c6 >> 8) / 16777216.0) and r2 == np.float32((0x6d5451fd >> 8) / 16777216.0)


def test_packed_padded_converters_reference_golden(oracle):
    """test/rep.jl:403-490: the 9 points 1..27 split 4 / 2 / 3 — _packed_to_padded, _list_to_padded, _padded_to_packed."""
    packed = np.arange(1, 28, dtype=np.float32).reshape(9, 3)   # == the Julia (3, 9) matrix read column by column
    items_len = [4, 2, 3]
    padded = np.zeros((3, 4, 3), np.float32)
    padded[0, :4], padded[1, :2], padded[2, :3] = packed[0:4], packed[4:6], packed[6:9]
    assert np.array_equal(oracle.np_packed_to_padded(packed, items_len, 0), padded)
    assert np.array_equal(oracle.np_list_to_padded([packed[0:4], packed[4:6], packed[6:9]], 0), padded)
    assert np.array_equal(oracle.np_padded_to_packed(padded, items_len), packed)
