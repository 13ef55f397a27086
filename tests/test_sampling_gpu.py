"""GPU parity of sample_points vs the oracle (seeded Philox draws: bit-exact face ids and points; injected
draws: bit-exact with the reference's barycentric arithmetic) plus the reference's own statistical test."""
import os

import numpy as np
import pytest
import torch

from fixtures import MESH3_FACES, MESH3_VERTS, pad, teapots

pytestmark = pytest.mark.gpu


def test_cfg4_seeded_bit_exact(f3d, oracle, golden_dir):
    """BASELINE configs[3] shape: 16 teapots, 10 000 samples per mesh, Philox seed 401."""
    vl, fl = teapots(16, golden_dir, oracle)
    m = f3d.TriMesh(vl, fl)
    pts, fidx = f3d.sample_points(m, 10000, seed=401, return_faces=True)
    vp, fp, vlen, flen = pad(vl, fl)
    opts, ofidx = oracle.sample_points(vp, fp, vlen, flen, 10000, seed=401)
    assert np.array_equal(fidx.cpu().numpy(), ofidx)
    assert np.array_equal(pts.cpu().numpy(), opts)
    # a different offset gives different draws, the same (seed, offset) the same ones
    p2 = f3d.sample_points(m, 10000, seed=401, offset=1)
    assert not torch.equal(p2, pts)
    assert torch.equal(f3d.sample_points(m, 10000, seed=401), pts)


def test_heterogeneous_batch_and_injected(f3d, oracle, golden_dir):
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vs, fs = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    vl, fl = MESH3_VERTS + [vt, vs], MESH3_FACES + [ft, fs]
    m = f3d.TriMesh(vl, fl)
    assert not m.equalised
    vp, fp, vlen, flen = pad(vl, fl)
    pts, fidx = f3d.sample_points(m, 777, seed=12345, return_faces=True)
    opts, ofidx = oracle.sample_points(vp, fp, vlen, flen, 777, seed=12345)
    assert np.array_equal(fidx.cpu().numpy(), ofidx) and np.array_equal(pts.cpu().numpy(), opts)
    assert all(int(fidx[i].max()) < flen[i] for i in range(m.N))
    # injected draws (bit-parity mode, rng(402) as in SURVEY §8d)
    rng = np.random.default_rng(402)
    S = 500
    jf = np.stack([rng.integers(0, flen[i], S) for i in range(m.N)]).astype(np.int32)
    r1 = rng.random((m.N, S), dtype=np.float32)
    r2 = rng.random((m.N, S), dtype=np.float32)
    pts, fidx = f3d.sample_points(m, S, inj_face=torch.from_numpy(jf), inj_r1=torch.from_numpy(r1),
                                  inj_r2=torch.from_numpy(r2), return_faces=True)
    opts, _ = oracle.sample_points(vp, fp, vlen, flen, S, inj_face=jf, inj_r1=r1, inj_r2=r2)
    assert np.array_equal(pts.cpu().numpy(), opts) and np.array_equal(fidx.cpu().numpy(), jf)


def test_sphere_radius(f3d, golden_dir):
    """test/transforms/mesh_func.jl:4-14 / test/cuda/metrics.jl:87-99: samples of the unit sphere have radius ≈ 1."""
    m = f3d.load_trimesh(os.path.join(golden_dir, "sphere.obj"))
    pts = f3d.sample_points(m, 1000)
    assert tuple(pts.shape) == (1, 1000, 3)
    assert torch.allclose(pts.norm(dim=-1), torch.ones(1, 1000, device="cuda"), rtol=1e-2)


def test_area_proportional(f3d, oracle, golden_dir):
    m = f3d.load_trimesh(os.path.join(golden_dir, "teapot.obj"))
    _, fidx = f3d.sample_points(m, 400000, seed=9, return_faces=True)
    areas = m.compute_faces_areas_packed().double()
    freq = torch.bincount(fidx[0].long(), minlength=areas.numel()).double() / 400000
    p = areas / areas.sum()
    big = p > 2e-3
    assert torch.allclose(freq[big], p[big], rtol=0.1)


def test_degenerate_mesh(f3d):
    """Zero total area: probabilities are 0/eps = 0 and the residual goes to the last face (mesh_func.jl:36-37)."""
    v = np.zeros((4, 3), np.float32)
    f = np.array([[0, 1, 2], [1, 2, 3]], np.int32)
    pts, fidx = f3d.sample_points(f3d.TriMesh([v], [f]), 64, seed=1, return_faces=True)
    assert bool((fidx == 1).all()) and bool((pts == 0).all())


def test_mesh_chamfer(f3d, golden_dir):
    """chamfer_distance(m, m) ≈ 0 with atol 1e-2 (test/metrics.jl:86-89): two independent samplings of one surface."""
    m = f3d.load_trimesh([os.path.join(golden_dir, "teapot.obj"), os.path.join(golden_dir, "sphere.obj")])
    loss = f3d.chamfer_distance(m, m)
    assert 0.0 <= float(loss.item()) <= 1e-2


def test_sample_points_backward(f3d, oracle, golden_dir):
    """Pullback of the barycentric combination: gverts[face corner k] += w_k * gsample.  Injected draws make the
    weights known analytically; the float64 restatement goes through torch autograd (atomics: rtol 1e-4)."""
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vs, fs = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    m0 = f3d.TriMesh([vt, vs], [ft, fs])
    verts = m0.get_verts_packed().clone().requires_grad_(True)
    m = f3d.TriMesh._from_packed(m0, verts)
    rng = np.random.default_rng(8)
    S = 3000
    jf = np.stack([rng.integers(0, n, S) for n in (ft.shape[0], fs.shape[0])]).astype(np.int32)
    r1 = rng.random((2, S), dtype=np.float32)
    r2 = rng.random((2, S), dtype=np.float32)
    pts = f3d.sample_points(m, S, inj_face=torch.from_numpy(jf), inj_r1=torch.from_numpy(r1), inj_r2=torch.from_numpy(r2))
    w = torch.randn_like(pts)
    (pts * w).sum().backward()
    x = m0.get_verts_padded().double().clone().requires_grad_(True)
    fp = torch.from_numpy(m0.get_faces_padded().astype(np.int64)).cuda()
    u = torch.from_numpy(np.sqrt(r1.astype(np.float64))).cuda()
    v = torch.from_numpy(r2.astype(np.float64)).cuda()
    wts = torch.stack([1 - u, u * (1 - v), u * v], dim=-1)                                  # (N, S, 3)
    tri = torch.stack([x[i][fp[i][torch.from_numpy(jf[i].astype(np.int64)).cuda()]] for i in range(2)])  # (N, S, 3, 3)
    ref_pts = (wts.unsqueeze(-1) * tri).sum(2)
    assert torch.allclose(pts.double(), ref_pts, rtol=1e-5, atol=1e-6)
    (ref_pts * w.double()).sum().backward()
    g_ref = torch.cat([x.grad[0, :vt.shape[0]], x.grad[1, :vs.shape[0]]])
    assert torch.allclose(verts.grad.double(), g_ref, rtol=1e-4, atol=1e-5)
