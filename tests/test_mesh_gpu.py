"""GPU parity of the TriMesh kernels (normals, areas, laplacian_loss, edge_loss) vs the oracle and the
reference's golden vectors, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from fixtures import (GOLD_FAREAS, GOLD_FNORMALS, GOLD_VNORMALS, MESH3_FACES, MESH3_VERTS, NORMALS_FACES, NORMALS_VERTS,
                      TEAPOT_LAPLACIAN_LOSS, pack, teapots)

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _meshes(oracle, golden_dir):
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vs, fs = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    return [("mesh3", MESH3_VERTS, MESH3_FACES), ("normals", NORMALS_VERTS, NORMALS_FACES), ("teapot", [vt], [ft]),
            ("teapot+sphere", [vt, vs], [ft, fs]), ("cfg4", *teapots(16, golden_dir, oracle))]


def test_reference_goldens(f3d):
    """test/rep.jl:178-388 — 4-decimal goldens for vertex normals, face normals and face areas."""
    m = f3d.TriMesh(NORMALS_VERTS, NORMALS_FACES)
    for mode in (f3d.NORMALS_REFERENCE_CPU, f3d.NORMALS_ACCUMULATE):
        vn = m.compute_verts_normals_packed(mode).cpu().numpy()
        assert np.allclose(vn, np.concatenate(GOLD_VNORMALS), rtol=1e-4, atol=1e-4)
        for got, gold in zip(m.compute_verts_normals_list(mode), GOLD_VNORMALS):
            assert np.allclose(got.cpu().numpy(), gold, rtol=1e-4, atol=1e-4)
    assert np.allclose(m.compute_faces_normals_packed().cpu().numpy(), np.concatenate(GOLD_FNORMALS), rtol=1e-4, atol=1e-4)
    assert np.allclose(m.compute_faces_areas_packed().cpu().numpy(), np.concatenate(GOLD_FAREAS), rtol=1e-4, atol=1e-4)
    pad = m.compute_verts_normals_padded().cpu().numpy()
    assert pad.shape == (2, 12, 3) and np.all(pad[1, 5:] == 0)   # rep.jl:303-304
    fpad = m.compute_faces_areas_padded().cpu().numpy()
    assert fpad.shape == (2, 4) and np.all(fpad[1, 2:] == 0)


def test_normals_areas_bit_exact(f3d, oracle, golden_dir):
    for name, vl, fl in _meshes(oracle, golden_dir):
        m = f3d.TriMesh(vl, fl)
        v, f = pack(vl, fl)
        areas, fn = oracle.faces_areas_normals(v, f)
        assert np.array_equal(m.compute_faces_areas_packed().cpu().numpy(), areas), name
        assert np.array_equal(m.compute_faces_normals_packed().cpu().numpy(), fn), name
        for mode in (0, 1):
            got = m.compute_verts_normals_packed(mode).cpu().numpy()
            assert np.array_equal(got, oracle.verts_normals(v, f, mode)), (name, mode)


def test_laplacian_loss(f3d, oracle, golden_dir):
    for name, vl, fl in _meshes(oracle, golden_dir):
        m = f3d.TriMesh(vl, fl)
        v, f = pack(vl, fl)
        got = float(f3d.laplacian_loss(m).item())
        ref = float(oracle.laplacian_loss(v, f))
        assert abs(got - ref) <= RTOL * abs(ref), (name, got, ref)
    m = f3d.load_trimesh(os.path.join(golden_dir, "teapot.obj"))
    got = float(f3d.laplacian_loss(m).item())
    assert abs(got - float(TEAPOT_LAPLACIAN_LOSS)) <= 1e-6 * got  # README.md:111-112
    # bitwise repeatable (no atomics in the reduction)
    assert f3d.laplacian_loss(m).item() == got


def test_edge_loss(f3d, oracle, golden_dir):
    for name, vl, fl in _meshes(oracle, golden_dir):
        m = f3d.TriMesh(vl, fl)
        v, f = pack(vl, fl)
        for target in (0.0, 0.05):
            got = float(f3d.edge_loss(m, target).item())
            ref = float(oracle.edge_loss(v, f, target))
            assert abs(got - ref) <= RTOL * abs(ref) + 1e-12, (name, target, got, ref)


def test_laplacian_backward(f3d, oracle, golden_dir):
    """Pullback vs a float64 torch autograd restatement of the same dense expression (the reference's own
    gradient test only asserts `isa Tuple`, test/metrics.jl:72)."""
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    m = f3d.TriMesh([vt], [ft])
    verts = m.get_verts_packed().clone().requires_grad_(True)
    m2 = f3d.TriMesh._from_packed(m, verts)
    (f3d.laplacian_loss(m2) * 3.0).backward()
    rowptr, colidx, vals = m.get_laplacian_packed()
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    Ld = torch.zeros((vt.shape[0], vt.shape[0]), dtype=torch.float64)
    Ld[torch.from_numpy(rows), torch.from_numpy(colidx.astype(np.int64))] = torch.from_numpy(vals.astype(np.float64))
    x = torch.from_numpy(vt.astype(np.float64)).requires_grad_(True)
    ((Ld @ x).norm(dim=1).mean() * 3.0).backward()
    assert torch.allclose(verts.grad.cpu().double(), x.grad, rtol=1e-4, atol=1e-7)


def test_laplacian_sharded_sum(f3d, oracle, golden_dir):
    """Mesh-axis sharding: shard losses with the global vertex count add up to the batch loss."""
    vl, fl = teapots(4, golden_dir, oracle)
    full = float(f3d.laplacian_loss(f3d.TriMesh(vl, fl)).item())
    tot = sum(v.shape[0] for v in vl)
    parts = [float(f3d.laplacian_loss(f3d.TriMesh(vl[a:b], fl[a:b]), verts_total=tot).item()) for a, b in ((0, 2), (2, 4))]
    assert abs(sum(parts) - full) <= 1e-6 * full


def test_edge_loss_backward(f3d, oracle, golden_dir):
    """edge_loss pullback vs float64 torch autograd of the same expression (reference: `gradient(...) isa Tuple`)."""
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    m = f3d.TriMesh([vt], [ft])
    for target in (0.0, 0.05):
        verts = m.get_verts_packed().clone().requires_grad_(True)
        (f3d.edge_loss(f3d.TriMesh._from_packed(m, verts), target) * 2.5).backward()
        e = torch.from_numpy(m.get_edges_packed().astype(np.int64))
        x = torch.from_numpy(vt.astype(np.float64)).requires_grad_(True)
        ((((x[e[:, 0]] - x[e[:, 1]]).norm(dim=1) - target) ** 2).mean() * 2.5).backward()
        assert torch.allclose(verts.grad.cpu().double(), x.grad, rtol=1e-4, atol=1e-8)


def test_fit_mesh_step(f3d, oracle, golden_dir):
    """The reference's training objective (examples/fit_mesh.jl:78-84): chamfer(sample(src + offset), sample(tgt)) +
    0.1 laplacian + edge loss, differentiated w.r.t. the offsets — a few SGD steps on the device must lower it."""
    vs, fs = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vt = (vt - vt.mean(0)) / np.abs(vt - vt.mean(0)).max()
    src, tgt = f3d.TriMesh([vs], [fs]), f3d.TriMesh([vt.astype(np.float32)], [ft])
    offs = torch.zeros_like(src.get_verts_packed(), requires_grad=True)
    opt = torch.optim.SGD([offs], lr=1.0, momentum=0.9)
    losses = []
    for it in range(12):
        opt.zero_grad()
        m = f3d.offset(src, offs)
        a = f3d.sample_points(m, 5000, seed=100 + it)
        b = f3d.sample_points(tgt, 5000, seed=200 + it)
        loss = f3d.chamfer_distance(a, b) + 0.1 * f3d.laplacian_loss(m) + f3d.edge_loss(m)
        loss.backward()
        assert torch.isfinite(offs.grad).all() and float(offs.grad.abs().max()) > 0
        opt.step()
        losses.append(float(loss.item()))
    assert losses[-1] < 0.8 * losses[0], losses


def test_packed_padded_converters(f3d, oracle):
    """_packed_to_padded / _padded_to_packed on the device (src/rep/utils.jl:131-185) against the reference's own golden
    (test/rep.jl:403-490: values and the gradient identities) and against the numpy restatement on ragged random items
    (an empty item included)."""
    packed = np.arange(1, 28, dtype=np.float32).reshape(9, 3)
    items_len = [4, 2, 3]
    tp = torch.from_numpy(packed).cuda().requires_grad_(True)
    padded = f3d.packed_to_padded(tp, items_len, 0)
    want = oracle.np_packed_to_padded(packed, items_len, 0)
    assert np.array_equal(padded.detach().cpu().numpy(), want)
    (0.5 * (padded ** 2).sum()).backward()                      # rep.jl:455-458: the gradient is the packed array itself
    assert np.array_equal(tp.grad.cpu().numpy(), packed)
    tq = torch.from_numpy(want).cuda().requires_grad_(True)
    back = f3d.padded_to_packed(tq, items_len)
    assert np.array_equal(back.detach().cpu().numpy(), packed)   # rep.jl:482
    (0.5 * (back ** 2).sum()).backward()                        # rep.jl:484-488: the gradient is the padded array (pads 0)
    assert np.array_equal(tq.grad.cpu().numpy(), want)
    assert np.array_equal(f3d.packed_to_padded(tp.detach(), items_len, -7.5).cpu().numpy(), oracle.np_packed_to_padded(packed, items_len, -7.5))
    rng = np.random.default_rng(5)
    lens = [0, 1, 37, 256, 5, 1000]
    big = rng.standard_normal((sum(lens), 3)).astype(np.float32)
    got = f3d.packed_to_padded(torch.from_numpy(big).cuda(), lens, 0)
    assert np.array_equal(got.cpu().numpy(), oracle.np_packed_to_padded(big, lens, 0))
    assert np.array_equal(f3d.padded_to_packed(got, lens).cpu().numpy(), big)


def test_trimesh_padded_views_on_device(f3d, oracle, golden_dir):
    """TriMesh's padded views come from the device converters: padded verts / normals / areas and padded faces (local ids,
    pads -1) equal the host restatement for a heterogeneous batch."""
    import os
    v1, f1 = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    v2, f2 = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    m = f3d.TriMesh([v1, v2, v1[:200]], [f1, f2, f1[(f1 < 200).all(1)]])
    lens_v = [len(v1), len(v2), 200]
    vp = m.get_verts_padded().cpu().numpy()
    assert np.array_equal(vp, oracle.np_packed_to_padded(m.get_verts_packed().cpu().numpy(), lens_v, 0))
    assert np.array_equal(m.faces_padded_device().cpu().numpy(), m.get_faces_padded())
    fa = m.compute_faces_areas_padded().cpu().numpy()
    lens_f = [len(f) for f in m.get_faces_list()]
    assert np.array_equal(fa, oracle.np_packed_to_padded(m.compute_faces_areas_packed().cpu().numpy(), lens_f, 0))
    vn = m.compute_verts_normals_padded().cpu().numpy()
    assert np.array_equal(vn, oracle.np_packed_to_padded(m.compute_verts_normals_packed().cpu().numpy(), lens_v, 0))
