"""The drop-in boundary on a machine WITHOUT a GPU: the C-ABI library loads and exports every symbol the
header declares, the host-only entry points work, argument errors come back as status codes with a message
(no exceptions across the ABI, no CUDA call needed), and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from fixtures import MESH3_FACES, MESH3_VERTS, pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flux3d_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"F3D_API\s+[\w\s\*]+?\b(f3d_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound(f3d):
    declared = _declared_symbols()
    assert len(declared) >= 23
    out = subprocess.check_output(["nm", "-D", "--defined-only", f3d.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (f3d_\w+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert exported <= set(declared), f"exports not declared in the header: {sorted(exported - set(declared))}"
    assert set(f3d._lib.SIGNATURES) == set(declared)  # the ctypes table mirrors the header one to one
    assert f3d._lib.lib().f3d_version() == 100


def test_library_is_sm100a_only(f3d):
    out = subprocess.check_output(["cuobjdump", "-lelf", f3d.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flux3d.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{fn} mentions the oracle"


def test_missing_library_fails_loudly(tmp_path):
    code = ("import os, sys; os.environ['FLUX3D_B200_LIB']=r'%s'; sys.path.insert(0, r'%s'); import flux3d_b200 as f\n"
            "try:\n    f._lib.lib()\nexcept f.Flux3DB200Error as e:\n    print('RAISED', e)\n") % (tmp_path / "nope.so", ROOT)
    out = subprocess.check_output(["python", "-c", code], text=True)
    assert "RAISED" in out and "no CPU or PyTorch fallback" in out


def test_argument_errors_are_status_codes(f3d):
    L = f3d._lib.lib()
    assert L.f3d_chamfer_fwd(None, None, 1, 1, 1, 1.0, 1.0, 0, None, None, None, None, None, 0, 0, None) == 1
    assert "null" in f3d._lib.last_error()
    dummy = C.c_void_p(256)  # never dereferenced: argument validation comes first
    assert L.f3d_chamfer_fwd(dummy, dummy, 0, 4, 4, 1.0, 1.0, 0, dummy, None, None, None, None, 0, 0, None) == 1
    assert "positive" in f3d._lib.last_error()
    assert L.f3d_chamfer_fwd(dummy, dummy, 2, 4, 4, 1.0, 1.0, 1, dummy, None, None, None, dummy, 1 << 20, 0, None) == 1  # B_total < B
    assert L.f3d_chamfer_fwd(dummy, dummy, 2, 4, 4, 1.0, 1.0, 0, dummy, None, None, None, None, 0, 0, None) == 3   # workspace
    assert "workspace" in f3d._lib.last_error()
    assert L.f3d_knn_graph(dummy, 1, 10, 3, 10, dummy, None, None, None, None, 0, 0, None) == 1  # K >= N
    assert L.f3d_knn_graph(dummy, 1, 100, 3, 64, dummy, None, None, None, None, 0, 0, None) == 1  # K > 63
    assert L.f3d_verts_normals(dummy, dummy, dummy, dummy, 4, 4, 7, dummy, None) == 1  # unknown mode
    assert L.f3d_sample_points(dummy, dummy, None, dummy, 1, 4, 4, 0, 1e-6, 0, 0, None, None, None, dummy, None, None, None, 0, None) == 1
    assert L.f3d_laplacian_loss(dummy, dummy, dummy, dummy, 4, 0, dummy, None, 0, None) == 3
    with pytest.raises(f3d.Flux3DB200Error):
        f3d._lib.check(1)
    assert L.f3d_chamfer_workspace_bytes(32, 4096, 4096) > 0 and L.f3d_chamfer_workspace_bytes(0, 1, 1) == 0


def test_host_wrappers_reject_bad_input(f3d):
    import torch
    with pytest.raises(f3d.Flux3DB200Error):  # CPU tensors: there is no CPU path
        f3d.chamfer_forward_raw(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3), 1.0, 1.0)
    with pytest.raises(ValueError):  # rep/mesh.jl:126-128
        f3d.TriMesh([np.zeros((3, 3), np.float32)], [], device="cpu")
    with pytest.raises(ValueError):
        f3d.TriMesh([np.zeros((3, 3), np.float32)], [np.array([[0, 1, 3]])], device="cpu")


def test_topology_build_matches_oracle(f3d, oracle, golden_dir):
    """f3d_mesh_topology_build_host is host code: edges / faces_to_edges / Laplacian CSR / vertex→corner CSR
    against the oracle's restatement of rep/mesh.jl:907-1002, on the reference's 3-mesh fixture and teapot+sphere."""
    vt, ft = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vs, fs = oracle.load_obj(os.path.join(golden_dir, "sphere.obj"))
    for vl, fl in ((MESH3_VERTS, MESH3_FACES), ([vt, vs], [ft, fs])):
        m = f3d.TriMesh(vl, fl, device="cpu")
        v, f = pack(vl, fl)
        assert np.array_equal(m.get_faces_packed(), f)
        edges, f2e = oracle.edges_packed(f, v.shape[0])
        assert np.array_equal(m.get_edges_packed(), edges)
        assert np.array_equal(m.get_faces_to_edges_packed(), f2e)
        rowptr, colidx, vals = oracle.laplacian_csr(edges, v.shape[0])
        r2, c2, v2 = m.get_laplacian_packed()
        assert np.array_equal(r2, rowptr) and np.array_equal(c2, colidx) and np.array_equal(v2, vals)
        t = m._topology()
        for vert in range(v.shape[0]):
            corners = t["v2c"][t["v2c_rowptr"][vert]:t["v2c_rowptr"][vert + 1]]
            expect = sorted((int(c) for c in np.flatnonzero(f.reshape(-1) == vert)), key=lambda c: (c % 3, c // 3))
            assert list(corners) == expect
    # padded / list getters (test/rep.jl:112-134)
    m = f3d.TriMesh(MESH3_VERTS, MESH3_FACES, device="cpu")
    vp = m.get_verts_padded().numpy()
    fp = m.get_faces_padded()
    for i, (v, f) in enumerate(zip(MESH3_VERTS, MESH3_FACES)):
        assert np.array_equal(vp[i, :len(v)], v) and np.all(vp[i, len(v):] == 0)
        assert np.array_equal(fp[i, :len(f)], f) and np.all(fp[i, len(f):] == -1)
        assert np.array_equal(m.get_verts_list()[i].numpy(), v)
    assert (m.N, m.V, m.F, m.equalised) == (3, 5, 7, False)


def test_edge_key_does_not_overflow(f3d):
    """The reference hashes edges in the face index type and overflows UInt32 past 65535 packed vertices
    (rep/mesh.jl:928-929); the build uses 64-bit keys."""
    nV = 70000
    faces = np.array([[0, 69998, 69999], [69997, 69998, 69999]], np.int32)
    m = f3d.TriMesh([np.zeros((nV, 3), np.float32)], [faces], device="cpu")
    assert np.array_equal(m.get_edges_packed(), np.array([[0, 69998], [0, 69999], [69997, 69998], [69997, 69999], [69998, 69999]]))


def test_shard_range(f3d):
    for total in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [f3d.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        f3d.shard_range(4, 4, 4)


def test_workspace_size_queries_are_host_only(f3d):
    """The *_workspace_bytes queries are pure host arithmetic (callable without a GPU): positive, 256-byte granular,
    monotone in the batch size, and 0 for invalid shapes — the caller sizes its device buffers from them."""
    L = f3d._lib.lib()
    w = L.f3d_chamfer_workspace_bytes(32, 4096, 4096)
    assert w > 0 and w % 256 == 0
    assert L.f3d_chamfer_workspace_bytes(64, 4096, 4096) > w
    assert L.f3d_chamfer_workspace_bytes(0, 4096, 4096) == 0 and L.f3d_chamfer_workspace_bytes(1, -5, 3) == 0
    wp = L.f3d_chamfer_pipe_workspace_bytes(32, 4096, 4096)
    # the host-array entry point adds the staging copies of both clouds to the sweep workspace
    assert wp >= w + 2 * 32 * 4096 * 12 and wp % 256 == 0
    assert L.f3d_chamfer_pipe_workspace_bytes(32, 0, 4096) == 0
    # ragged shapes: pads inside the last row block / column tile are accounted for
    assert L.f3d_chamfer_workspace_bytes(3, 257, 1025) >= L.f3d_chamfer_workspace_bytes(3, 256, 1024)


def test_workspace_sizes_are_host_only_and_sane(f3d):
    """The *_workspace_bytes entry points are pure host arithmetic (no CUDA call): callable on a machine without a GPU, monotone in
    the batch size, and the kNN workspace exists exactly where the TMA-fed Gram-filter path (knn_gram.cu) can serve the shape."""
    L = f3d._lib.lib()
    c2 = L.f3d_chamfer_workspace_bytes(32, 4096, 4096)
    assert 0 < c2 < (1 << 28)
    assert L.f3d_chamfer_workspace_bytes(64, 4096, 4096) > c2 > L.f3d_chamfer_workspace_bytes(8, 4096, 4096)
    assert L.f3d_chamfer_pipe_workspace_bytes(32, 4096, 4096) >= c2 + 2 * 32 * 4096 * 12    # + the two staging copies
    # kNN: (B, N, F, K)
    k3 = L.f3d_knn_graph_workspace_bytes(32, 1024, 3, 20)
    k64 = L.f3d_knn_graph_workspace_bytes(32, 1024, 64, 20)
    assert k3 >= 32 * 1024 * 128 and k64 >= 32 * 1024 * 256            # the operand image: 128 bytes per point and 32 features
    assert k64 > k3
    for shape in ((32, 4096, 3, 20), (32, 1024, 100, 20), (32, 1024, 3, 40), (2, 100, 3, 20)):   # N > 2048, F > 64, K > 31, too few chunks
        assert L.f3d_knn_graph_workspace_bytes(*shape) == 256
    assert L.f3d_sample_points_workspace_bytes(16, 2256) == 0           # the CDF fits in shared memory: fused path
    assert L.f3d_sample_points_workspace_bytes(4, 100000) >= 4 * 100000 * 8
