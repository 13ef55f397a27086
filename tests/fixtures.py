"""Literal fixtures transcribed from the reference's own tests (data, not code) — 0-based faces here.

MESH3: test/rep.jl:59-93 == test/metrics.jl:10-43 (three small meshes: 3/4/5 verts, 1/2/7 faces).
NORMALS: test/rep.jl:178-260 (two meshes with hand-checked vertex normals, face normals and face areas,
4-decimal goldens; constructed so that no vertex repeats in the same corner slot, hence both normals
modes must reproduce them)."""
import numpy as np

MESH3_VERTS = [
    np.array([[0.1, 0.3, 0.5], [0.5, 0.2, 0.1], [0.6, 0.8, 0.7]], np.float32),
    np.array([[0.1, 0.3, 0.3], [0.6, 0.7, 0.8], [0.2, 0.3, 0.4], [0.1, 0.5, 0.3]], np.float32),
    np.array([[0.7, 0.3, 0.6], [0.2, 0.4, 0.8], [0.9, 0.5, 0.2], [0.2, 0.3, 0.4], [0.9, 0.3, 0.8]], np.float32),
]
MESH3_FACES = [
    np.array([[1, 2, 3]], np.int32) - 1,
    np.array([[1, 2, 3], [2, 3, 4]], np.int32) - 1,
    np.array([[2, 3, 1], [1, 2, 4], [3, 4, 2], [5, 4, 3], [5, 1, 2], [5, 4, 2], [5, 3, 2]], np.int32) - 1,
]

NORMALS_VERTS = [
    np.array([[0.1, 0.3, 0.0], [0.5, 0.2, 0.0], [0.6, 0.8, 0.0], [0.0, 0.3, 0.2], [0.0, 0.2, 0.5], [0.0, 0.8, 0.7],
              [0.5, 0.0, 0.2], [0.6, 0.0, 0.5], [0.8, 0.0, 0.7], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]],
             np.float32),
    np.array([[0.1, 0.3, 0.0], [0.5, 0.2, 0.0], [0.0, 0.3, 0.2], [0.0, 0.2, 0.5], [0.0, 0.8, 0.7]], np.float32),
]
NORMALS_FACES = [
    np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]], np.int32) - 1,
    np.array([[1, 2, 3], [3, 4, 5]], np.int32) - 1,
]
GOLD_VNORMALS = [
    np.array([[0, 0, 1]] * 3 + [[-1, 0, 0]] * 3 + [[0, 1, 0]] * 3 + [[0, 0, 0]] * 3, np.float32),
    np.array([[-0.2408, -0.9631, -0.1204], [-0.2408, -0.9631, -0.1204], [-0.9389, -0.3414, -0.0427],
              [-1.0, 0.0, 0.0], [-1.0, 0.0, 0.0]], np.float32),
]
GOLD_FNORMALS = [
    np.array([[0, 0, 1], [-1, 0, 0], [0, 1, 0], [0, 0, 0]], np.float32),
    np.array([[-0.2408, -0.9631, -0.1204], [-1.0, 0.0, 0.0]], np.float32),
]
GOLD_FAREAS = [np.array([0.125, 0.1, 0.02, 0.0], np.float32), np.array([0.0415, 0.1], np.float32)]

TEAPOT_LAPLACIAN_LOSS = np.float32(0.05888283)  # README.md:111-112


def pack(verts_list, faces_list):
    """packed verts (ΣV,3) and packed faces (ΣF,3) with global ids — src/rep/mesh.jl:884-896."""
    offs = np.cumsum([0] + [v.shape[0] for v in verts_list])
    return (np.concatenate(verts_list).astype(np.float32),
            np.concatenate([f + offs[i] for i, f in enumerate(faces_list)]).astype(np.int32))


def pad(verts_list, faces_list):
    N, V, F = len(verts_list), max(v.shape[0] for v in verts_list), max(f.shape[0] for f in faces_list)
    vp = np.zeros((N, V, 3), np.float32)
    fp = np.full((N, F, 3), -1, np.int32)
    for i, (v, f) in enumerate(zip(verts_list, faces_list)):
        vp[i, :v.shape[0]] = v
        fp[i, :f.shape[0]] = f
    return vp, fp, np.array([v.shape[0] for v in verts_list], np.int32), np.array([f.shape[0] for f in faces_list], np.int32)


def teapots(n, golden_dir, oracle):
    """cfg4 meshes: n copies of teapot.obj, copy i scaled by (1+0.01 i) and shifted by 0.1 i."""
    import os
    v, f = oracle.load_obj(os.path.join(golden_dir, "teapot.obj"))
    vl = [(v * np.float32(1 + 0.01 * i) + np.float32(0.1 * i)).astype(np.float32) for i in range(n)]
    return vl, [f.copy() for _ in range(n)]
