"""flux3d.jl_b200 — Blackwell-native (sm_100a) implementation of the Flux3D.jl batched 3D-metric hot
path behind the reference's own interface: chamfer_distance, the DGCNN kNN graph, sample_points,
laplacian_loss / compute_verts_normals_packed, with PointCloud / TriMesh containers.

Import as ``flux3d_b200`` (see flux3d_b200.py at the repo root).  All compute is in
libflux3d_b200.so (csrc/, C ABI in include/flux3d_b200.h); torch only provides device memory, streams
and torch.distributed plumbing."""
from . import _lib
from ._lib import Flux3DB200Error, LIB_PATH
from .pcloud import PointCloud
from .metrics import (FLAG_FMA, FLAG_NONE, chamfer_distance, chamfer_forward_raw, nearest_neighbors)

__all__ = ["PointCloud", "chamfer_distance", "chamfer_forward_raw", "nearest_neighbors", "FLAG_FMA", "FLAG_NONE",
           "Flux3DB200Error", "LIB_PATH"]
__version__ = "0.1.0"
