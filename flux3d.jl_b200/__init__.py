"""flux3d.jl_b200 — Blackwell-native (sm_100a) implementation of the Flux3D.jl batched 3D-metric hot
path behind the reference's own interface: chamfer_distance, the DGCNN kNN graph, sample_points,
laplacian_loss / compute_verts_normals_packed, with PointCloud / TriMesh containers.

Import as ``flux3d_b200`` (see flux3d_b200.py at the repo root).  All compute is in
libflux3d_b200.so (csrc/, C ABI in include/flux3d_b200.h); torch only provides device memory, streams
and torch.distributed plumbing."""
from . import _lib
from ._lib import Flux3DB200Error, LIB_PATH
from .pcloud import PointCloud
from .metrics import (FLAG_CUDA_CORES, FLAG_EXACT_SWEEP, FLAG_FMA, FLAG_NONE, FLAG_SWEEP_ONLY, FLAG_TENSOR, chamfer_distance, chamfer_forward_host, chamfer_forward_raw,
                      nearest_neighbors)
from .mesh import (NORMALS_ACCUMULATE, NORMALS_REFERENCE_CPU, TriMesh, compute_faces_areas_packed,
                   compute_faces_normals_packed, compute_verts_normals_packed, edge_loss, get_edges_packed,
                   get_laplacian_packed, get_verts_packed, laplacian_loss, load_trimesh, offset, packed_to_padded,
                   padded_to_packed)
from .sampling import sample_points
from .dgcnn import create_single_knn_graph, edgeconv_features, knn_graph
from .distributed import Communicator, allreduce_loss_, chamfer_distance_sharded, shard_range
from .graph import CapturedStep, capture_step

__all__ = ["CapturedStep", "capture_step", "PointCloud", "TriMesh", "chamfer_distance", "chamfer_forward_raw", "chamfer_forward_host", "nearest_neighbors", "laplacian_loss",
           "edge_loss", "sample_points", "knn_graph", "create_single_knn_graph", "edgeconv_features",
           "compute_verts_normals_packed", "compute_faces_normals_packed", "compute_faces_areas_packed",
           "get_verts_packed", "get_edges_packed", "get_laplacian_packed", "load_trimesh", "offset", "packed_to_padded",
           "padded_to_packed",
           "shard_range", "chamfer_distance_sharded", "allreduce_loss_", "Communicator",
           "NORMALS_REFERENCE_CPU", "NORMALS_ACCUMULATE", "FLAG_FMA", "FLAG_NONE", "FLAG_EXACT_SWEEP", "FLAG_SWEEP_ONLY", "FLAG_TENSOR", "FLAG_CUDA_CORES", "Flux3DB200Error", "LIB_PATH"]
__version__ = "0.1.0"
