"""ctypes binding of libflux3d_b200.so — the same symbols the Julia ccall shim binds
(include/flux3d_b200.h).  There is NO fallback: if the library is missing or a call fails, this
raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLUX3D_B200_LIB", os.path.join(_HERE, "libflux3d_b200.so"))

_f32p = C.c_void_p  # device pointers travel as plain addresses
_i32p = C.c_void_p
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/flux3d_b200.h one to one
SIGNATURES = {
    "f3d_version": (C.c_int32, []),
    "f3d_last_error": (C.c_int32, [C.c_char_p, C.c_size_t]),
    "f3d_chamfer_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "f3d_chamfer_fwd": (C.c_int32, [_f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                    C.c_int32, _f32p, _f32p, _i32p, _i32p, _vp, C.c_size_t, C.c_int32, _vp]),
    "f3d_chamfer_pipe_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "f3d_chamfer_pipe_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "f3d_chamfer_pipe_run": (C.c_int32, [_vp, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                         C.c_int32, _f32p, _f32p, _vp, C.c_size_t, C.c_int32, _vp, _vp]),
    "f3d_chamfer_pipe_destroy": (C.c_int32, [_vp]),
    "f3d_chamfer_bwd": (C.c_int32, [_f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                    C.c_int32, _i32p, _i32p, _f32p, _f32p, _f32p, _vp]),
    "f3d_knn_graph_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "f3d_knn_graph": (C.c_int32, [_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p, _f32p, _f32p,
                                  _f32p, _vp, C.c_size_t, C.c_int32, _vp]),
    "f3d_faces_areas_normals": (C.c_int32, [_f32p, _i32p, C.c_int32, C.c_int32, _f32p, _f32p, _vp]),
    "f3d_mesh_topology_build_host": (C.c_int32, [_i32p, C.c_int32, C.c_int32, _i32p, C.POINTER(C.c_int32), _i32p,
                                                 _i32p, _i32p, _f32p, _i32p, _i32p]),
    "f3d_packed_to_padded": (C.c_int32, [_vp, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, _vp, _vp]),
    "f3d_padded_to_packed": (C.c_int32, [_vp, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "f3d_verts_normals": (C.c_int32, [_f32p, _i32p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, _f32p, _vp]),
    "f3d_laplacian_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "f3d_laplacian_loss": (C.c_int32, [_f32p, _i32p, _i32p, _f32p, C.c_int32, C.c_int32, _f32p, _vp, C.c_size_t, _vp]),
    "f3d_laplacian_loss_bwd": (C.c_int32, [_f32p, _i32p, _i32p, _f32p, C.c_int32, C.c_int32, _f32p, _f32p, _vp,
                                           C.c_size_t, _vp]),
    "f3d_edge_loss_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "f3d_edge_loss": (C.c_int32, [_f32p, _i32p, C.c_int32, C.c_int32, C.c_float, _f32p, _vp, C.c_size_t, _vp]),
    "f3d_sample_points_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "f3d_sample_points": (C.c_int32, [_f32p, _i32p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_double, C.c_uint64, C.c_uint64, _i32p, _f32p, _f32p, _f32p, _i32p, _f32p,
                                      _vp, C.c_size_t, _vp]),
    "f3d_sample_points_replayable": (C.c_int32, [_f32p, _i32p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_double, C.c_uint64, C.c_uint64, _vp, _i32p, _f32p, _f32p, _f32p, _i32p, _f32p,
                                                 _vp, C.c_size_t, _vp]),
    "f3d_sample_points_bwd": (C.c_int32, [_f32p, _i32p, _f32p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _f32p, _vp]),
    "f3d_edge_loss_bwd": (C.c_int32, [_f32p, _i32p, _i32p, C.c_int32, C.c_int32, C.c_float, _f32p, _f32p, _vp]),
    "f3d_comm_unique_id_host": (C.c_int32, [_vp]),
    "f3d_comm_init": (C.c_int32, [C.c_int32, C.c_int32, _vp, C.POINTER(C.c_void_p)]),
    "f3d_allreduce_sum_f32": (C.c_int32, [_vp, _f32p, C.c_int32, _vp]),
    "f3d_comm_enable_p2p": (C.c_int32, [_vp, _vp]),
    "f3d_chamfer_fwd_allreduce": (C.c_int32, [_vp, _f32p, _f32p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                              C.c_int32, _f32p, _i32p, _i32p, _vp, C.c_size_t, C.c_int32, _vp]),
    "f3d_comm_destroy": (C.c_int32, [_vp]),
}

_lib = None


class Flux3DB200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built — there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Flux3DB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(flux3d.jl_b200 has no CPU or PyTorch fallback).")
        L = C.CDLL(LIB_PATH)
        missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing:  # a stale / partial build is an error, never a reason to fall back
            raise Flux3DB200Error(f"{LIB_PATH} does not export {missing}; rebuild it (make -C flux3d.jl_b200/csrc)")
        _lib = L
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().f3d_last_error(buf, 512)
    return buf.value.decode()


def check(status: int):
    if status != 0:
        raise Flux3DB200Error(f"libflux3d_b200 status {status}: {last_error()}")


def ptr(t):
    """Device (or host, for *_host arguments) address of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


# ---- caller-owned device workspaces (the library allocates nothing; torch owns the memory) ----------
_ws_cache: dict = {}


def workspace(key, nbytes: int, device):
    """A >= nbytes uint8 device buffer, cached per (op key, device, stream) the way CUDA.jl would keep a
    scratch CuArray.  256-byte aligned (torch's caching allocator guarantees 512)."""
    import torch
    k = (key, str(device), torch.cuda.current_stream(device).cuda_stream)
    t = _ws_cache.get(k)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[k] = t
    return t


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
