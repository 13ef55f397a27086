"""Development aid: A/B device time of f3d_chamfer_fwd between library builds (argv: .so paths) over several shapes,
bound directly with ctypes (works for builds that predate newer entry points)."""
import ctypes as C, sys
import torch
shapes = [(32, 4096, 4096), (32, 8192, 8192), (16, 10000, 10000), (2, 1024, 1024), (8, 1000, 500), (64, 2048, 2048), (1, 16384, 16384)]
for path in sys.argv[1:]:
    L = C.CDLL(path)
    L.f3d_chamfer_workspace_bytes.restype = C.c_size_t
    L.f3d_chamfer_workspace_bytes.argtypes = [C.c_int32] * 3
    L.f3d_chamfer_fwd.restype = C.c_int32
    L.f3d_chamfer_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
    out = []
    for (B, N, M) in shapes:
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.rand((B, N, 3), device="cuda", generator=g); Bc = torch.rand((B, M, 3), device="cuda", generator=g)
        ws = torch.empty(L.f3d_chamfer_workspace_bytes(B, N, M), dtype=torch.uint8, device="cuda")
        res = torch.empty(3, device="cuda"); nnA = torch.empty((B, N), dtype=torch.int32, device="cuda"); nnB = torch.empty((B, M), dtype=torch.int32, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        def run(idx):
            rc = L.f3d_chamfer_fwd(A.data_ptr(), Bc.data_ptr(), B, N, M, 1.0, 1.0, 0, res.data_ptr(), res.data_ptr() + 4,
                                   nnA.data_ptr() if idx else None, nnB.data_ptr() if idx else None, ws.data_ptr(), ws.numel(), 0, st)
            assert rc == 0
        ts = []
        for idx in (False, True):
            for _ in range(5): run(idx)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30): run(idx)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 30 * 1e3)
        out.append("%dx%dx%d %.1f/%.1f (%.6g)" % (B, N, M, ts[0], ts[1], res[0].item()))
    print(path.split("/")[-1], " | ".join(out), flush=True)
