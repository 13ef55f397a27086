"""Development aid: A/B device time of f3d_knn_graph (cfg3: B=32 N=1024 K=20, F=3 and F=64, CUDA-core path forced) between builds."""
import ctypes as C, sys
import torch
for path in sys.argv[1:]:
    L = C.CDLL(path)
    L.f3d_knn_graph.restype = C.c_int32
    L.f3d_knn_graph.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
    out = []
    for (B, N, F, K, flags) in [(32, 1024, 3, 20, 0), (32, 1024, 3, 10, 0), (32, 1024, 64, 20, 4), (32, 4096, 3, 20, 0), (8, 2048, 6, 40, 0)]:
        g = torch.Generator(device="cuda").manual_seed(3)
        X = torch.randn((B, N, F), device="cuda", generator=g)
        idx = torch.empty((B, N, K), dtype=torch.int32, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        run = lambda: L.f3d_knn_graph(X.data_ptr(), B, N, F, K, idx.data_ptr(), None, None, None, None, 0, flags, st)
        for _ in range(5): assert run() == 0
        torch.cuda.synchronize(); print("ok", B, N, F, K, flush=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): run()
        e1.record(); torch.cuda.synchronize()
        out.append("B%d N%d F%d K%d: %.1f us (chk %d)" % (B, N, F, K, e0.elapsed_time(e1) / 30 * 1e3, int(idx.long().sum().item() % 100003)))
    print(path.split("/")[-1], " | ".join(out), flush=True)
