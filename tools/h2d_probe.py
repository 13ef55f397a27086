"""Where does the e2e time go?  H2D bandwidth by path (torch copy, raw cudaMemcpyAsync through the library's pipe)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flux3d_b200 as f3d

B, N, M = 32, 4096, 4096
A = torch.from_numpy(np.random.default_rng(201).random((B, N, 3), dtype=np.float32)).pin_memory()
Bc = torch.from_numpy(np.random.default_rng(202).random((B, M, 3), dtype=np.float32)).pin_memory()
dA, dB = torch.empty_like(A, device="cuda"), torch.empty_like(Bc, device="cuda")
nbytes = A.nbytes + Bc.nbytes

def ev_time(fn, reps=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, (time.perf_counter() - t0) / reps * 1e6

def torch_copy():
    dA.copy_(A, non_blocking=True); dB.copy_(Bc, non_blocking=True)
g, w = ev_time(torch_copy)
print(f"torch copy_ x2 (current stream): {g:.1f} us gpu, {w:.1f} us wall -> {nbytes/g/1e3:.1f} GB/s")
big = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device="cuda")
g, w = ev_time(lambda: dbig.copy_(big, non_blocking=True), 20)
print(f"torch copy_ 64 MiB: {g:.1f} us -> {big.nbytes/g/1e3:.1f} GB/s")
side = torch.cuda.Stream()
def torch_copy_side():
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dA.copy_(A, non_blocking=True); dB.copy_(Bc, non_blocking=True)
    torch.cuda.current_stream().wait_stream(side)
g, w = ev_time(torch_copy_side)
print(f"torch copy_ x2 (side stream + joins): {g:.1f} us gpu, {w:.1f} us wall")
def fwd():
    f3d.chamfer_forward_raw(dA, dB, 1.0, 1.0, want_indices=False)
g, w = ev_time(fwd)
print(f"chamfer_forward_raw (device resident, warm L2): {g:.1f} us gpu, {w:.1f} us wall")
def old_e2e():
    a = A.to("cuda", non_blocking=True); b = Bc.to("cuda", non_blocking=True)
    return f3d.chamfer_forward_raw(a, b, 1.0, 1.0, want_indices=False)[0].item()
g, w = ev_time(old_e2e)
print(f"old e2e (2 torch copies + fwd + item): {g:.1f} us gpu, {w:.1f} us wall")
for chunks in (8, 32, 128):
    g, w = ev_time(lambda: f3d.chamfer_forward_host(A, Bc, uploaders=chunks))
    print(f"pipe chunks={chunks} device loss, no sync: {g:.1f} us gpu, {w:.1f} us wall")
    g, w = ev_time(lambda: f3d.chamfer_forward_host(A, Bc, uploaders=chunks).item())
    print(f"pipe chunks={chunks} + item: {g:.1f} us gpu, {w:.1f} us wall")
    g, w = ev_time(lambda: f3d.chamfer_forward_host(A, Bc, uploaders=chunks, to_host=True))
    print(f"pipe chunks={chunks} to_host (mapped slot): {g:.1f} us gpu, {w:.1f} us wall")
g, w = ev_time(lambda: f3d.chamfer_distance(A, Bc).item())
print(f"public chamfer_distance(host, host).item(): {w:.1f} us wall")
# CPU cost of the C call alone
L = f3d._lib.lib()
t0 = time.perf_counter()
for _ in range(100):
    x = f3d.chamfer_forward_host(A, Bc, uploaders=0)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"pipe uploaders=0 host-side issue time: {(t1-t0)/100*1e6:.1f} us/call")
