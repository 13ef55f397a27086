#!/usr/bin/env python
"""Secondary measurements for the other SURVEY §8 rows (bench.py is the headline contract: chamfer).

For each op: device time per call (CUDA events, 5 warm-ups, L2 flushed between timed calls), the algorithmic
bytes per call (SURVEY §8d / DESIGN.md §3) and the HBM fraction they imply, and the CPU oracle timed on a
bounded sample beside it.  One JSON line per op on stdout.

    python tools/bench_ops.py [--steps 30] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import flux3d_b200 as f3d


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timed(fn, steps, flush):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        flush.zero_()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    return ts[len(ts) // 2] * 1e3  # median, µs


def cpu_time(fn, min_s=2.0):
    fn()
    t0 = time.perf_counter()
    n = 0
    while n < 2 or time.perf_counter() - t0 < min_s:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from oracle import oracle as O
    from fixtures import pad, teapots
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = hbm_peak()
    out = []

    def report(name, units, unit_name, us, alg_bytes, cpu_us=None, cpu_note="", extra=None):
        line = {"op": name, "us_per_call": us, "value": units / (us * 1e-6), "unit": unit_name + "/s",
                "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (us * 1e-6) / 1e9,
                "hbm_frac_of_measured": alg_bytes / (us * 1e-6) / 1e9 / peak}
        if cpu_us is not None:
            line["cpu_oracle"] = {"us_per_call": cpu_us, "value": units / (cpu_us * 1e-6), "cores": O.num_threads(), "note": cpu_note}
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)
        out.append(line)

    # ---- cfg3: DGCNN EdgeConv kNN graph, B=32 N=1024 K=20, F=3 and F=64 ------------------------------------
    rng = np.random.default_rng(301)
    X3 = rng.standard_normal((32, 1024, 3)).astype(np.float32)
    X3 = ((X3 - X3.mean(1, keepdims=True)) / X3.std(axis=(1, 2), keepdims=True)).astype(np.float32)
    X64 = np.random.default_rng(302).standard_normal((32, 1024, 64)).astype(np.float32)
    for name, X in (("knn_graph cfg3 F=3 K=20 (idx only)", X3), ("knn_graph cfg3 F=64 K=20 (idx only)", X64)):
        t = torch.from_numpy(X).cuda()
        B, N, F = X.shape
        us = timed(lambda: f3d.knn_graph(t, 20), args.steps, flush)
        cpu = None if args.no_cpu else cpu_time(lambda: O.knn_graph(X[:4], 20)) * (B / 4)
        report(name, B * N * N, "pairs", us, X.nbytes + B * N * 20 * 4, cpu, "oracle brute force + qsort, first 4 clouds scaled to 32")
    t3 = torch.from_numpy(X3).cuda()
    us = timed(lambda: f3d.knn_graph(t3, 20, want_edge=True), args.steps, flush)
    report("knn_graph cfg3 F=3 K=20 + edge features (2F,K,N,B)", 32 * 1024 * 1024, "pairs", us, X3.nbytes + 32 * 1024 * 20 * (4 + 24))
    t64 = torch.from_numpy(X64).cuda()
    us = timed(lambda: f3d.knn_graph(t64, 20, want_edge=True), args.steps, flush)
    report("knn_graph cfg3 F=64 K=20 + edge features (2F,K,N,B)", 32 * 1024 * 1024, "pairs", us, X64.nbytes + 32 * 1024 * 20 * (4 + 512))

    # ---- cfg4: 16 teapots (V=1202, F=2256), S=10000 --------------------------------------------------------
    gold = os.path.join(ROOT, "tests", "golden")
    vl, fl = teapots(16, gold, O)
    m = f3d.TriMesh(vl, fl)
    m._topology(); m.get_verts_padded(); m.faces_padded_device(); m._topo_device("rowptr"); m._topo_device("v2c")
    nV, nF, nE = 16 * 1202, 16 * 2256, 16 * 3456
    vp, fp, vlen, flen = pad(vl, fl)
    vpk = np.concatenate(vl)
    fpk = m.get_faces_packed()
    us = timed(lambda: f3d.sample_points(m, 10000, seed=401), args.steps, flush)
    cpu = None if args.no_cpu else cpu_time(lambda: O.sample_points(vp, fp, vlen, flen, 10000, seed=401))
    report("sample_points cfg4 (16 meshes x 10000)", 160000, "samples", us, nV * 12 + nF * 12 + 160000 * 12, cpu, "oracle, 1 thread")
    us = timed(lambda: f3d.laplacian_loss(m), args.steps, flush)
    cpu = None if args.no_cpu else cpu_time(lambda: O.laplacian_loss(vpk, fpk))
    report("laplacian_loss cfg4", nV, "vertices", us, nV * 12 + (2 * nE + nV) * 8 + (nV + 1) * 4 + 4, cpu, "oracle incl. its topology build, 1 thread")
    for mode, nm in ((0, "REFERENCE_CPU"), (1, "ACCUMULATE")):
        us = timed(lambda: m.compute_verts_normals_packed(mode), args.steps, flush)
        cpu = None if args.no_cpu else cpu_time(lambda: O.verts_normals(vpk, fpk, mode))
        report(f"compute_verts_normals_packed cfg4 ({nm})", nV, "vertices", us, nV * 12 + nF * 12 + nV * 12 + nF * 12 + (nV + 1) * 4, cpu, "oracle, 1 thread")
    us = timed(lambda: m.compute_faces_areas_packed(), args.steps, flush)
    report("compute_faces_areas_packed cfg4", nF, "faces", us, nV * 12 + nF * 12 + nF * 4)
    us = timed(lambda: f3d.edge_loss(m), args.steps, flush)
    report("edge_loss cfg4", nE, "edges", us, nV * 12 + nE * 8 + 4)
    # the fit_mesh objective's forward: sample both meshes + chamfer + laplacian + edge (examples/fit_mesh.jl:78-84)
    m2 = f3d.TriMesh([v * np.float32(1.05) for v in vl], fl)
    m2._topology(); m2.get_verts_padded(); m2.faces_padded_device()

    def fit_step():
        a = f3d.sample_points(m, 10000, seed=1)
        b = f3d.sample_points(m2, 10000, seed=2)
        return f3d.chamfer_distance(a, b) + 0.1 * f3d.laplacian_loss(m) + f3d.edge_loss(m)
    us = timed(fit_step, args.steps, flush)
    report("fit_mesh objective forward cfg4 (2x sample_points + chamfer S=10000 + laplacian + edge)", 16 * 10000 * 10000, "pairs", us,
           2 * (nV * 12 + nF * 12 + 160000 * 12) + 2 * 160000 * 12)

    # the whole fit_mesh step — forward AND pullbacks — eager and as ONE CUDA-graph launch (flux3d_b200.capture_step)
    nVt = sum(len(v) for v in vl)
    delta = torch.zeros((nVt, 3), device="cuda", requires_grad=True)
    delta.grad = torch.zeros_like(delta)
    c1 = torch.zeros(1, dtype=torch.int64, device="cuda"); c2 = torch.zeros(1, dtype=torch.int64, device="cuda")

    def train_step():
        delta.grad.zero_()
        md = f3d.offset(m, delta)
        a = f3d.sample_points(md, 10000, seed=1, counter=c1)
        b = f3d.sample_points(m2, 10000, seed=2, counter=c2)
        loss = f3d.chamfer_distance(a, b) + 0.1 * f3d.laplacian_loss(md) + f3d.edge_loss(md)
        loss.backward()
        return loss
    us_eager = timed(train_step, args.steps, flush)
    graphed = f3d.capture_step(train_step)
    us_graph = timed(graphed, args.steps, flush)
    report("fit_mesh step cfg4, forward + pullbacks, ONE CUDA-graph launch", 16 * 10000 * 10000, "pairs", us_graph,
           2 * (nV * 12 + nF * 12 + 160000 * 12) + 2 * 160000 * 12, extra={"eager_us_per_call": us_eager})

    # ---- chamfer backward at cfg2 --------------------------------------------------------------------------
    A = torch.rand((32, 4096, 3), device="cuda")
    Bc = torch.rand((32, 4096, 3), device="cuda")
    _, _, nnA, nnB = f3d.chamfer_forward_raw(A, Bc, 1.0, 1.0)
    gA, gB, gout = torch.empty_like(A), torch.empty_like(Bc), torch.ones(1, device="cuda")
    L, ptr = f3d._lib.lib(), f3d._lib.ptr
    stream = torch.cuda.current_stream().cuda_stream

    def bwd():   # the C entry point itself (through autograd the host side of one call takes longer than the kernel)
        f3d._lib.check(L.f3d_chamfer_bwd(ptr(A), ptr(Bc), 32, 4096, 4096, 1.0, 1.0, 0, ptr(nnA), ptr(nnB), ptr(gout), ptr(gA), ptr(gB), stream))
    us = timed(bwd, args.steps, flush)
    report("chamfer backward cfg2", 2 * 32 * 4096, "points", us, 2 * 32 * 4096 * (12 + 4 + 12) + 2 * 32 * 4096 * 12)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
