#!/usr/bin/env python
"""bench.py — the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg5]

Metric (BASELINE.json): Chamfer point-pairs/s = B*N*M / t, whole job over all N GPUs.
A step = one chamfer_distance forward (both directions + the scalar loss; for N>1 also the one all-reduce of
the per-shard loss) over one batch of synthetic clouds.  Workload per GPU (weak scaling): cfg2 = BASELINE
configs[1], B=32 N=M=4096 Float32, U[0,1)^3, seeds 201/202 (+rank); --workload cfg5 gives the per-GPU shard of
configs[4] (B=32 per GPU, N=M=8192).

value      device-timed (CUDA events around every step, summed; inputs resident in HBM; L2 flushed between steps
           outside the event pairs), max over ranks.
e2e        the same step through the public API flux3d_b200.chamfer_distance with PINNED HOST inputs: H2D of
           both clouds and the D2H read of the loss are inside the timed region (C ABI: f3d_chamfer_pipe_run —
           the batch crosses PCIe in 4 chunks, chunk k+1 in flight while chunk k is swept).
roofline   the dominant kernel (chamfer_filter_sweep_kernel) timed alone, live, with CUDA events on its launch stream
           (F3D_FLAG_SWEEP_ONLY).  The binding roof is FP32 issue, not HBM (0.006 algorithmic bytes per pair):
           achieved = ALGORITHMIC 8 lane-instructions/pair (3 FSUB, 3 FMUL, 2 FADD: the reference's bit-exact direct
           form; the two min-updates not counted) * pairs / t against SMs*128 lanes*f_max.  The kernel
           reaches the same bit-exact result with 4 executed FP32 lane-ops per pair (expanded-form filter +
           certified exact re-evaluation), reported as roofline.executed; the HBM view BASELINE.json asks for is
           reported beside it under roofline.hbm.
cpu_baseline  the reference's CPU algorithm (per batch element a KD-tree build + 1-NN queries per direction,
           serial, src/metrics/pcloud.jl:54-70) restated with scipy's cKDTree, 1 thread, on this box's host cores.

--impl reference times that same CPU restatement with all host threads (the reference is pure Julia; no Julia
exists in this image) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"cfg2": dict(B=32, N=4096, M=4096, seeds=(201, 202)),
             "cfg5": dict(B=32, N=8192, M=8192, seeds=(501, 502))}
LANE_INSTR_PER_PAIR = 8
SMS, LANES = 148, 128


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def make_inputs(wl, rank):
    import numpy as np
    w = WORKLOADS[wl]
    A = np.random.default_rng(w["seeds"][0] + 1000 * rank).random((w["B"], w["N"], 3), dtype=np.float32)
    B = np.random.default_rng(w["seeds"][1] + 1000 * rank).random((w["B"], w["M"], 3), dtype=np.float32)
    return A, B


SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
t_end = time.time() + 600  # never outlive a bench that died without stopping us
while time.time() < t_end:
    print(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
          nv.nvmlDeviceGetCurrentClocksEventReasons(h), flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (every ~2 ms, NVML — the same counters as
    B200_PROFILING.md's nvidia-smi clocks line, fast enough for a millisecond-scale region).  The sampler is a separate
    PROCESS: a thread in this interpreter takes the GIL at every wake-up and shows up in the wall-clock e2e number."""

    def __init__(self, gpu_index):
        self.idx, self.proc, self.err = gpu_index, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", SAMPLER_SRC, str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.PIPE, text=True)
            self.first = self.proc.stdout.readline()  # "max <MHz>": the sampler is up before the timed region starts
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def stop(self):
        samples, max_mhz = [], None
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, errtxt = self.proc.communicate(timeout=5)
            except Exception as e:  # pragma: no cover
                out, errtxt = "", repr(e)
            for line in [self.first] + out.splitlines():
                f = line.split()
                if len(f) == 2 and f[0] == "max":
                    max_mhz = float(f[1])
                elif len(f) == 3:
                    samples.append((float(f[0]), float(f[1]), int(f[2])))
            if not samples:
                self.err = (errtxt or "").strip()[-200:]
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"no samples ({self.err})"]}
        import pynvml as nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, b in bits.items() if any(r & b for _, _, r in samples))
        sm = sorted(c for c, _, _ in samples)
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_min_mhz": float(sm[0]), "sm_max_mhz": max_mhz,
                "power_w_max": max(p for _, p, _ in samples), "samples": len(sm), "reasons": reasons}


def kdtree_step(A, B, workers):
    from oracle import oracle as O
    return float(O.kdtree_chamfer(A, B, workers=workers))


def run_reference(args):
    """Reference arm: the reference's CPU path (KD-tree chamfer) on the host cores, all threads it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    A, B = make_inputs(args.workload, 0)
    cores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        kdtree_step(A[:2], B[:2], -1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = kdtree_step(A, B, -1)
    dt = time.perf_counter() - t0
    pairs = wl["B"] * wl["N"] * wl["M"]
    value = pairs * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "chamfer_point_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"chamfer_distance B={wl['B']} N={wl['N']} M={wl['M']} Float32 ({args.workload})",
                   "note": "reference CPU algorithm (KD-tree 1-NN per batch element and direction, "
                           "src/metrics/pcloud.jl:54-70) restated with scipy cKDTree; the Julia reference cannot run "
                           "in this image; rank 0 only, one full batch per step"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"full batch per step, {args.steps} steps, cKDTree workers=-1"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss}))


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import flux3d_b200 as f3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    multi = world > 1
    if multi:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    Bn, N, M = wl["B"], wl["N"], wl["M"]
    B_total = Bn * world
    hA, hB = make_inputs(args.workload, rank)
    pA, pB = torch.from_numpy(hA).pin_memory(), torch.from_numpy(hB).pin_memory()
    dA, dB = pA.to(dev), pB.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out = (torch.empty(3, dtype=torch.float32, device=dev), None, None)
    stream = torch.cuda.current_stream(dev)

    # N > 1: the single exchange of the path (the sum of the shard losses) is fused into the finalize kernel — peer
    # mailboxes mapped over NVLink (f3d_comm_enable_p2p); if peer mapping is not possible, one NCCL all-reduce instead
    comm, exchange = None, "none"
    if multi:
        exchange = "nccl_allreduce"
        try:
            if os.environ.get("F3D_BENCH_EXCHANGE", "fused") == "nccl":  # A/B aid
                raise RuntimeError("F3D_BENCH_EXCHANGE=nccl")
            comm = f3d.Communicator(rank, world, dev).enable_p2p()
            exchange = "fused_peer_mailboxes"
        except Exception as e:  # pragma: no cover
            comm = None
            print(f"[rank {rank}] peer mailboxes unavailable ({e}); using NCCL all-reduce", file=sys.stderr)
        ok = torch.tensor([1.0 if comm is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks take the same path
        if ok.item() == 0.0:
            comm, exchange = None, "nccl_allreduce"

    def step(flags=0):
        if multi and not flags:
            return f3d.chamfer_distance_sharded(dA, dB, B_total, comm=comm).reshape(1)
        loss, _, _, _ = f3d.chamfer_forward_raw(dA, dB, 1.0, 1.0, batch_total=B_total, want_indices=False, flags=flags, out=out)
        return loss

    def barrier():
        torch.cuda.synchronize(dev)
        if multi:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for e0, e1 in evs:
            flush.zero_()           # L2 flush, outside the event pair
            e0.record(stream)
            fn()
            e1.record(stream)
        barrier()
        ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
        if multi:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = timed(step, args.steps)
    loss_val = float(step().item())
    # dominant kernel alone (live, CUDA events on the launch stream)
    for _ in range(3):
        step(f3d.metrics.FLAG_SWEEP_ONLY)
    sweep_ms = timed(lambda: step(f3d.metrics.FLAG_SWEEP_ONLY), args.steps) / args.steps
    # end to end through the public API: pinned host inputs → H2D → kernels → D2H of the loss
    def e2e_step():
        with torch.no_grad():
            # host arrays in: f3d_chamfer_pipe_run uploads chunk k+1 while chunk k is swept (one C call per step)
            if multi:
                return float(f3d.chamfer_distance_sharded(pA, pB, B_total, comm=comm, to_host=comm is not None).item())
            return float(f3d.chamfer_distance(pA, pB).item())
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_loss = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if multi:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        pk = peaks()
        pairs_step = B_total * N * M
        value = pairs_step * args.steps / (total_ms * 1e-3)
        pairs_launch = Bn * N * M
        lane_peak = SMS * LANES * pk["sm_max_mhz"] * 1e6
        achieved = LANE_INSTR_PER_PAIR * pairs_launch / (sweep_ms * 1e-3)
        alg_bytes = 12 * Bn * (N + M) + 4
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        line = {
            "metric": "chamfer_point_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"chamfer_distance B={Bn}/GPU (global {B_total}) N={N} M={M} Float32 ({args.workload}), "
                                   "U[0,1)^3, forward incl. loss reduction" + (f" + cross-rank loss sum ({exchange})" if multi else ""),
                       "parallelism": f"batch-sharded x{world}", "l2": "flushed between steps (256 MiB memset outside the event pairs)",
                       "arithmetic": "results bit-identical to the direct form ((dx*dx)+(dy*dy))+(dz*dz) without FMA contraction "
                                     "(expanded-form FP32 filter, every reported distance/index re-evaluated exactly)"},
            "clocks": clocks,
            "e2e": {"value": pairs_step * args.steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(pA.nbytes + pB.nbytes),
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": 2 * args.steps,  # chamfer_filter_sweep_kernel + chamfer_filter_finalize_kernel per step (+1 memset node)
            "roofline": {"bound": "fp32_issue", "kernel": "chamfer_filter_sweep_kernel", "achieved": achieved / 1e12,
                         "peak": lane_peak / 1e12, "unit": "Tlane-instr/s", "frac": achieved / lane_peak,
                         "kernel_ms": sweep_ms, "kernel_share_of_step": sweep_ms / (total_ms / args.steps),
                         "pairs_per_s_kernel": pairs_launch / (sweep_ms * 1e-3),
                         "peak_source": f"{SMS} SMs x {LANES} lanes x {pk['sm_max_mhz']:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz, {pk['source']})",
                         "traffic": traffic,
                         "executed": {"fp32_lane_ops_per_pair": 4, "achieved": 4 * pairs_launch / (sweep_ms * 1e-3) / 1e12,
                                      "frac": 4 * pairs_launch / (sweep_ms * 1e-3) / lane_peak,
                                      "note": "3 FFMA2 + 1 FADD2 per two pairs actually issued by the filter sweep"},
                         "hbm": {"bound": "hbm", "achieved": alg_bytes / (sweep_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": alg_bytes / (sweep_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": alg_bytes,
                                 "note": f"of {pk['source']}; 0.006 B/pair: HBM cannot bind a brute-force sweep"}},
            "loss": loss_val, "e2e_loss": e2e_loss,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded CPU sample: first 8 batch elements of the same workload, 1 thread (the reference is serial)
            nb = min(8, Bn)
            kdtree_step(hA[:1], hB[:1], 1)
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or time.perf_counter() - t0 < 5.0:
                kdtree_step(hA[:nb], hB[:nb], 1)
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            from oracle import oracle as O
            t1 = time.perf_counter()
            O.chamfer_distance(hA[:nb], hB[:nb])
            dt_bf = time.perf_counter() - t1
            line["cpu_baseline"] = {"value": nb * N * M / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                    "sample": f"first {nb} of {Bn} batch elements, {reps} reps, KD-tree (scipy cKDTree) "
                                              "restatement of src/metrics/pcloud.jl:54-70",
                                    "host_cores_available": os.cpu_count(),
                                    "brute_force_oracle_all_cores": {"value": nb * N * M / dt_bf, "unit": "pairs/s",
                                                                     "cores": O.num_threads()}}
        print(json.dumps(line))
    if multi:
        if comm is not None:
            comm.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
