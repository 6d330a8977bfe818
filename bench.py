#!/usr/bin/env python
"""bench.py - Horn-Schunck Mpixel-iterations/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload 1080p|4k|kitti|batch256|slab16k] [--window 3|5] [--iters T] [--k K] [--textbook]

A *step* is one complete solve of the workload: gradient/coefficient stage + T Jacobi sweeps.
  value  : whole-job Mpixel-iterations/s with the frames already resident in HBM
           (CUDA events on the launching stream, max over ranks)
  e2e    : the same metric through the reference-facing call hs_solve() with HOST buffers:
           H2D of both uint8 frames and D2H of u, v as float64 (the reference's CV_64FC1) inside
           the timed region
  roofline: the fused Jacobi kernel against the measured HBM peak, 32 algorithmic bytes per
           pixel-iteration (SURVEY.md 8d / DESIGN.md), duration from CUDA events around the sweeps
  cpu_baseline / --impl reference: the reference's CPU path (oracle/hs_oracle.py::cv_flow, the
           line-by-line cv2 restatement of hornSchunck.cpp - C++ OpenCV is not in this image, so
           hornSchunck.cpp itself cannot be compiled) on a bounded sample of the same workload.
N > 1 (torchrun, one rank per GPU): the default workloads give every rank its own frame pair
(independent pairs, no communication, weak scaling); `batch256` deals the 256 pairs of BASELINE
config 4 to the ranks (strong scaling); `slab16k` is the row-slab decomposition of one 16384^2
pair with halo exchange (config 5, strong scaling).  `kitti` is config 1 on the reference's own
bundled frame pair.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (height, width, default iterations, BASELINE.json config it restates)
    "1080p": (1080, 1920, 1000, "configs[1]: synthetic textured 1920x1080 pair, alpha=1, 1000 Jacobi iterations"),
    "4k": (2160, 3840, 2000, "configs[2]: synthetic 3840x2160 pair, 2000 iterations"),
    "kitti": (375, 1242, 100, "configs[0]: the bundled 1242x375 pair 000050_10/11 (tests/golden fixture), w=5, 100 iterations"),
    "slab16k": (16384, 16384, 5000, "configs[4]: one 16384x16384 pair, 5000 iterations, row slabs + halo exchange"),
    "batch256": (1080, 1920, 500, "configs[3]: 256 independent 1080p pairs, 500 iterations each, split over the GPUs"),
}
ALGO_BYTES_PER_PIXEL_ITER = 32.0       # u,v read 8 + Ix,Iy,It,inv read 16 + u,v write 8 (fp32)
METRIC = "HS Mpixel-iter/s"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_lookup(workload, window, k, pixel_iterations_per_launch, launch_seconds, peak):
    """DRAM bytes of ONE launch of the dominant kernel from the committed ncu captures (profiles/traffic.json,
    written by tools/traffic_merge.py: dram__bytes_read.sum + dram__bytes_write.sum per pixel-iteration for this
    (workload, window, k)), scaled to the pixel-iterations of the benched launch.  None - never a stale
    constant - when no capture exists for the configuration."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None, "no profiles/traffic.json", None
    e = table.get(f"{workload}_w{window}_k{k}")
    if not e:
        return None, f"no ncu capture for {workload} w={window} k={k}", None
    byts = e["dram_bytes_per_pixel_iteration"] * pixel_iterations_per_launch
    gbs = byts / launch_seconds / 1e9
    return (byts, f"profiles/traffic.json[{workload}_w{window}_k{k}]: ncu dram__bytes_read+write.sum of a "
                  f"{e['sweeps']}-sweep launch = {e['dram_bytes_per_pixel_iteration']:.3f} B per pixel-iteration, scaled",
            {"gbs": gbs, "frac_of_peak": gbs / peak, "algorithmic_over_dram_bytes": 32.0 / e["dram_bytes_per_pixel_iteration"]})


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.t0 = [], None, index, 0.0

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def mark(self):
        """Start of the window whose samples are reported (samples before it are dropped)."""
        self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if r[0] < self.t0:
                continue
            r = r[1:]
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU algorithm on the box's host cores
# ---------------------------------------------------------------------------------------------
REF_BINARY = os.path.join(ROOT, "oracle", "_ref", "hs_ref")   # built by `make -C oracle ref` iff OpenCV C++ exists


def reference_solve_seconds(prev, nxt, window, iters, alpha, threads=0):
    """Wall seconds of ONE getFlow of the reference's CPU path on these frames, and what ran:
    ("reference", threads, description) when the unmodified hornSchunck.cpp could be compiled here
    (oracle/_ref/hs_ref), else ("port", ...) = the line-by-line cv2 restatement."""
    if os.path.exists(REF_BINARY):
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            pa, pb = os.path.join(d, "a.raw"), os.path.join(d, "b.raw")
            np.ascontiguousarray(prev).tofile(pa); np.ascontiguousarray(nxt).tofile(pb)
            out = subprocess.run([REF_BINARY, pa, pb, str(prev.shape[0]), str(prev.shape[1]), str(window), str(iters),
                                  repr(float(alpha))], capture_output=True, text=True, check=True).stdout
        fields = dict(kv.split("=") for kv in out.split())
        return float(fields["seconds"]), "reference", int(fields["threads"]), "HornSchunckOF/hornSchunck.cpp compiled with g++ -O3 (oracle/_ref/hs_ref)"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cv2
    import hs_oracle
    cv2.setNumThreads(int(threads) if threads else -1)        # 1 = the single-threaded reference of configs[0]; -1 = default
    t0 = time.perf_counter()
    hs_oracle.cv_flow(prev, nxt, window, iters, alpha)
    return (time.perf_counter() - t0, "port", cv2.getNumThreads(),
            f"oracle/hs_oracle.py::cv_flow = hornSchunck.cpp:19-75 through cv2 {cv2.__version__} "
            "(C++ OpenCV absent: hornSchunck.cpp not compilable here)")


def cpu_reference_rate(prev, nxt, window, alpha, seconds_budget, threads=0):
    """Time the reference's CPU path on a bounded number of sweeps of this workload."""
    t2, kind, nthr, what = reference_solve_seconds(prev, nxt, window, 2, alpha, threads)   # incl. the gradient stage
    per_iter = max(t2 / 2.0, 1e-4)
    iters = int(min(max(seconds_budget / per_iter, 4), 400))
    dt, kind, nthr, what = reference_solve_seconds(prev, nxt, window, iters, alpha, threads)
    rate = prev.shape[0] * prev.shape[1] * iters / dt / 1e6
    return rate, nthr, iters, dt, kind, what


def run_reference(args, rank, world):
    if rank != 0:
        return
    from cpp_optical_flow_b200 import synth
    H, W, T_default, cfgname = WORKLOADS[args.workload]
    if args.workload == "slab16k":       # bounded sample: a 16384 x 512 strip of the 16K^2 frame
        H = 512
    prev, nxt = synth.frame_pair(H, W)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cv2
    import hs_oracle
    if args.workload == "kitti":
        g = os.path.join(ROOT, "tests", "golden")
        prev = cv2.imread(os.path.join(g, "kitti_000050_10_gray.png"), cv2.IMREAD_UNCHANGED)
        nxt = cv2.imread(os.path.join(g, "kitti_000050_11_gray.png"), cv2.IMREAD_UNCHANGED)
    iters = args.ref_iters
    for _ in range(args.warmup):
        reference_solve_seconds(prev, nxt, args.window, max(1, iters // 4), 1.0)
    dt, kind, threads, what = 0.0, "port", 1, ""
    for _ in range(args.steps):
        t, kind, threads, what = reference_solve_seconds(prev, nxt, args.window, iters, 1.0)
        dt += t
    val = H * W * iters * args.steps / dt / 1e6
    sample = f"{W}x{H} frame pair, {iters} of {args.iters or T_default} sweeps per step (throughput is per sweep)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixel-iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "bundled frame pair of the reference" if args.workload == "kitti" else "synthetic",
            "config": {"workload": f"{args.workload}: {cfgname}", "window": args.window, "alpha": 1.0,
                       "iterations": args.iters or T_default},
            "cpu_baseline": {"value": val, "unit": "Mpixel-iter/s", "cores": threads, "kind": kind,
                             "sample": sample, "what": what},
            "e2e": {"value": val, "unit": "Mpixel-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import cpp_optical_flow_b200 as pkg
    from cpp_optical_flow_b200 import hs_ctypes as HC, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload == "slab16k":
        from cpp_optical_flow_b200 import slab
        with ClockSampler(local_rank) as clocks:
            rec = slab.bench_slab_record(args, rank, local_rank, world, ALGO_BYTES_PER_PIXEL_ITER, measured_hbm_peak,
                                         iterations=args.iters or None, steps=args.steps, exchange=args.slab_exchange,
                                         clock_sampler=clocks)
            time.sleep(0.05)
        check = slab.slab_bit_identity_check(rank, local_rank=local_rank, world=world) if not args.no_slab_check else None
        if rank == 0:
            line = {"metric": METRIC, "value": rec["value"], "unit": "Mpixel-iter/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": "slab16k: " + rec["workload"], "window": rec["window"], "alpha": 1.0,
                               "iterations": rec["iterations"], "temporal_k": rec["temporal_k"], "parallelism": rec["exchange"],
                               "halo_rows_per_exchange": rec["halo_rows_per_exchange"],
                               "halo_bytes_per_exchange_per_seam": rec["halo_bytes_per_exchange_per_seam"],
                               "exchanges_per_step": rec["exchanges_per_step"], "l2": rec["l2"]},
                    "roofline": dict(rec["roofline"], traffic=None), "e2e": None, "gpu_launches": rec["gpu_launches"],
                    "clocks": clocks.summary(), "slab_bit_identical": check}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    if args.workload == "batch256":
        return run_batch256(args, rank, local_rank, world)

    H, W, T_default, cfgname = WORKLOADS[args.workload]
    T = args.iters or T_default
    window = args.window
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    data = "synthetic"
    if args.workload == "kitti":
        import cv2
        g = os.path.join(ROOT, "tests", "golden")
        prev = cv2.imread(os.path.join(g, "kitti_000050_10_gray.png"), cv2.IMREAD_UNCHANGED)
        nxt = cv2.imread(os.path.join(g, "kitti_000050_11_gray.png"), cv2.IMREAD_UNCHANGED)
        data = "bundled frame pair of the reference (HornSchunckOF/img/leftimage/000050_10/11.png, gray)"
    else:
        prev, nxt = synth.video_pair(rank, H, W) if world > 1 else synth.frame_pair(H, W)

    solver = pkg.Solver(W, H, window, T, 1.0, device=local_rank, temporal_k=args.k, stream=stream.cuda_stream,
                        flags=HC.FLAG_TEXTBOOK if args.textbook else 0)
    solver.upload(prev, nxt)
    solver.sync()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()

    with torch.cuda.stream(stream), ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            solver.solve_device(); solver.sync()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        iter_ms, prep_ms, launches = [], [], 0
        barrier(); torch.cuda.synchronize()
        clocks.mark()
        wall0 = time.perf_counter()
        for s, e in ev:
            flush.zero_()                      # cold L2 at the start of every step
            s.record(stream)
            solver.solve_device()
            e.record(stream)
            solver.sync()
            t = solver.timing()
            iter_ms.append(t.iterate_ms); prep_ms.append(t.prepare_ms); launches += t.launches
        torch.cuda.synchronize(); barrier()
        wall = time.perf_counter() - wall0
        time.sleep(0.05)                       # let the last nvidia-smi sample of the window arrive
        step_ms = [s.elapsed_time(e) for s, e in ev]
    tm = solver.timing()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    work = float(H) * W * T * args.steps * world            # pixel-iterations, all ranks
    value = work / total_s / 1e6

    # ---- end to end through hs_solve with pinned host buffers -------------------------------
    hp = torch.from_numpy(prev).pin_memory(); hn = torch.from_numpy(nxt).pin_memory()
    hu = torch.empty((H, W), dtype=torch.float64).pin_memory(); hv = torch.empty((H, W), dtype=torch.float64).pin_memory()
    lib = HC.load_library()

    def solve_host():
        rc = lib.hs_solve(solver._ctx, hp.data_ptr(), W, 0, hn.data_ptr(), W, 0, hu.data_ptr(), W * 8, 0,
                          hv.data_ptr(), W * 8, 0, HC.HS_F64)
        if rc:
            raise RuntimeError(lib.hs_last_error(solver._ctx))

    e2e_steps = max(3, min(args.steps, 10))

    def time_calls(fn):
        """Wall time of e2e_steps synchronous calls of `fn` (host buffers in, host buffers out); the L2 flush
        between steps is outside the timed region.  Max over ranks -> whole-job throughput."""
        fn()
        barrier(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(e2e_steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            stream.synchronize()
            t0 = time.perf_counter()
            fn()
            tot += time.perf_counter() - t0
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(H) * W * T * e2e_steps * world / float(t.item()) / 1e6

    e2e_value = time_calls(solve_host)
    et = solver.timing()
    # the same call with float32 outputs (half the D2H bytes; what a caller that does not need CV_64FC1 gets)
    hu32 = torch.empty((H, W), dtype=torch.float32).pin_memory(); hv32 = torch.empty((H, W), dtype=torch.float32).pin_memory()

    def solve_host_f32():
        rc = lib.hs_solve(solver._ctx, hp.data_ptr(), W, 0, hn.data_ptr(), W, 0, hu32.data_ptr(), W * 4, 0,
                          hv32.data_ptr(), W * 4, 0, HC.HS_F32)
        if rc:
            raise RuntimeError(lib.hs_last_error(solver._ctx))
    e2e_f32 = time_calls(solve_host_f32)
    # ... and from PAGEABLE host memory, which is what a cv::Mat hands the adapter unless it registers it
    pu = np.empty((H, W), np.float64); pv = np.empty((H, W), np.float64)
    pp, pn = np.array(prev), np.array(nxt)

    def solve_pageable():
        rc = lib.hs_solve(solver._ctx, pp.ctypes.data, W, 0, pn.ctypes.data, W, 0, pu.ctypes.data, W * 8, 0,
                          pv.ctypes.data, W * 8, 0, HC.HS_F64)
        if rc:
            raise RuntimeError(lib.hs_last_error(solver._ctx))
    e2e_pageable = time_calls(solve_pageable)

    # ---- the same through the streaming front-end (frame sequences: each frame uploaded once,
    #      H2D / solve / D2H of consecutive pairs overlapped) ------------------------------------
    import ctypes
    idx = ctypes.c_int(-1)
    for src in (hp, hn, hp):                                  # warm-up: lazy allocations of the front-end
        lib.hs_video_push(solver._ctx, src.data_ptr(), W, hu.data_ptr(), W * 8, hv.data_ptr(), W * 8, HC.HS_F64,
                          ctypes.byref(idx))
    lib.hs_video_flush(solver._ctx, hu.data_ptr(), W * 8, hv.data_ptr(), W * 8, ctypes.byref(idx))
    lib.hs_video_reset(solver._ctx)
    n_frames = e2e_steps + 1
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_frames):
        src = hp if i % 2 == 0 else hn
        rc = lib.hs_video_push(solver._ctx, src.data_ptr(), W, hu.data_ptr(), W * 8, hv.data_ptr(), W * 8, HC.HS_F64,
                               ctypes.byref(idx))
        if rc:
            raise RuntimeError(lib.hs_last_error(solver._ctx))
    lib.hs_video_flush(solver._ctx, hu.data_ptr(), W * 8, hv.data_ptr(), W * 8, ctypes.byref(idx))
    torch.cuda.synchronize()
    st = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
    stream_value = float(H) * W * T * (n_frames - 1) * world / float(st.item()) / 1e6

    # ---- independent pairs with TWO contexts in flight: pair i + 1 is queued (hs_upload, hs_solve_device,
    #      hs_download: all asynchronous on the context's stream) before pair i is waited for with hs_sync ----
    solver2 = pkg.Solver(W, H, window, T, 1.0, device=local_rank, temporal_k=args.k, stream=stream.cuda_stream,
                         flags=HC.FLAG_TEXTBOOK if args.textbook else 0)
    hu2 = torch.empty((H, W), dtype=torch.float64).pin_memory(); hv2 = torch.empty((H, W), dtype=torch.float64).pin_memory()
    ctxs = [(solver, hu, hv), (solver2, hu2, hv2)]

    def enqueue(i):
        sv, ou, ov = ctxs[i & 1]
        sv.solve_async_raw(hp.data_ptr(), hn.data_ptr(), W, 0, ou.data_ptr(), ov.data_ptr(), W * 8, 0, HC.HS_F64)

    for i in range(2):
        enqueue(i)
    solver.solve_wait(); solver2.solve_wait()
    n_pairs = 2 * e2e_steps
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    enqueue(0)
    for i in range(n_pairs):
        if i + 1 < n_pairs:
            enqueue(i + 1)
        ctxs[i & 1][0].solve_wait()
    pt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    pipelined_value = float(H) * W * T * n_pairs * world / float(pt.item()) / 1e6
    solver2.close()

    solver.close()
    del flush, hp, hn, hu, hv
    torch.cuda.empty_cache()

    # ---- BASELINE configs[4] beside the headline: ONE 16384^2 pair in row slabs over the same N GPUs
    #      (the only path with a real exchange step), plus a bit-identity check of that path --------
    slab_rec = None
    if args.workload == "1080p" and not args.no_slab and not args.textbook:
        from cpp_optical_flow_b200 import slab
        with ClockSampler(local_rank) as sclocks:
            slab_rec = slab.bench_slab_record(args, rank, local_rank, world, ALGO_BYTES_PER_PIXEL_ITER, measured_hbm_peak,
                                              steps=args.slab_steps, exchange=args.slab_exchange, clock_sampler=sclocks)
            time.sleep(0.05)
        check = slab.slab_bit_identity_check(rank, world, local_rank)
        single = None
        if world > 1:        # the same code on ONE GPU (rank 0), for the parallel efficiency of this very run
            if rank == 0:
                single = slab.bench_slab_record(args, 0, local_rank, 1, ALGO_BYTES_PER_PIXEL_ITER, measured_hbm_peak, steps=1)
            dist.barrier()
        if rank == 0:
            slab_rec["clocks"] = sclocks.summary()
            slab_rec["slab_bit_identical"] = check
            v1 = single["value"] if single else slab_rec["value"]
            slab_rec["single_gpu_value"] = v1
            slab_rec["slab_parallel_efficiency"] = slab_rec["value"] / (world * v1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the Jacobi sweeps) -----------------------------------
    peak, peak_src = measured_hbm_peak()
    it_s = float(np.mean(iter_ms)) / 1e3
    achieved = ALGO_BYTES_PER_PIXEL_ITER * H * W * T / it_s / 1e9
    sweep_launches = launches // args.steps - 1              # minus the gradient kernel
    traffic, traffic_src, dram = traffic_lookup(args.workload, window, tm.temporal_k, float(H) * W * T / max(sweep_launches, 1),
                                                it_s / max(sweep_launches, 1), peak)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "dram": dram, "peak_source": peak_src,
                "kernel": "k_jacobi_tile" if tm.kernel_id == 1 else "k_jacobi_generic",
                "launches_per_step": sweep_launches, "avg_launch_us": it_s / max(sweep_launches, 1) * 1e6,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PIXEL_ITER * H * W * T / max(sweep_launches, 1),
                "fused_sweeps_per_launch": tm.temporal_k,
                "note": "32 B per pixel-iteration, no credit for temporal blocking: k fused sweeps per HBM round "
                        "trip is why frac can exceed 1"}

    # The unit K3 is really bound by ("What bounds K3"): the FP32 pipe.  Lane-operations per STAGED
    # pixel-sweep of the canonical arithmetic, the valid fraction of a staged 128 x 48 tile at this k, and the
    # SM clock sampled under load give the pipe's roof for this configuration.
    clk = clocks.summary()
    if tm.kernel_id == 1 and not args.textbook and clk and clk.get("sm_mhz"):
        a = window - window // 2 - 1
        rl, rr, k = a, window - 1 - a, tm.temporal_k
        adds = window / 2.0                               # adds of a paired window sum per pixel and direction (w=3: 1.5, w=5: 2.5)
        lane_ops = 2 * 2 * adds + 2 + 5                   # u and v, two directions; mean scaling; the update
        vx = 128 - ((rl * k + 3) // 4) * 4 - ((rr * k + 3) // 4) * 4
        hyt = rl * k + ((rl * k) & 1)
        vy = (48 - hyt - rr * k) & ~1
        valid = max(vx, 0) * max(vy, 0) / (128.0 * 48.0)
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        roof = sms * 128 * clk["sm_mhz"] * 1e6 / lane_ops * valid / 1e6          # Mpixel-iter/s
        roofline["fp32_pipe"] = {"lane_ops_per_staged_pixel_sweep": lane_ops, "valid_fraction_of_a_staged_tile": valid,
                                 "roof_mpixel_iter_s": roof, "frac": (value / world) / roof,
                                 "note": "128 FP32 lanes per SM and clock at the sampled SM clock; FADD2/FMUL2/FFMA2 occupy "
                                         "the pipe for two cycles (tools/fp_pipe_probe.cu), so packing does not raise this roof"}

    cpu = None
    if world == 1 and not args.no_cpu and not args.textbook:
        # headline: ONE thread, as BASELINE configs[0] / BASELINE.md 4.3 specify; all host cores beside it
        rate, threads, its, dt, kind, what = cpu_reference_rate(prev, nxt, window, 1.0, args.cpu_seconds, threads=1)
        rate_all, threads_all, its_all, dt_all, _, _ = cpu_reference_rate(prev, nxt, window, 1.0, args.cpu_seconds / 2)
        cpu = {"value": rate, "unit": "Mpixel-iter/s", "cores": threads, "kind": kind,
               "sample": f"same {W}x{H} pair, {its} of {T} sweeps ({dt:.1f} s)", "what": what,
               "all_cores": {"value": rate_all, "cores": threads_all,
                             "sample": f"same pair, {its_all} of {T} sweeps ({dt_all:.1f} s)"}}

    line = {"metric": METRIC, "value": value, "unit": "Mpixel-iter/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": data,
            "config": {"workload": f"{args.workload}: {cfgname}", "window": window, "alpha": 1.0, "iterations": T,
                       "pairs_per_rank_per_step": 1, "parallelism": f"independent pairs x{world}" if world > 1 else "1 gpu",
                       "l2": "512 MiB memset before every step (cold L2 at step start; events exclude it)",
                       "temporal_k": tm.temporal_k,
                       **({"mode": "textbook Horn-Schunck (HS_FLAG_TEXTBOOK): NOT the reference's arithmetic"} if args.textbook else {})},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "Mpixel-iter/s", "h2d_bytes_per_step": 2 * H * W,
                    "d2h_bytes_per_step": 2 * H * W * 8, "steps": e2e_steps,
                    "note": "hs_solve: pinned host uint8 frames in, pinned float64 u, v out (the reference's CV_64FC1)",
                    "value_f32_outputs": e2e_f32, "value_pageable_buffers": e2e_pageable,
                    "stream_value": stream_value,
                    "value_two_contexts_in_flight": pipelined_value,
                    "two_contexts_note": "independent pairs through hs_solve_async / hs_solve_wait on two contexts that share one "
                                         "compute stream: the copies of one pair overlap the sweeps of the other (L2 not flushed between pairs)",
                    "stream_note": "hs_video_push: consecutive pairs of a frame sequence, one frame uploaded per pair, "
                                   "H2D/solve/D2H overlapped (L2 not flushed between pairs)",
                    "last_step_ms": {"h2d": et.h2d_ms, "prepare": et.prepare_ms, "iterate": et.iterate_ms,
                                     "d2h": et.d2h_ms, "total": et.total_ms}},
            "gpu_launches": launches, "clocks": clocks.summary(),
            "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
            "prepare_ms": float(np.mean(prep_ms)), "iterate_ms": float(np.mean(iter_ms))}
    if cpu:
        line["cpu_baseline"] = cpu
    if slab_rec:
        line["slab16k"] = slab_rec
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_batch256(args, rank, local_rank, world):
    """BASELINE config 4: 256 independent 1080p pairs (synthetic video), 500 sweeps each, pairs dealt
    to the ranks, no communication.  Total work is fixed -> strong scaling.  `value`: frames resident
    in HBM; `e2e`: every pair comes from / goes back to pinned host memory through hs_solve."""
    import torch
    import torch.distributed as dist
    import cpp_optical_flow_b200 as pkg
    from cpp_optical_flow_b200 import hs_ctypes as HC, synth
    H, W, T_default, cfgname = WORKLOADS["batch256"]
    T = args.iters or T_default
    pairs_total, B = 256, 4                                   # B pairs per hs_solve call (one launch)
    if pairs_total % (world * B):
        raise SystemExit("batch256 needs a GPU count that divides 64")
    calls = pairs_total // world // B
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    # B distinct pairs per rank, reused for every call of the rank (generating 256 distinct 1080p
    # pairs on the host would take minutes and change nothing for the device)
    pairs = [synth.video_pair(rank * B + j, H, W) for j in range(B)]
    hp = torch.from_numpy(np.stack([p[0] for p in pairs])).pin_memory()
    hn = torch.from_numpy(np.stack([p[1] for p in pairs])).pin_memory()
    hu = torch.empty((B, H, W), dtype=torch.float64).pin_memory()
    hv = torch.empty((B, H, W), dtype=torch.float64).pin_memory()
    solver = pkg.Solver(W, H, args.window, T, 1.0, batch=B, device=local_rank, temporal_k=args.k,
                        stream=stream.cuda_stream)
    lib = HC.load_library()

    def barrier():
        if world > 1:
            dist.barrier()

    def solve_host():
        rc = lib.hs_solve(solver._ctx, hp.data_ptr(), W, H * W, hn.data_ptr(), W, H * W, hu.data_ptr(), W * 8,
                          H * W * 8, hv.data_ptr(), W * 8, H * W * 8, HC.HS_F64)
        if rc:
            raise RuntimeError(lib.hs_last_error(solver._ctx))

    solver.upload(hp.numpy(), hn.numpy()); solver.sync()
    for _ in range(args.warmup):
        solver.solve_device(); solver.sync(); solve_host()
    launches = 0
    with torch.cuda.stream(stream), ClockSampler(local_rank) as clocks:
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); torch.cuda.synchronize(); clocks.mark()
        s_ev.record(stream)
        for _ in range(args.steps * calls):
            solver.solve_device()
            launches += 2
        e_ev.record(stream)
        torch.cuda.synchronize(); barrier()
        dev_ms = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=dev)
        t0 = time.perf_counter()
        for _ in range(args.steps * calls):
            solve_host()
        torch.cuda.synchronize()
        serial_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        # The same calls with TWO contexts in flight (what a caller with many independent pairs does): call i + 1 is
        # queued - upload, solve, download, all asynchronous on its context's stream - before call i is waited
        # for, so the copies of one batch run under the sweeps of the other.  Every call still moves its own
        # frames in and its own float64 flow out inside the timed region.
        solver2 = pkg.Solver(W, H, args.window, T, 1.0, batch=B, device=local_rank, temporal_k=args.k, stream=stream.cuda_stream)
        hu2 = torch.empty((B, H, W), dtype=torch.float64).pin_memory()
        hv2 = torch.empty((B, H, W), dtype=torch.float64).pin_memory()
        ctxs = [(solver, hu, hv), (solver2, hu2, hv2)]

        def enqueue(i):
            sv, ou, ov = ctxs[i & 1]
            sv.solve_async_raw(hp.data_ptr(), hn.data_ptr(), W, H * W, ou.data_ptr(), ov.data_ptr(), W * 8, H * W * 8, HC.HS_F64)

        for i in range(2):                                    # warm the second context
            enqueue(i)
        solver.solve_wait(); solver2.solve_wait()
        barrier(); torch.cuda.synchronize()
        ncalls = args.steps * calls
        t0 = time.perf_counter()
        enqueue(0)
        for i in range(ncalls):
            if i + 1 < ncalls:
                enqueue(i + 1)
            ctxs[i & 1][0].solve_wait()                       # call i is complete: its flow is in host memory
        host_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        solver2.close()
        time.sleep(0.05)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX); dist.all_reduce(host_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(serial_s, op=dist.ReduceOp.MAX)
    work = float(pairs_total) * H * W * T * args.steps
    if rank == 0:
        peak, src = measured_hbm_peak()
        value = work / (float(dev_ms.item()) / 1e3) / 1e6
        achieved = ALGO_BYTES_PER_PIXEL_ITER * work / world / (float(dev_ms.item()) / 1e3) / 1e9
        tk = solver.timing().temporal_k
        launch_s = float(dev_ms.item()) / 1e3 / (args.steps * calls)
        traffic, traffic_src, dram = traffic_lookup("batch256", args.window, tk, float(B) * H * W * T, launch_s, peak)
        line = {"metric": METRIC, "value": value, "unit": "Mpixel-iter/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(dev_ms.item()) / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"batch256: {cfgname}", "window": args.window, "alpha": 1.0, "iterations": T,
                           "pairs_per_call": B, "calls_per_rank_per_step": calls,
                           "parallelism": f"256 pairs over {world} GPU(s), no communication",
                           "l2": "4 pairs per launch = 200 MB working set per GPU, larger than L2"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_source": traffic_src, "dram": dram,
                             "peak_source": src, "kernel": "k_jacobi_tile", "note": "per GPU; one launch = one call of 4 pairs"},
                "e2e": {"value": work / float(host_s.item()) / 1e6, "unit": "Mpixel-iter/s",
                        "h2d_bytes_per_step": 2 * H * W * pairs_total // world,
                        "d2h_bytes_per_step": 2 * H * W * 8 * pairs_total // world,
                        "note": "hs_solve_async / hs_solve_wait per call of 4 pairs, pinned host buffers, float64 flow out; two "
                                "contexts on one compute stream in flight, so the copies of one call overlap the sweeps of the other",
                        "value_serial_calls": work / float(serial_s.item()) / 1e6},
                "gpu_launches": launches, "clocks": clocks.summary()}
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--window", type=int, default=0,
                    help="windowSize; default 3 (the north star's 3x3), 5 for --workload kitti (main.cpp:94)")
    ap.add_argument("--iters", type=int, default=0, help="override the workload's iteration count")
    ap.add_argument("--k", type=int, default=0, help="fused sweeps per launch (0 = library default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-iters", type=int, default=16, help="--impl reference: sweeps per step (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="skip the 16K^2 row-slab sub-record of the default run")
    ap.add_argument("--slab-steps", type=int, default=2, help="timed 16K^2 solves of the sub-record")
    ap.add_argument("--slab-exchange", default="peer", choices=["peer", "nccl"],
                    help="row slabs: in-kernel exchange over peer memory (default) or NCCL send/recv between launches")
    ap.add_argument("--no-slab-check", action="store_true")
    ap.add_argument("--textbook", action="store_true",
                    help="non-parity extra: cube gradients + weighted 3x3 average (HS_FLAG_TEXTBOOK); no CPU baseline")
    args = ap.parse_args()
    if args.window <= 0:
        args.window = 5 if args.workload == "kitti" else 3
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun for N>1", file=sys.stderr)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
