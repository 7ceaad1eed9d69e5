#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 tape-replay engine.

Metric (BASELINE.json): compiled-tape ``gradient!`` evals/s of ``f(a,b) = sum(a'*b + a*b')`` at 4096x4096 FP64
(configs[1]).  A "step" is one gradient! replay (forward sweep + seeded reverse sweep + result extraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 4096] [--dtype f64|f32]

* ``value``     device-resident inputs/results (D2D copies only), CUDA events, max over ranks.
* ``e2e``       page-locked HOST inputs and HOST results, every step uploads its inputs and downloads its results; consecutive steps
                pipelined over four engine copies of the tape (rd.GradientPipeline = rdb_gradient_async + rdb_tape_wait); the
                one-synchronous-call-per-step number rides along (``e2e.synchronous_calls``), and ``e2e.host_link`` is the measured
                host<->device copy bandwidth with every rank copying (the platform ceiling of this leg).
* ``roofline``  dominant kernel (the tcgen05 int8-slice GEMM behind every FP64 ``*``): tensor-pipe operations issued per DENSE launch
                (28 int8 products of 2 n^3: the two forward products) / CUDA-event time of those launches including the slicing
                passes of both operands, against 2 x the driver-measured bf16 burst (MEASURED_PEAKS.json).  The four adjoint
                products meet deriv(output) of `sum` - a constant fill with one live digit plane - and issue 7 slice pairs
                (counted on the device, rdb_tape_slice_pairs); they are bound by operand delivery and reported under
                ``roofline.launch_kinds``; ``other_configs.cfg2_without_plane_trimming`` is the same leg with every plane
                multiplied.  FP64-equivalent TFLOP/s next to a cuBLAS DGEMM measured in the same run; ``traffic`` from the committed
                ncu capture named in profiles/traffic.json.
* ``cpu_baseline`` the oracle (reference-equivalent CPU replay, numpy/OpenBLAS) on the box's host cores.
* N > 1: gradient! of one scalar function does not shard ("replicas only"): every rank replays its
  own tape copy, value = N evals / max-over-ranks time.  The same run then measures the paths that DO shard -
  jacobian! by output-seed rows and hessian! by column chunks (results left sharded; the product's NCCL gather timed apart) -
  and reports them under "jacobian"/"hessian"; N = 1 also runs cfg 1, cfg 2 in FP32 and cfg 3 ("other_configs").
* ``--impl reference``: the reference-equivalent CPU replay alone (rank 0 only), same metric/config.
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
import __graft_entry__ as graft  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--jac-n", type=int, default=8192)
    ap.add_argument("--hess-n", type=int, default=16384)
    ap.add_argument("--hess-large-n", type=int, default=65536,
                    help="second hessian! leg: at n=16384 one B200 writes the whole 2.1 GB Hessian in ~0.45 ms, too little work to "
                         "shard 8 ways; n=65536 (34 GB) shows the scaling of the same code path (0 = skip)")
    ap.add_argument("--gemm", default="auto", choices=["auto", "dmma"],
                    help="FP64 `*` kernels: auto = exact int8 slices on tcgen05 where eligible (default), dmma = DMMA only")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extra", action="store_true", help="skip the jacobian!/hessian! legs")
    ap.add_argument("--only-value", action="store_true", help="device-resident leg only: print {value, ms_per_step, parity} and exit")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def pause(self):
        """stop sampling (the lines so far are kept); start() resumes"""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            try:
                self.thread.join(timeout=2)
            except Exception:
                pass
            self.proc = None
            self.ran = True

    def stop(self):
        self.pause()
        if not getattr(self, "ran", False):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def timed_steps(torch, dist, world, step, steps, warmup):
    """W untimed warm-ups, then exactly K steps between barrier+synchronize, CUDA events, max over ranks (ms)."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def cpu_replay(rd, orc, n, npdt, evals):
    """Reference-equivalent CPU replay (oracle): best-of wall time per eval."""
    rng = np.random.default_rng(1)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((8, 8)).astype(npdt), rng.random((8, 8)).astype(npdt)))
    _ = tape
    a0 = np.asfortranarray(rng.random((n, n)).astype(npdt)); b0 = np.asfortranarray(rng.random((n, n)).astype(npdt))
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0))
    ot = orc.OracleTape(tape.program_bytes())
    a, b = rd.workloads.inputs_gradient_example(n, npdt)
    times = []
    for _ in range(evals):
        t0 = time.perf_counter()
        ot.gradient([a, b])
        times.append(time.perf_counter() - t0)
    return times


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms set their BLAS pool to all usable cores themselves so that
    the reference number does not depend on how the script was launched.  Returns the thread count in effect."""
    try:
        ncores = len(os.sched_getaffinity(0))
    except Exception:
        ncores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncores)
    except Exception:
        pass
    return blas_threads()


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE any page-locked buffer is
    allocated, so that first touch places the staging memory on the GPU's own NUMA node.  Returns what was found."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        info["numa_node"] = int(open(path + "/numa_node").read().strip())
        cl = open(path + "/local_cpulist").read().strip()
        cpus = set()
        for part in cl.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        info["cpus"] = cl
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as e:
        info["error"] = str(e)[:80]
    return info


def pcie_probe(torch, dist, world, nbytes=256 << 20, reps=4):
    """Host<->device copy bandwidth of page-locked buffers with ALL ranks copying at once, both directions concurrently: the
    platform ceiling of the end-to-end leg (every evaluation moves 268 MB each way per GPU)."""
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def go():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    go(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    gbps = reps * nbytes / dt / 1e9
    return {"GBps_each_way_per_gpu_all_ranks_copying": gbps, "evals_per_s_bound_per_gpu": gbps * 1e9 / (2 * 4096 * 4096 * 8),
            "note": "concurrent H2D + D2H of 256 MiB page-locked buffers on every rank, max over ranks"}


def workload_config(n, dtype):
    """`config` is identical on both arms (the driver compares them); descriptive extras live in `config_detail`."""
    return {"workload": f"compiled-tape gradient! of sum(a'*b + a*b') at {n}x{n} {dtype} (BASELINE configs[1])", "n": n}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    rd = graft.load_package(); orc = graft.load_oracle()
    npdt = np.float64 if args.dtype == "f64" else np.float32
    n = args.n
    cores = use_all_host_cores()
    times = cpu_replay(rd, orc, n, npdt, args.warmup + args.steps)[args.warmup:]
    total = float(np.sum(times))
    val = args.steps / total
    line = {
        "impl": "reference", "metric": "gradient! evals/s", "value": val, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(n, args.dtype),
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full-size evals of the numpy/OpenBLAS oracle (reference-equivalent CPU replay; "
                                   f"Julia is not installed so ReverseDiff.jl itself cannot run), {cores} BLAS threads"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def committed_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
    (profiles/traffic.json names the capture file, the command and the commit it was taken at) - never a literal in this file."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return float(rec["dram_bytes_per_launch"]), f"{rec['source']} @ {rec['commit']}"
    except Exception as e:
        return None, f"profiles/traffic.json unavailable ({e})"


def measure_blas_peak(torch, n, tdt):
    a = torch.rand(n, n, dtype=tdt, device="cuda"); b = torch.rand(n, n, dtype=tdt, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    best = 1e30
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def measure_int8_peak(torch, n=8192):
    """cuBLASLt int8 GEMM (torch._int_mm) as the measured denominator of the int8 tensor pipe; None if unavailable."""
    try:
        a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()   # column-major operand
        for _ in range(2):
            torch._int_mm(a, b)
        best = 1e30
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    rd = graft.load_package()
    ctx = rd.engine.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)     # engine launches on torch's stream: events see it

    npdt = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    es = np.dtype(npdt).itemsize
    n = args.n
    rng = np.random.default_rng(1)
    a0 = np.asfortranarray(rng.random((n, n)).astype(npdt)); b0 = np.asfortranarray(rng.random((n, n)).astype(npdt))
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0))
    mode = 1 if args.gemm == "dmma" else 0
    ct = rd.compile(tape, ctx=ctx, gemm_f64_mode=mode)
    a, b = rd.workloads.inputs_gradient_example(n, npdt)

    # ---- device-resident leg ---------------------------------------------------------------------
    a_d = rd.engine.device_array(a); b_d = rd.engine.device_array(b)    # column-major images of a, b
    ga_d = torch.empty(n * n, dtype=tdt, device="cuda"); gb_d = torch.empty(n * n, dtype=tdt, device="cuda")

    def step_dev():
        rd.gradient_((ga_d, gb_d), ct, (a_d, b_d))

    sampler = ClockSampler(local)
    l0 = ct.handle.launch_count()
    step_dev()   # first touch outside the timing
    l_per_step = ct.handle.launch_count() - l0
    if rank == 0:
        sampler.start()
    ms = timed_steps(torch, dist, world, step_dev, args.steps, args.warmup)
    if rank == 0:
        sampler.pause()                                       # the host-side parity check below is not part of any timed region
    value = world * args.steps / (ms * 1e-3)

    # parity of what was just timed: closed form, O(n^2)
    ga = ga_d.cpu().numpy().reshape(n, n, order="F").astype(np.float64)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    err = np.max(np.abs(ga - (b64.sum(1)[:, None] + b64.sum(0)[None, :]))) / np.max(np.abs(ga))
    tol = 1e-12 if args.dtype == "f64" else 1e-5
    if not err < tol:
        raise SystemExit(f"parity check failed: {err}")

    if args.only_value:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms / args.steps, "max_rel_err_vs_closed_form": float(err),
                              "gemm_guard": ct.handle.guard_stats()}), flush=True)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- end-to-end leg: pinned host inputs and results --------------------------------------------
    def pinned(arr):
        t = torch.empty(arr.size, dtype=tdt).pin_memory()
        h = t.numpy().reshape(arr.shape, order="F")
        h[...] = arr
        return t, h
    keep = []
    (ta_, a_h), (tb_, b_h) = pinned(a), pinned(b)
    (tga, ga_h), (tgb, gb_h) = pinned(np.zeros((n, n), npdt)), pinned(np.zeros((n, n), npdt))
    keep += [ta_, tb_, tga, tgb]
    res_h = (rd.DiffResult(0.0, ga_h), rd.DiffResult(0.0, gb_h))

    def step_e2e():
        rd.gradient_(res_h, ct, (a_h, b_h))

    # (a) one synchronous call per step: upload, sweeps and download of one evaluation, nothing overlapping the next one
    ms_e2e_sync = timed_steps(torch, dist, world, step_e2e, max(3, args.steps // 2), 3)
    e2e_sync_steps = max(3, args.steps // 2)
    # (b) the headline end-to-end number: the same per-step work (every step uploads its inputs from page-locked host arrays and
    # downloads its two gradients and the value), consecutive evaluations pipelined over three engine copies of the tape
    # (rd.GradientPipeline = rdb_tape_set_input + rdb_gradient_async, rdb_tape_wait): evaluation k+1 uploads and k-1 downloads
    # while k computes.  The timed region ends when the LAST result byte is in host memory (drain).
    (tga2, ga_h2), (tgb2, gb_h2) = pinned(np.zeros((n, n), npdt)), pinned(np.zeros((n, n), npdt))
    keep += [tga2, tgb2]
    res_h2 = (rd.DiffResult(0.0, ga_h2), rd.DiffResult(0.0, gb_h2))
    # depth 4 (measured, profiles/e2e_sweep.sh: 126.7 / 137.7 / 144.2 evals/s at depth 2 / 3 / 4): a slot's own upload -> sweeps ->
    # download chain takes ~17 ms while every engine - copy in, copy out, SMs - is busy only ~6 ms per evaluation
    depth = int(os.environ.get("RDB_BENCH_DEPTH", "4"))
    res_slots = [res_h, res_h2]
    host_results = [(ga_h, gb_h), (ga_h2, gb_h2)]
    for _ in range(depth - 2):
        (tgx, ga_hx), (tgy, gb_hx) = pinned(np.zeros((n, n), npdt)), pinned(np.zeros((n, n), npdt))
        keep += [tgx, tgy]
        res_slots.append((rd.DiffResult(0.0, ga_hx), rd.DiffResult(0.0, gb_hx)))
        host_results.append((ga_hx, gb_hx))
    res_slots = res_slots[:depth]; host_results = host_results[:depth]
    pipe = rd.GradientPipeline(tape, depth=depth, device=local, gemm_f64_mode=mode)
    e2e_steps = max(40, args.steps)                           # (the first upload and the last download overlap nothing: amortised over the run)

    def run_pipeline(k):
        for j in range(k):
            pipe.submit(res_slots[j % depth], (a_h, b_h))
        pipe.drain()
    run_pipeline(2 * depth)                                        # warm-up: all slots compile their kernels and size their workspaces
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for ga_x, gb_x in host_results:
        ga_x[...] = 0; gb_x[...] = 0
    lp0 = pipe.launch_count()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    run_pipeline(e2e_steps)                                   # ends with every result in host memory
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3                 # host clock around a region that begins and ends with an idle device
    e2e_launches = pipe.launch_count() - lp0
    if world > 1:
        tt = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
        dist.barrier()
    e2e_val = world * e2e_steps / (ms_e2e * 1e-3)
    pcie = pcie_probe(torch, dist, world)
    clocks = sampler.stop() if rank == 0 else None           # sampled every 20 ms during the timed regions (device-resident, end-to-end)
    gb_ref = a64.sum(1)[:, None] + a64.sum(0)[None, :]
    for ga_x, gb_x in host_results:
        assert np.max(np.abs(ga_x - ga)) <= tol * np.max(np.abs(ga))
        assert np.max(np.abs(gb_x - gb_ref)) <= tol * np.max(np.abs(gb_ref))
    del pipe

    # ---- roofline of the dominant kernel: per-launch CUDA events inside the engine --------------------
    def gemm_stats(m):
        ctp = rd.compile(tape, ctx=ctx, profile=1, gemm_f64_mode=m)
        for _ in range(2):
            rd.gradient_((ga_d, gb_d), ctp, (a_d, b_d))
        ctp.handle.stats()
        reps_ = 3
        for _ in range(reps_):
            rd.gradient_((ga_d, gb_d), ctp, (a_d, b_d))
        st = json.loads(ctp.handle.stats())
        gs = ctp.handle.guard_stats()
        del ctp
        return st, reps_, gs
    stats, reps, gstats = gemm_stats(mode)
    g_ms = sum(s["ms"] for s in stats if s["step"].startswith("gemm"))
    g_fl = sum(s["flops"] for s in stats if s["step"].startswith("gemm"))
    g_cnt = sum(s["count"] for s in stats if s["step"].startswith("gemm"))
    tot_ms = sum(s["ms"] for s in stats)
    fp_equiv = g_fl / (g_ms * 1e-3) / 1e12                    # FP64/FP32-equivalent TFLOP/s of the `*` instructions
    ew = [s for s in stats if not s["step"].startswith("gemm")]
    ew_bytes = sum(s["bytes"] for s in ew); ew_ms = sum(s["ms"] for s in ew)
    blas_peak = measure_blas_peak(torch, n, tdt)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    peak_tag = "MEASURED_PEAKS.json bf16_tflops" if "bf16_tflops" in peaks else "fallback 1590 TFLOP/s bf16"
    obits = 7 if os.environ.get("RDB_OZAKI_BITS", "8") == "7" else 8        # librdb200's default: 8-bit digits, 7 slices
    slices = int(os.environ.get("RDB_OZAKI_SLICES", "7" if obits == 8 else "8"))
    if args.dtype == "f64" and mode == 0:
        # tcgen05 kind::i8 path: every FP64 GEMM is S(S+1)/2 exact int8 slice products of the same shape (28 at S = 7 x 8 bits)
        ops_per_flop = slices * (slices + 1) / 2
        kernel = (f"oz2::umma_ozaki2c_kernel<{slices},2> (TMA multicast in 2-CTA clusters + tcgen05.mma kind::i8, int32 TMEM accumulators; "
                  f"{slices} slices of {obits} bits) + k_slice_i8/k_row_absmax")
        unit = "TOP/s (int8)"
        i8 = measure_int8_peak(torch)
        # denominator: the int8 pipe issues at twice the bf16 rate, so 2 x the driver-measured bf16 burst; the cuBLASLt int8 GEMM
        # measured in this run is reported next to it (this kernel's int8 rate exceeds it, so it cannot serve as the peak)
        pipe_peak = max(2.0 * bf16_peak, i8 or 0.0)
        peak_src = (f"2 x {peak_tag} (kind::i8 issues at twice the bf16 rate; nominal dense 4500); cuBLASLt int8 GEMM 8192^3 "
                    f"(torch._int_mm) measured in this run = {i8:.0f}" if i8 else
                    f"2 x {peak_tag} (kind::i8 issues at twice the bf16 rate; nominal dense 4500); torch._int_mm unavailable")
        traffic, traffic_src = committed_traffic("int8")
    elif args.dtype == "f64":
        ops_per_flop = 1.0
        kernel = "d64::dgemm_dmma_kernel (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4)"
        unit, pipe_peak = "TFLOP/s (fp64)", blas_peak
        peak_src = "cuBLAS DGEMM measured in this run (MEASURED_PEAKS.json holds no FP64 figure)"
        traffic, traffic_src = committed_traffic("dmma")
    else:
        ops_per_flop = 3.0
        kernel = "umma::umma_gemm_kernel<tf32> (TMA + tcgen05.mma kind::tf32, 3xTF32 split) + k_split_tf32"
        unit = "TFLOP/s (tf32)"
        try:
            old = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True
            tf32 = measure_blas_peak(torch, 8192, torch.float32)
            torch.backends.cuda.matmul.allow_tf32 = old
        except Exception:
            tf32 = None
        if tf32 is not None and tf32 > 2 * blas_peak:
            pipe_peak = tf32
            peak_src = f"cuBLAS TF32 GEMM 8192^3 (torch.matmul, allow_tf32) measured in this run; for context 0.5 x {peak_tag} = {0.5 * bf16_peak:.0f}"
        else:
            pipe_peak = 0.5 * bf16_peak
            peak_src = f"0.5 x {peak_tag} (tf32 issues at half the bf16 rate)"
        traffic, traffic_src = None, None
    achieved = fp_equiv * ops_per_flop
    launch_kinds = None
    if args.dtype == "f64" and mode == 0 and gstats.get("checked", 0) > 0:
        # Slice pairs actually issued (rdb_tape_slice_pairs): the kernels skip digit planes that are identically zero.  The two forward
        # products multiply dense operands (28 pairs); the four adjoint products meet deriv(output) of `sum`, a constant fill with ONE
        # live plane (7 pairs): tensor work shrinks 4x and those launches are bound by operand delivery (L2 -> shared memory), not by
        # the tensor pipe.  `frac` is therefore taken on the DENSE launches - the kernel in the regime the roofline describes - and
        # the all-launch figure is reported next to it.
        pairs_per_launch_all = gstats["slice_pairs"] / gstats["checked"]
        fwd = [s_ for s_ in stats if s_["step"] == "gemm_forward"]
        rev = [s_ for s_ in stats if s_["step"].startswith("gemm_reverse")]
        f_ms = sum(s_["ms"] for s_ in fwd); f_cnt = sum(s_["count"] for s_ in fwd)
        r_ms = sum(s_["ms"] for s_ in rev); r_cnt = sum(s_["count"] for s_ in rev)
        tot_pairs = gstats["slice_pairs"]
        dense_pairs = ops_per_flop
        rev_pairs = (tot_pairs - dense_pairs * f_cnt * gstats["checked"] / max(g_cnt, 1)) / max(r_cnt * gstats["checked"] / max(g_cnt, 1), 1)
        ops_launch = 2.0 * n ** 3
        launch_kinds = {
            "dense_forward": {"launches_per_step": f_cnt // reps, "avg_launch_ms": f_ms / max(f_cnt, 1), "slice_pairs_per_launch": dense_pairs,
                              "int8_tops": dense_pairs * ops_launch / (f_ms / max(f_cnt, 1) * 1e-3) / 1e12},
            "trimmed_reverse": {"launches_per_step": r_cnt // reps, "avg_launch_ms": r_ms / max(r_cnt, 1), "slice_pairs_per_launch": rev_pairs,
                                "int8_tops": rev_pairs * ops_launch / (r_ms / max(r_cnt, 1) * 1e-3) / 1e12,
                                "bound": "operand delivery L2 -> shared memory (one live plane of deriv(output))"},
            "all_launches": {"slice_pairs_per_launch": pairs_per_launch_all,
                             "int8_tops": pairs_per_launch_all * ops_launch / (g_ms / max(g_cnt, 1) * 1e-3) / 1e12},
        }
        achieved = launch_kinds["dense_forward"]["int8_tops"]
    roofline = {
        "bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": pipe_peak, "unit": unit, "frac": achieved / pipe_peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "note": "achieved = tensor-pipe operations actually issued by the DENSE launches of the kernel (the two forward products: "
                "28 slice pairs x 2mnk) / their CUDA-event time INCLUDING the slicing passes of both operands; launch_kinds holds the "
                "trimmed adjoint launches and the all-launch average (slice pairs counted on the device, rdb_tape_slice_pairs)",
        "ops_per_algorithmic_flop": ops_per_flop,
        "fp_equivalent": {"achieved_tflops": fp_equiv, "cublas_gemm_tflops_same_run": blas_peak, "ratio_vs_cublas": fp_equiv / blas_peak,
                          "what": f"algorithmic 2mnk flops of the six {n}^3 `*` instructions / their time, next to cuBLAS "
                                  f"{'D' if args.dtype == 'f64' else 'S'}GEMM {n}^3 (best of 6)"},
        "launch_kinds": launch_kinds,
        "algorithmic_flops_per_launch": 2.0 * n ** 3, "launches_per_step": g_cnt // reps,
        "avg_launch_ms": g_ms / max(g_cnt, 1), "gemm_share_of_step": g_ms / tot_ms,
        "elementwise": {"bound": "hbm", "achieved": ew_bytes / (ew_ms * 1e-3) / 1e9 if ew_ms > 0 else None, "peak": hbm_peak,
                        "unit": "GB/s", "frac": (ew_bytes / (ew_ms * 1e-3) / 1e9) / hbm_peak if ew_ms > 0 else None,
                        "ms_per_step": ew_ms / reps,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
        "steps": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items()} for s in stats],
    }
    if args.dtype == "f64" and mode == 0:
        roofline["cublaslt_int8_tops_same_run"] = i8
        # the DMMA kernel on the same tape, same run, for context
        st1, r1, _ = gemm_stats(1)
        d_ms = sum(s["ms"] for s in st1 if s["step"].startswith("gemm")); d_fl = sum(s["flops"] for s in st1 if s["step"].startswith("gemm"))
        roofline["dmma_path_same_run"] = {"gemm_tflops": d_fl / (d_ms * 1e-3) / 1e12, "frac_of_cublas_dgemm": d_fl / (d_ms * 1e-3) / 1e12 / blas_peak,
                                          "ms_per_eval_all_steps": sum(s["ms"] for s in st1) / r1}

    # ---- CPU baseline (rank 0, N=1): reference-equivalent replay on the host cores ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        orc = graft.load_oracle()
        cores = use_all_host_cores()
        times = cpu_replay(rd, orc, n, npdt, 2)
        cpu = {"value": 1.0 / min(times), "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": f"2 full-size evals ({n}x{n} {args.dtype}) of the numpy/OpenBLAS oracle, best taken; "
                         f"{cores} BLAS threads of {os.cpu_count()} host cores"}

    extra = {}
    if not args.skip_extra:
        extra = sharded_legs(args, torch, dist, rd, ctx, rank, world)
        if rank == 0 and world == 1:
            extra["other_configs"] = other_configs(args, torch, rd, ctx)

    if rank == 0:
        line = {
            "metric": "gradient! evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(n, args.dtype),
            "config_detail": {"tape": "*(IDX a', b) ; *(a, IDX b') ; + ; sum", "flops_per_eval": 12.0 * n ** 3, "gemm_kernels": args.gemm,
                              "l2": "working set (8 x %d MB) larger than the 126 MB L2" % (n * n * es // 2 ** 20),
                              "multi_gpu": "replicas only (a scalar function's gradient does not shard)" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": 2 * n * n * es, "d2h_bytes_per_step": 2 * n * n * es + es,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "gpu_launches": int(e2e_launches),
                    "how": f"rd.GradientPipeline(depth={depth}): rdb_tape_set_input + rdb_gradient_async per step, rdb_tape_wait before a slot is "
                           "reused; pinned host inputs and results; host clock from the first submit to the last result byte in host memory",
                    "synchronous_calls": {"value": world * e2e_sync_steps / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / e2e_sync_steps,
                                          "how": "one rdb_gradient per step, returning when its results are in host memory"},
                    "host_link": pcie, "rank0_numa": numa},
            "gpu_launches": int(l_per_step * args.steps),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity": {"max_rel_err_vs_closed_form": float(err), "tol": tol},
            "gemm_guard": ct.handle.guard_stats(),
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    _ = keep


def adaptive_steps(torch, step, min_steps=10, min_seconds=0.25, cap=2000):
    """enough timed steps for >= min_seconds of device time (the hessian! shard of one rank is tens of microseconds)"""
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1), 1e-3)
    return int(min(cap, max(min_steps, np.ceil(min_seconds * 1e3 / est))))


def agree_steps(torch, dist, world, k):
    if world > 1:
        t = torch.tensor([k], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k = int(t.item())
    return k


def timed_once(torch, dist, world, fn, reps=3):
    """best of `reps` single calls, CUDA events, max over ranks (ms) - for the collectives"""
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = min(best, ms)
    return best


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def sharded_legs(args, torch, dist, rd, ctx, rank, world):
    """jacobian! rows and hessian! columns sharded over the ranks: every rank replays its own tape copy over
    its shard, no collective on the data path; results are left sharded in device memory for the timed region, and the product's
    own gather (rdb_gather_result: NCCL inside librdb200.so, blocks placed into the reference's column-major result) is timed
    separately."""
    out = {}
    peaks = load_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bf16_peak = peaks.get("bf16_tflops", 1590.0)
    comm = rd.make_comm(ctx) if world > 1 else None
    # ---- jacobian!: f(x) = (A*x + x) .* tanh.(B*x), output-seed rows [rank*m/G, (rank+1)*m/G)
    n = args.jac_n
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    ct = rd.compile(tape, ctx=ctx)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    x_d = torch.from_numpy(x).cuda()
    rb, re = rd.shard_range(n, world, rank)
    J_d = torch.empty((re - rb) * n, dtype=torch.float64, device="cuda")

    def step():
        rd.jacobian_(J_d, ct, x_d, rows=(rb, re))
    k = agree_steps(torch, dist, world, adaptive_steps(torch, step))
    ms = timed_steps(torch, dist, world, step, k, 3)
    per = ms / k
    gather_ms = None
    if comm is not None:
        full = torch.empty(n * n, dtype=torch.float64, device="cuda")
        gather_ms = timed_once(torch, dist, world, lambda: rd.gather_jacobian_(full, J_d, ct, comm))
        # the gathered matrix IS the reference's result layout: spot-check rank 0's copy against the closed form
        if rank == 0:
            s_ = A @ x + x; t_ = np.tanh(B @ x)
            rows_chk = np.array([0, n // 3, n - 1])
            Jrows = full.reshape(n, n).T[torch.from_numpy(rows_chk).cuda()].cpu().numpy()
            Jcf = t_[rows_chk, None] * (A[rows_chk] + np.eye(n)[rows_chk]) + (s_ * (1 - t_ ** 2))[rows_chk, None] * B[rows_chk]
            assert np.max(np.abs(Jrows - Jcf)) < 1e-12 * np.max(np.abs(Jcf)), "gathered Jacobian does not match the closed form"
        del full
    flops_row = 4.0 * n * n                                     # two adjoint GEMV-columns per seed row
    tf_equiv = flops_row * n / (per * 1e-3) / 1e12              # whole job: all n rows per `per` ms
    out["jacobian"] = {"metric": "jacobian! rows/s", "value": n / (per * 1e-3), "unit": "rows/s", "n": n, "ms_per_jacobian": per,
                       "timed_steps": k, "rows_per_rank": re - rb, "result": "left sharded in device memory for the timed region",
                       "gather_ms": gather_ms, "gather": "rdb_gather_result (NCCL broadcasts + block placement into the column-major m x n result)",
                       "roofline": {"bound": "tensor", "algorithmic_flops_per_row": flops_row, "achieved_fp64_equiv_tflops_per_gpu": tf_equiv / world,
                                    "achieved": tf_equiv / world * 28.0, "peak": 2.0 * bf16_peak, "unit": "TOP/s (int8)",
                                    "frac": tf_equiv / world * 28.0 / (2.0 * bf16_peak),
                                    "note": "28 int8 slice products per FP64 product (7 slices of 8 bits); includes the replicated forward "
                                            "sweep, seeding, elementwise adjoints and row extraction of every block"},
                       "gemm_guard": ct.handle.guard_stats()}
    del ct, J_d
    # ---- hessian!: extended Rosenbrock, column chunks
    for key, n in (("hessian", args.hess_n), ("hessian_large", args.hess_large_n)):
        if n <= 0:
            continue
        cb, ce = rd.shard_range(n, world, rank)
        need = n * (ce - cb) * 8
        free, _total = torch.cuda.mem_get_info()
        if need * 1.5 + (40 << 30) > free:
            out[key] = {"skipped": f"needs {need / 2**30:.1f} GiB result + tangent workspace, {free / 2**30:.1f} GiB free"}
            continue
        tape = rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1))
        # a context with a stream of its own: the launch-bound sweep replays as one CUDA graph (the legacy default stream cannot
        # be captured); every call is synchronous, so the events on torch's stream still bracket the work
        hctx = rd.engine.Context(torch.cuda.current_device())
        hstream = torch.cuda.Stream()
        hctx.set_stream(hstream.cuda_stream)                  # a non-default torch stream: the timing events are recorded on it
        ct = rd.compile(tape, ctx=hctx)
        # device-result hessian! calls are only enqueued (rdb_tape_set_async): a rank's shard is tens of microseconds of device
        # work, so consecutive calls overlap their launch latency; the timed region ends with a device synchronisation
        ct.handle.set_async(True)
        x_d = torch.from_numpy(rd.workloads.inputs_rosenbrock(n, seed=2)).cuda()
        torch.cuda.synchronize()
        H_d = torch.empty(n * (ce - cb), dtype=torch.float64, device="cuda")

        def hstep():
            rd.hessian_(H_d, ct, x_d, cols=(cb, ce))
        with torch.cuda.stream(hstream):
            k = agree_steps(torch, dist, world, adaptive_steps(torch, hstep))
            ms = timed_steps(torch, dist, world, hstep, k, 3)
        torch.cuda.synchronize()
        ct.handle.set_async(False)
        per = ms / k
        # parity of what was just timed (closed form): the block's non-zeros and the count of exact zeros
        xs = x_d
        cc = torch.arange(cb, ce, device="cuda")
        Hb = H_d.reshape(ce - cb, n)                          # row j = column cb + j of H
        o_idx = (cc // 2) * 2
        o = xs[o_idx]; e_ = xs[o_idx + 1]
        even = (cc % 2 == 0)
        diag_ref = torch.where(even, 1200 * o * o - 400 * e_ + 2, torch.full_like(o, 200.0))
        off_ref = -400 * o
        j = torch.arange(ce - cb, device="cuda")
        diag = Hb[j, cc]; off = Hb[j, torch.where(even, cc + 1, cc - 1)]
        herr = float(torch.maximum((diag - diag_ref).abs().max() / diag_ref.abs().max(), (off - off_ref).abs().max() / off_ref.abs().max()))
        if not (herr < 1e-12 and int(torch.count_nonzero(H_d)) == 2 * (ce - cb)):
            raise SystemExit(f"hessian parity check failed: {herr}")
        gather_ms = None
        if comm is not None and key == "hessian":
            full = torch.empty(n * n, dtype=torch.float64, device="cuda")
            gather_ms = timed_once(torch, dist, world, lambda: rd.gather_hessian_(full, H_d, ct, comm))
            del full
        gbps = n * (ce - cb) * 8 / (per * 1e-3) / 1e9
        out[key] = {"metric": "hessian! rows/s", "value": n / (per * 1e-3), "unit": "rows/s", "n": n, "ms_per_hessian": per,
                    "timed_steps": k, "cols_per_rank": ce - cb, "result": "left sharded in device memory for the timed region",
                    "gather_ms": gather_ms, "max_rel_err_vs_closed_form": herr, "calls": "asynchronous (rdb_tape_set_async), one fused kernel per call",
                    "roofline": {"bound": "hbm", "algorithmic_bytes_per_row": n * 8, "achieved": gbps, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": gbps / hbm_peak, "note": "per GPU: the dense n x cols result block is the traffic (SURVEY section 8d)"}}
        del ct, H_d, x_d, hctx
        torch.cuda.empty_cache()
    if comm is not None:
        comm.close()
    return out


def other_configs(args, torch, rd, ctx):
    """The BASELINE configs that are not the headline, one short measurement each (rank 0, N = 1), so the driver's record holds
    them: cfg 1 (100 x 100, the only size the reference publishes a number for), cfg 2 in FP32, cfg 3 (MLP)."""
    out = {}
    # the headline leg once more in a child process with plane trimming switched off (the knob is read once per process): every
    # digit plane of every operand is multiplied, which is what a tape whose `*` adjoints are dense (no constant-fill
    # deriv(output)) pays - the same-run number to hold next to `value`
    if args.dtype == "f64" and args.gemm == "auto":
        try:
            import subprocess
            e = dict(os.environ); e["RDB_OZAKI_TRIM"] = "0"
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--only-value", "--steps", str(max(5, args.steps // 2)), "--warmup", "3",
                                "--n", str(args.n)], env=e, capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            d = json.loads(line)
            out["cfg2_without_plane_trimming"] = {"evals_per_s": d["value"], "ms_per_eval": d["ms_per_step"],
                                                  "max_rel_err_vs_closed_form": d["max_rel_err_vs_closed_form"],
                                                  "slice_pairs_per_eval": d["gemm_guard"]["slice_pairs"] / max(d["gemm_guard"]["checked"], 1) * 6,
                                                  "how": "child process, RDB_OZAKI_TRIM=0: all 28 slice pairs of all six products"}
        except Exception as ex:  # noqa: BLE001 - a reporting extra must not sink the bench line
            out["cfg2_without_plane_trimming"] = {"error": repr(ex)[:200]}

    def timeit(step, steps, warm=3):
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    # cfg 1: 100 x 100 FP64, host arrays through the public call (wall clock, synchronous calls): README reports 429.65 us
    n = 100
    rng = np.random.default_rng(1)
    c1ctx = rd.engine.Context(torch.cuda.current_device())        # own stream: the sweep replays as one CUDA graph
    ct = rd.compile(rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((n, n)), rng.random((n, n)))), ctx=c1ctx)
    a, b = rd.workloads.inputs_gradient_example(n)
    res = (np.zeros((n, n), order="F"), np.zeros((n, n), order="F"))
    for _ in range(20):
        rd.gradient_(res, ct, (a, b))
    t0 = time.perf_counter()
    for _ in range(500):
        rd.gradient_(res, ct, (a, b))
    us = (time.perf_counter() - t0) / 500 * 1e6
    ok = np.max(np.abs(res[0] - (b.sum(1)[:, None] + b.sum(0)[None, :]))) < 1e-12 * np.max(np.abs(res[0]))
    out["cfg1_100x100_f64"] = {"us_per_gradient_host_arrays_wall_clock": us, "evals_per_s": 1e6 / us, "reference_readme_us": 429.65,
                               "vs_readme": 429.65 / us, "parity_ok": bool(ok),
                               "note": "README.md timing is of unknown CPU hardware (BASELINE.md); includes the Python binding"}
    del ct, c1ctx
    # cfg 2 in FP32 (3xTF32 on tcgen05)
    n = args.n
    a0 = np.asfortranarray(rng.random((n, n)).astype(np.float32)); b0 = np.asfortranarray(rng.random((n, n)).astype(np.float32))
    ct = rd.compile(rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0), dtype=np.float32), ctx=ctx)
    a, b = rd.workloads.inputs_gradient_example(n, np.float32)
    ad, bd = rd.engine.device_array(a), rd.engine.device_array(b)
    ga = torch.empty(n * n, dtype=torch.float32, device="cuda"); gb = torch.empty_like(ga)
    ms = timeit(lambda: rd.gradient_((ga, gb), ct, (ad, bd)), 10)
    b64 = b.astype(np.float64)
    ref = b64.sum(1)[:, None] + b64.sum(0)[None, :]
    err = float(np.max(np.abs(ga.cpu().numpy().reshape(n, n, order="F") - ref)) / np.max(np.abs(ref)))
    out["cfg2_f32"] = {"evals_per_s": 1e3 / ms, "ms_per_eval": ms, "fp32_equiv_tflops": 12.0 * n ** 3 / ms / 1e9, "max_rel_err_vs_closed_form": err, "tol": 1e-5}
    del ct, ga, gb, ad, bd
    # cfg 3: MLP, X 1024 x 65536 frozen in the tape
    d = h = o = 1024; N = 65536
    X, rec = rd.workloads.inputs_mlp(d, h, o, N, seed=1)
    tape = rd.GradientTape(rd.workloads.make_f_mlp(X), rec)
    ct = rd.compile(tape, ctx=ctx)
    _, ws = rd.workloads.inputs_mlp(d, h, o, N, seed=2)
    wd = [rd.engine.device_array(w) for w in ws]
    gs = [torch.empty(w.size, dtype=torch.float64, device="cuda") for w in ws]
    ms = timeit(lambda: rd.gradient_(tuple(gs), ct, tuple(wd)), 5)
    Xd = torch.from_numpy(X).cuda(); W1, b1, W2, b2 = [torch.from_numpy(np.ascontiguousarray(w)).cuda() for w in ws]
    Hh = torch.tanh(W1 @ Xd + b1[:, None]); Y = torch.tanh(W2 @ Hh + b2[:, None])
    dZ2 = 1 - Y * Y; dZ1 = (W2.T @ dZ2) * (1 - Hh * Hh)
    refs = [dZ1 @ Xd.T, dZ1.sum(1), dZ2 @ Hh.T, dZ2.sum(1)]
    errs = []
    for g, r, w in zip(gs, refs, ws):
        gg = g.reshape(w.shape[::-1]).T if w.ndim == 2 else g
        errs.append(float((gg - r).abs().max() / r.abs().max()))
    S = h * N
    out["cfg3_mlp_f64"] = {"evals_per_s": 1e3 / ms, "ms_per_eval": ms, "gemm_flops_per_eval": 10.0 * 1024 ** 2 * N,
                           "elementwise_bytes_algorithmic": 14 * S * 8, "max_rel_err_vs_torch_fp64_backprop": max(errs), "tol": 1e-12,
                           "gemm_guard": ct.handle.guard_stats()}
    return out


if __name__ == "__main__":
    main()
