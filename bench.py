#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 tape-replay engine.

Metric (BASELINE.json): compiled-tape ``gradient!`` evals/s of ``f(a,b) = sum(a'*b + a*b')`` at 4096x4096 FP64
(configs[1]).  A "step" is one gradient! replay (forward sweep + seeded reverse sweep + result extraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 4096] [--dtype f64|f32]

* ``value``     device-resident inputs/results (D2D copies only), CUDA events, max over ranks.
* ``e2e``       same call with pinned HOST inputs and HOST results: H2D + replay + D2H inside the timed region.
* ``roofline``  dominant kernel (the DMMA GEMM): 12 n^3 flop per eval / summed CUDA-event time of the six GEMM
                launches, against a cuBLAS DGEMM measured in the same run (MEASURED_PEAKS.json has no FP64 entry).
* ``cpu_baseline`` the oracle (reference-equivalent CPU replay, numpy/OpenBLAS) on the box's host cores.
* N > 1: gradient! of one scalar function does not shard ("replicas only"): every rank replays its
  own tape copy, value = N evals / max-over-ranks time.  The same run then measures the paths that DO shard —
  jacobian! by output-seed rows and hessian! by column chunks — and reports them under "jacobian"/"hessian".
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


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    rd = graft.load_package(); orc = graft.load_oracle()
    npdt = np.float64 if args.dtype == "f64" else np.float32
    n = args.n
    times = cpu_replay(rd, orc, n, npdt, args.warmup + args.steps)[args.warmup:]
    total = float(np.sum(times))
    val = args.steps / total
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "gradient! evals/s", "value": val, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"compiled-tape gradient! of sum(a'*b + a*b') at {n}x{n} {args.dtype} (BASELINE configs[1])", "n": n},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full-size evals of the numpy/OpenBLAS oracle (reference-equivalent CPU replay; "
                                   f"Julia is not installed so ReverseDiff.jl itself cannot run), {cores} BLAS threads"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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

    if rank == 0:
        sampler.start()
    ms_e2e = timed_steps(torch, dist, world, step_e2e, max(3, args.steps // 2), 3)
    e2e_steps = max(3, args.steps // 2)
    e2e_val = world * e2e_steps / (ms_e2e * 1e-3)
    clocks = sampler.stop() if rank == 0 else None           # sampled every 20 ms during both timed regions (device-resident, end-to-end)
    assert np.max(np.abs(ga_h - ga)) <= tol * np.max(np.abs(ga))
    gb_ref = a64.sum(1)[:, None] + a64.sum(0)[None, :]
    assert np.max(np.abs(gb_h - gb_ref)) <= tol * np.max(np.abs(gb_ref))

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
        del ctp
        return st, reps_
    stats, reps = gemm_stats(mode)
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
        traffic = 1.0387e9 + 0.1262e9   # dram__bytes_read+write per 4096^3 launch, ncu --set full (profiles/prof_ozaki2c_r1c.txt)
    elif args.dtype == "f64":
        ops_per_flop = 1.0
        kernel = "d64::dgemm_dmma_kernel (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4)"
        unit, pipe_peak = "TFLOP/s (fp64)", blas_peak
        peak_src = "cuBLAS DGEMM measured in this run (MEASURED_PEAKS.json holds no FP64 figure)"
        traffic = 1.084e9 + 0.121e9     # profiles/prof_dgemm_r1.txt
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
        traffic = None
    achieved = fp_equiv * ops_per_flop
    roofline = {
        "bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": pipe_peak, "unit": unit, "frac": achieved / pipe_peak,
        "traffic": traffic, "peak_source": peak_src,
        "note": "achieved = tensor-pipe operations actually issued (algorithmic 2mnk x ops_per_flop) / CUDA-event time of the `*` "
                "steps INCLUDING their operand slicing passes",
        "ops_per_algorithmic_flop": ops_per_flop,
        "fp_equivalent": {"achieved_tflops": fp_equiv, "cublas_gemm_tflops_same_run": blas_peak, "ratio_vs_cublas": fp_equiv / blas_peak,
                          "what": f"algorithmic 2mnk flops of the six {n}^3 `*` instructions / their time, next to cuBLAS "
                                  f"{'D' if args.dtype == 'f64' else 'S'}GEMM {n}^3 (best of 6)"},
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
        st1, r1 = gemm_stats(1)
        d_ms = sum(s["ms"] for s in st1 if s["step"].startswith("gemm")); d_fl = sum(s["flops"] for s in st1 if s["step"].startswith("gemm"))
        roofline["dmma_path_same_run"] = {"gemm_tflops": d_fl / (d_ms * 1e-3) / 1e12, "frac_of_cublas_dgemm": d_fl / (d_ms * 1e-3) / 1e12 / blas_peak,
                                          "ms_per_eval_all_steps": sum(s["ms"] for s in st1) / r1}

    # ---- CPU baseline (rank 0, N=1): reference-equivalent replay on the host cores ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        orc = graft.load_oracle()
        times = cpu_replay(rd, orc, n, npdt, 2)
        cores = blas_threads()
        cpu = {"value": 1.0 / min(times), "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": f"2 full-size evals ({n}x{n} {args.dtype}) of the numpy/OpenBLAS oracle, best taken; "
                         f"{cores} BLAS threads of {os.cpu_count()} host cores"}

    extra = {}
    if not args.skip_extra:
        extra = sharded_legs(args, torch, dist, rd, ctx, rank, world)

    if rank == 0:
        line = {
            "metric": "gradient! evals/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"compiled-tape gradient! of sum(a'*b + a*b') at {n}x{n} {args.dtype} (BASELINE configs[1])",
                       "n": n, "tape": "*(IDX a', b) ; *(a, IDX b') ; + ; sum", "flops_per_eval": 12.0 * n ** 3, "gemm_kernels": args.gemm,
                       "l2": "working set (8 x %d MB) larger than the 126 MB L2" % (n * n * es // 2 ** 20),
                       "multi_gpu": "replicas only (a scalar function's gradient does not shard)" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": 2 * n * n * es, "d2h_bytes_per_step": 2 * n * n * es + es,
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(l_per_step * args.steps),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity": {"max_rel_err_vs_closed_form": float(err), "tol": tol},
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    _ = keep


def sharded_legs(args, torch, dist, rd, ctx, rank, world):
    """jacobian! rows and hessian! columns sharded over the ranks, results left sharded in device memory
    (SURVEY.md §8e); the all-gather of the result blocks is timed separately."""
    out = {}
    # ---- jacobian!: f(x) = (A*x + x) .* tanh.(B*x), output-seed rows [rank*m/G, (rank+1)*m/G)
    n = args.jac_n
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    ct = rd.compile(tape, ctx=ctx)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    x_d = torch.from_numpy(x).cuda()
    rows = n // world
    rb, re = rd.shard_range(n, world, rank)
    J_d = torch.empty((re - rb) * n, dtype=torch.float64, device="cuda")

    def step():
        rd.jacobian_(J_d, ct, x_d, rows=(rb, re))
    ms = timed_steps(torch, dist, world, step, 3, 2)
    full = torch.empty(world * rows * n, dtype=torch.float64, device="cuda") if (n % world == 0 and world > 1) else None
    gather_ms = None
    if full is not None:
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_gather_into_tensor(full, J_d); e1.record(); torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
    out["jacobian"] = {"metric": "jacobian! rows/s", "value": n * 3 / (ms * 1e-3), "unit": "rows/s", "n": n, "ms_per_jacobian": ms / 3,
                       "rows_per_rank": rows, "result": "left sharded in device memory", "allgather_ms": gather_ms}
    del ct, J_d, full
    # ---- hessian!: extended Rosenbrock, column chunks
    for key, n in (("hessian", args.hess_n), ("hessian_large", args.hess_large_n)):
        if n <= 0:
            continue
        cols = n // world
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
        ct = rd.compile(tape, ctx=hctx)
        x_d = torch.from_numpy(rd.workloads.inputs_rosenbrock(n, seed=2)).cuda()
        torch.cuda.synchronize()
        H_d = torch.empty(n * (ce - cb), dtype=torch.float64, device="cuda")

        def hstep():
            rd.hessian_(H_d, ct, x_d, cols=(cb, ce))
        ms = timed_steps(torch, dist, world, hstep, 3, 3)
        out[key] = {"metric": "hessian! rows/s", "value": n * 3 / (ms * 1e-3), "unit": "rows/s", "n": n, "ms_per_hessian": ms / 3,
                    "cols_per_rank": cols, "result": "left sharded in device memory",
                    "result_write_GBps_per_gpu": n * (ce - cb) * 8 / (ms / 3 * 1e-3) / 1e9}
        del ct, H_d, x_d, hctx
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
