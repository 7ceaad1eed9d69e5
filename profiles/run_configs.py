#!/usr/bin/env python
"""Measure all five BASELINE.json configs on one B200 (device-resident inputs, CUDA events, 3 warm-ups) and print one JSON
object per config.  Used for profiles/README.md; the headline contract lives in bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

rd = graft.load_package()
ctx = rd.engine.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)


def timeit(step, steps=5, warm=3):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def stats_of(tape, run, **opts):
    ctp = rd.compile(tape, ctx=ctx, profile=1, **opts)
    run(ctp); run(ctp)
    ctp.handle.stats()
    run(ctp)
    return json.loads(ctp.handle.stats())


def dev(x):
    return rd.engine.device_array(x)      # same shape, column-major memory


out = []
which = sys.argv[1:] or ["1", "2", "2f32", "3", "4", "5"]

if "1" in which:   # cfg 1: 100x100 FP64 (README: 429.65 us per gradient! on unstated CPU hardware)
    n = 100
    rng = np.random.default_rng(1)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((n, n)), rng.random((n, n))))
    ctx1 = rd.engine.Context(0)          # own (capturable) stream: the whole sweep replays as one CUDA graph
    a, b = rd.workloads.inputs_gradient_example(n)
    ad, bd = dev(a), dev(b); ga, gb = torch.empty(n * n, dtype=torch.float64, device="cuda"), torch.empty(n * n, dtype=torch.float64, device="cuda")
    gh, gbh = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
    res = {}
    for label, g in (("graph", 1), ("eager", 0)):
        ct = rd.compile(tape, ctx=ctx1, use_graph=g)
        def wall(step, reps=300, warm=30):
            for _ in range(warm):
                step()
            t0 = time.perf_counter()
            for _ in range(reps):
                step()
            return (time.perf_counter() - t0) / reps * 1e6
        res[label] = {"us_per_eval_device_arrays": wall(lambda: rd.gradient_((ga, gb), ct, (ad, bd))),
                      "us_per_eval_host_arrays": wall(lambda: rd.gradient_((gh, gbh), ct, (a, b)))}
    out.append({"cfg": 1, "what": "gradient! 100x100 f64 (wall clock per synchronous call)", **res,
                "evals_per_s_host_arrays_graph": 1e6 / res["graph"]["us_per_eval_host_arrays"], "readme_cpu_us": 429.65})

for key, dt, tdt in (("2", np.float64, torch.float64), ("2f32", np.float32, torch.float32)):
    if key in which:
        n = 4096
        rng = np.random.default_rng(1)
        tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((n, n)).astype(dt), rng.random((n, n)).astype(dt)))
        ct = rd.compile(tape, ctx=ctx)
        a, b = rd.workloads.inputs_gradient_example(n, dt)
        ad, bd = dev(a), dev(b); ga, gb = torch.empty(n * n, dtype=tdt, device="cuda"), torch.empty(n * n, dtype=tdt, device="cuda")
        ms = timeit(lambda: rd.gradient_((ga, gb), ct, (ad, bd)))
        out.append({"cfg": key, "what": f"gradient! 4096x4096 {np.dtype(dt).name}", "ms_per_eval": ms, "evals_per_s": 1e3 / ms,
                    "tflops_equiv": 12 * n ** 3 / ms / 1e9, "steps": stats_of(tape, lambda c: rd.gradient_((ga, gb), c, (ad, bd)))})
        del ct

if "3" in which:   # cfg 3: MLP, X 1024 x 65536 frozen in the tape
    d = h = o = 1024; N = 65536
    X, rec = rd.workloads.inputs_mlp(d, h, o, N, seed=1)
    tape = rd.GradientTape(rd.workloads.make_f_mlp(X), rec)
    ct = rd.compile(tape, ctx=ctx)
    _, ws = rd.workloads.inputs_mlp(d, h, o, N, seed=2)
    wd = [dev(w) for w in ws]
    gs = [torch.empty(w.size, dtype=torch.float64, device="cuda") for w in ws]
    ms = timeit(lambda: rd.gradient_(tuple(gs), ct, tuple(wd)), 3, 2)
    # parity at full size: standard backprop in torch fp64 (independent of the engine)
    Xd = torch.from_numpy(X).cuda(); W1, b1, W2, b2 = [torch.from_numpy(np.asarray(w)).cuda() for w in ws]
    H = torch.tanh(W1 @ Xd + b1[:, None]); Y = torch.tanh(W2 @ H + b2[:, None])
    dZ2 = 1 - Y * Y; dZ1 = (W2.T @ dZ2) * (1 - H * H)
    refs = [dZ1 @ Xd.T, dZ1.sum(1), dZ2 @ H.T, dZ2.sum(1)]
    errs = []
    for g, r, w in zip(gs, refs, ws):
        gg = g.reshape(w.shape[::-1]).T if w.ndim == 2 else g
        errs.append(float((gg - r).abs().max() / r.abs().max()))
    S = h * N
    out.append({"cfg": 3, "what": "gradient! MLP sum(tanh.(W2*tanh.(W1*X .+ b1) .+ b2)), X 1024x65536 f64", "ms_per_eval": ms, "evals_per_s": 1e3 / ms,
                "gemm_flops": 10.0 * 1024 ** 2 * N, "elementwise_bytes_algorithmic": 14 * S * 8, "max_rel_err_vs_torch_backprop": max(errs),
                "steps": stats_of(tape, lambda c: rd.gradient_(tuple(gs), c, tuple(wd)))})
    del ct, Xd, H, Y, dZ2, dZ1

if "4" in which:   # cfg 4: jacobian! 8192 -> 8192
    n = 8192
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    ct = rd.compile(tape, ctx=ctx)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    xd = dev(x); J = torch.empty(n * n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: rd.jacobian_(J, ct, xd), 3, 2)
    Ad, Bd, xt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(x).cuda()
    s = Ad @ xt + xt; v = Bd @ xt; t = torch.tanh(v)
    Jcf = t[:, None] * (Ad + torch.eye(n, dtype=torch.float64, device="cuda")) + (s * (1 - t * t))[:, None] * Bd
    Jg = J.reshape(n, n).T          # column-major (rows = outputs) viewed row-major
    err = float((Jg - Jcf).abs().max() / Jcf.abs().max())
    out.append({"cfg": 4, "what": "jacobian! f(x)=(A*x+x).*tanh.(B*x), R^8192->R^8192 f64", "ms_per_jacobian": ms, "rows_per_s": n / ms * 1e3,
                "gemm_flops": 4.0 * n ** 3, "tflops_equiv": 4.0 * n ** 3 / ms / 1e9, "max_rel_err_vs_closed_form": err,
                "steps": stats_of(tape, lambda c: rd.jacobian_(J, c, xd))})
    del ct, Ad, Bd, Jcf

if "5" in which:   # cfg 5: hessian! extended Rosenbrock n = 16384
    n = 16384
    tape = rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1))
    ct = rd.compile(tape, ctx=ctx)
    x = rd.workloads.inputs_rosenbrock(n, seed=2)
    xd = dev(x); H = torch.empty(n * n, dtype=torch.float64, device="cuda")
    ms = timeit(lambda: rd.hessian_(H, ct, xd), 3, 2)
    Hm = H.reshape(n, n)
    xt = torch.from_numpy(x).cuda(); oo, ee = xt[0::2], xt[1::2]
    idx = torch.arange(n // 2, device="cuda")
    Hcf = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    Hcf[2 * idx, 2 * idx] = 1200 * oo ** 2 - 400 * ee + 2
    Hcf[2 * idx, 2 * idx + 1] = -400 * oo; Hcf[2 * idx + 1, 2 * idx] = -400 * oo
    Hcf[2 * idx + 1, 2 * idx + 1] = 200
    err = float((Hm - Hcf).abs().max() / Hcf.abs().max())
    zeros_exact = bool((Hm[Hcf == 0] == 0).all())
    Hs = torch.empty(n * (n // 8), dtype=torch.float64, device="cuda")
    hctx = rd.engine.Context(0)                                                    # own stream: the sweep replays as a CUDA graph
    cth = rd.compile(tape, ctx=hctx)
    torch.cuda.synchronize()
    ms_graph = timeit(lambda: rd.hessian_(H, cth, xd), 10, 3)
    ms_shard = timeit(lambda: rd.hessian_(Hs, cth, xd, cols=(0, n // 8)), 20, 3)     # what ONE rank does at 8 GPUs
    out.append({"cfg": 5, "what": "hessian! extended Rosenbrock n=16384 f64 (forward-over-reverse)", "ms_per_hessian": ms, "rows_per_s": n / ms * 1e3,
                "ms_per_hessian_graph": ms_graph, "ms_one_eighth_column_shard_graph": ms_shard, "speedup_if_8_ranks": ms_graph / ms_shard,
                "result_bytes": n * n * 8, "result_write_GBps": n * n * 8 / ms / 1e6, "max_rel_err_vs_closed_form": err,
                "structural_zeros_exact": zeros_exact, "steps": stats_of(tape, lambda c: rd.hessian_(H, c, xd))})

if "scalar" in which:   # §8 row a13: scalar-instruction tapes (the reference's own HessianTape; loop-written gradients)
    orc = graft.load_oracle()

    def rosen_H(xt, n):
        oo, ee = xt[0::2], xt[1::2]
        idx = torch.arange(n // 2, device="cuda")
        Hcf = torch.zeros(n, n, dtype=torch.float64, device="cuda")
        Hcf[2 * idx, 2 * idx] = 1200 * oo ** 2 - 400 * ee + 2
        Hcf[2 * idx, 2 * idx + 1] = -400 * oo; Hcf[2 * idx + 1, 2 * idx] = -400 * oo
        Hcf[2 * idx + 1, 2 * idx + 1] = 200
        return Hcf
    for n in [int(v) for v in os.environ.get("SCALAR_HESS_N", "1024,4096").split(",")]:
        t0 = time.perf_counter()
        tape = rd.HessianTape(rd.workloads.f_rosenbrock_loop, rd.workloads.inputs_rosenbrock(n, seed=1), mode="reverse")
        rec_s = time.perf_counter() - t0
        blk = tape.tape.program.instrs[0].scalar
        ct = rd.compile(tape, ctx=ctx)
        x = rd.workloads.inputs_rosenbrock(n, seed=2)
        xd = dev(x); H = torch.empty(n * n, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: rd.hessian_(H, ct, xd), 3, 2)
        Hm = H.reshape(n, n).T
        Hcf = rosen_H(torch.from_numpy(x).cuda(), n)
        err = float((Hm - Hcf).abs().max() / Hcf.abs().max())
        zeros_exact = bool((Hm[Hcf == 0] == 0).all())
        # CPU: the same tape through the oracle's C loops (one forward sweep + one reverse sweep per row), a sample of rows
        ot = orc.OracleTape(tape.program_bytes())
        rows = min(n, 256)
        t0 = time.perf_counter(); ot.jacobian([x], rows=range(rows)); cpu_s = time.perf_counter() - t0
        out.append({"cfg": "5-literal", "what": f"hessian! extended Rosenbrock n={n} f64 on the reference's reverse-over-reverse scalar tape",
                    "scalar_instructions": len(blk), "record_s": rec_s, "ms_per_hessian": ms, "rows_per_s": n / ms * 1e3,
                    "max_rel_err_vs_closed_form": err, "structural_zeros_exact": zeros_exact,
                    "cpu_port_rows_per_s": rows / cpu_s, "cpu_sample": f"{rows} rows, oracle/scalar_block.c loops, 1 thread",
                    "steps": stats_of(tape, lambda c: rd.hessian_(H, c, xd))})
        del ct, H
    for name, f, n in (("rosenbrock_loop", rd.workloads.f_rosenbrock_loop, 16384), ("ackley_loop", rd.workloads.f_ackley_loop, 16384)):
        tape = rd.GradientTape(f, rd.workloads.inputs_rosenbrock(n, seed=1))
        blk = tape.tape.program.instrs[0].scalar
        x = rd.workloads.inputs_rosenbrock(n, seed=2)
        xd = dev(x); g = torch.empty(n, dtype=torch.float64, device="cuda")
        res = {}
        for label, ug in (("graph", 1), ("eager", 0)):
            ctg = rd.compile(tape, ctx=rd.engine.Context(0) if ug else ctx, use_graph=ug)
            res["ms_per_eval_" + label] = timeit(lambda: rd.gradient_(g, ctg, xd), 5, 3)
        ot = orc.OracleTape(tape.program_bytes())
        ot.gradient([x])
        t0 = time.perf_counter()
        for _ in range(5):
            _, (go,) = ot.gradient([x])
        cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
        err = float(np.max(np.abs(g.cpu().numpy() - go)) / np.max(np.abs(go)))
        out.append({"cfg": "scalar-gradient", "what": f"gradient! of {name} n={n} f64 (one seed: only dataflow parallelism)",
                    "scalar_instructions": len(blk), **res, "cpu_port_ms_per_eval": cpu_ms, "max_rel_err_vs_oracle": err,
                    "steps": stats_of(tape, lambda c: rd.gradient_(g, c, xd))})

for o in out:
    print(json.dumps(o))
_ = time
