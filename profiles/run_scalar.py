#!/usr/bin/env python
"""Short driver for ncu captures of the scalar-instruction kernels: one reference-style (reverse-over-reverse) Hessian of the
extended Rosenbrock function (seed-lane reverse kernel) and one gradient! of the loop-written function (slot-lane kernels)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

rd = graft.load_package()
ctx = rd.engine.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n = int(os.environ.get("SCALAR_HESS_N", "4096"))
reps = int(os.environ.get("SCALAR_REPS", "2"))
tape = rd.HessianTape(rd.workloads.f_rosenbrock_loop, rd.workloads.inputs_rosenbrock(n, seed=1), mode="reverse")
ct = rd.compile(tape, ctx=ctx)
x = torch.from_numpy(rd.workloads.inputs_rosenbrock(n, seed=2)).cuda()
H = torch.empty(n * n, dtype=torch.float64, device="cuda")
for _ in range(reps):
    rd.hessian_(H, ct, x)
torch.cuda.synchronize()
ng = int(os.environ.get("SCALAR_GRAD_N", "16384"))
gt = rd.GradientTape(rd.workloads.f_rosenbrock_loop, rd.workloads.inputs_rosenbrock(ng, seed=1))
cg = rd.compile(gt, ctx=ctx, use_graph=0)
xg = torch.from_numpy(rd.workloads.inputs_rosenbrock(ng, seed=2)).cuda()
g = torch.empty(ng, dtype=torch.float64, device="cuda")
for _ in range(reps):
    rd.gradient_(g, cg, xg)
torch.cuda.synchronize()
print("ok", float(H[0]), float(g[0]))
