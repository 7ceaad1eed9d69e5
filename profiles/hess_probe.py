"""hessian! (extended Rosenbrock, device result, asynchronous calls) at n = 16384: time per call for the whole matrix and for the
column shards one rank owns at 2 / 4 / 8 GPUs - the per-call fixed cost is what bounds the multi-GPU scaling of this config."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
rd = graft.load_package()
n = int(os.environ.get("N", "16384"))
ctx = rd.engine.Context(0)
st = torch.cuda.Stream()
ctx.set_stream(st.cuda_stream)
ct = rd.compile(rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1)), ctx=ctx)
ct.handle.set_async(True)
xd = torch.from_numpy(rd.workloads.inputs_rosenbrock(n, seed=2)).cuda()
torch.cuda.synchronize()
base = None
with torch.cuda.stream(st):
    for world in (1, 2, 4, 8):
        cols = n // world
        H = torch.empty(n * cols, dtype=torch.float64, device="cuda")
        for _ in range(5):
            rd.hessian_(H, ct, xd, cols=(0, cols))
        torch.cuda.synchronize()
        reps = 400
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            rd.hessian_(H, ct, xd, cols=(0, cols))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        base = base or us
        print(f"shard 1/{world}: {us:8.1f} us per call, {n * cols * 8 / us / 1e3:7.1f} GB/s, speed-up {base / us:.2f}x", flush=True)
        del H
# reference points: a plain device fill of the same blocks (torch's fill kernel and cudaMemsetAsync), back to back
with torch.cuda.stream(st):
    for world in (1, 2, 4, 8):
        cols = n // world
        H = torch.empty(n * cols, dtype=torch.float64, device="cuda")
        for name, fn in (("torch zero_", lambda: H.zero_()), ("cudaMemsetAsync", lambda: torch.cuda.cudart().cudaMemsetAsync(H.data_ptr(), 0, H.numel() * 8, st.cuda_stream) if hasattr(torch.cuda.cudart(), "cudaMemsetAsync") else H.zero_())):
            for _ in range(5): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(400): fn()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 400 * 1e3
            print(f"{name} 1/{world}: {us:8.1f} us, {n * cols * 8 / us / 1e3:7.1f} GB/s", flush=True)
        del H
