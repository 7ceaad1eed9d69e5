"""One int8-slice FP64 GEMM shape through rdb_gemm (for ncu captures and quick timings): MN, K from the environment."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
rd = graft.load_package()
ctx = rd.engine.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_gemm_f64_mode(2)
m = n = int(os.environ.get("MN", "4096")); k = int(os.environ.get("K", "4096")); reps = int(os.environ.get("REPS", "5"))
A = torch.rand(k, m, dtype=torch.float64, device="cuda"); B = torch.rand(n, k, dtype=torch.float64, device="cuda")
C = torch.zeros(n, m, dtype=torch.float64, device="cuda")
f = lambda: ctx.gemm(0, 0, 0, m, n, k, 1.0, A.data_ptr(), m, B.data_ptr(), k, 0.0, C.data_ptr(), m)
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"M=N={m} K={k}: {ms:.3f} ms per call (slicing included), {2.0*m*n*k/ms/1e9:.1f} TFLOP/s FP64-equivalent", flush=True)
