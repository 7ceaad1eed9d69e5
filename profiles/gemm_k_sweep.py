"""int8-slice FP64 GEMM at M = N = 4096 and a sweep of K: is the per-tile cost proportional to K (mainloop-bound) or does it carry
a fixed overhead (launch, TMEM alloc, pipeline fill, epilogue)?  Run under `ncu --metrics gpu__time_duration.sum` for kernel-only
durations, or plainly for CUDA-event times of the whole rdb_gemm call (slicing included)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
rd = graft.load_package()
ctx = rd.engine.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_gemm_f64_mode(2)
m = n = int(os.environ.get("MN", "4096"))
for k in (512, 1024, 2048, 4096, 8192, 16384):
    A = torch.rand(k, m, dtype=torch.float64, device="cuda")      # column-major m x k
    B = torch.rand(n, k, dtype=torch.float64, device="cuda")      # column-major k x n
    C = torch.zeros(n, m, dtype=torch.float64, device="cuda")
    for _ in range(2):
        ctx.gemm(0, 0, 0, m, n, k, 1.0, A.data_ptr(), m, B.data_ptr(), k, 0.0, C.data_ptr(), m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.gemm(0, 0, 0, m, n, k, 1.0, A.data_ptr(), m, B.data_ptr(), k, 0.0, C.data_ptr(), m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"K={k}: {ms:.3f} ms per call (slicing included), {2.0*m*n*k/ms/1e9:.1f} TFLOP/s FP64-equivalent", flush=True)
