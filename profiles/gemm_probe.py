"""Time rdb_gemm (FP64 auto mode) at a few shapes with beta = 0 / 1 and dense / diagonal operands."""
import sys, json
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import __graft_entry__ as g
rd = g.load_package()
ctx = rd.engine.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
def run(n, beta, diag):
    A = torch.rand(n, n, dtype=torch.float64, device="cuda")
    B = torch.diag(torch.rand(n, dtype=torch.float64, device="cuda")) if diag else torch.rand(n, n, dtype=torch.float64, device="cuda")
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    f = lambda: ctx.gemm(0, 1, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, beta, C.data_ptr(), n)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return ms, 2 * n ** 3 / ms / 1e9
for n in (4096, 8192):
    for beta in (0.0, 1.0):
        for diag in (False, True):
            ms, tf = run(n, beta, diag)
            print(json.dumps({"n": n, "beta": beta, "diag_B": diag, "ms": round(ms, 3), "tflops_equiv": round(tf, 1)}))
