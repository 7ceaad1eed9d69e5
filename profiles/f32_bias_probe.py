"""Signed error of the FP32 `*` kernel (3xTF32 on tcgen05) on all-positive data: is the tensor core's truncating accumulation a
BIAS (same sign for every element) that survives a sum over the result?"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
rd = graft.load_package()
ctx = rd.engine.Context(0)
for n, k in ((1024, 1024), (2048, 2048), (4096, 4096), (1024, 16384)):
    for dist in ("uniform01", "normal"):
        rng = np.random.default_rng(1)
        gen = rng.random if dist == "uniform01" else rng.standard_normal
        A = np.asfortranarray(gen((n, k)).astype(np.float32)); B = np.asfortranarray(gen((k, n)).astype(np.float32))
        dA, dB = rd.engine.device_array(A), rd.engine.device_array(B)
        dC = torch.zeros(n * n, dtype=torch.float32, device="cuda")
        ctx.gemm(1, 0, 0, n, n, k, 1.0, dA.data_ptr(), n, dB.data_ptr(), k, 0.0, dC.data_ptr(), n)
        ctx.synchronize()
        C = dC.cpu().numpy().reshape(n, n, order="F").astype(np.float64)
        ref = (torch.from_numpy(A).cuda().double() @ torch.from_numpy(B).cuda().double()).cpu().numpy()
        err = C - ref
        scale = np.max(np.abs(ref))
        sg = torch.from_numpy(A).cuda() @ torch.from_numpy(B).cuda()        # cuBLAS SGEMM, for comparison
        err_s = sg.cpu().numpy().astype(np.float64) - ref
        print(f"n={n} k={k} {dist}: engine max|err|/max {np.max(np.abs(err))/scale:.2e} mean signed err/max {np.mean(err)/scale:+.2e} "
              f"sum rel err {abs(C.sum()-ref.sum())/abs(ref.sum()):.2e} | cuBLAS SGEMM max {np.max(np.abs(err_s))/scale:.2e} mean {np.mean(err_s)/scale:+.2e}", flush=True)
