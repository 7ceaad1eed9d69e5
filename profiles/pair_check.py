"""Quick correctness probe of the cta_group::2 int8-slice kernel through rdb_gemm (mode 2): a few shapes incl. ragged edges, all
transposes, beta != 0, against numpy; prints max error relative to |A||B|."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
rd = graft.load_package()
ctx = rd.engine.Context(0)
ctx.set_gemm_f64_mode(2)
rng = np.random.default_rng(0)
worst = 0.0
for (m, n, k) in [(256, 256, 256), (512, 384, 640), (1024, 1024, 1024), (257, 2049, 263), (300, 270, 1000), (4096, 4096, 512), (384, 64 * 5 + 7, 300)]:
    for ta, tb in [(0, 0), (1, 0), (0, 1), (1, 1)]:
        A = rng.standard_normal((k, m) if ta else (m, k)); B = rng.standard_normal((n, k) if tb else (k, n)); C0 = rng.standard_normal((m, n))
        dA, dB = rd.engine.device_array(A), rd.engine.device_array(B)
        for alpha, beta in [(1.0, 0.0), (-0.5, 2.0)]:
            dC = rd.engine.device_array(C0).contiguous() if False else torch.from_numpy(np.ascontiguousarray(C0.T)).cuda()
            ctx.gemm(0, ta, tb, m, n, k, alpha, dA.data_ptr(), A.shape[0], dB.data_ptr(), B.shape[0], beta, dC.data_ptr(), m)
            ctx.synchronize()
            got = dC.cpu().numpy().T
            opA = A.T if ta else A; opB = B.T if tb else B
            ref = alpha * (opA @ opB) + beta * C0
            scale = np.abs(opA) @ np.abs(opB) + np.abs(C0) * abs(beta)
            err = float(np.max(np.abs(got - ref) / scale))
            worst = max(worst, err)
            if err > 2e-13:
                print("MISMATCH", (m, n, k), (ta, tb), (alpha, beta), err, flush=True)
print("pair_check worst component-wise error", worst, "guard", ctx.gemm_guard(), flush=True)
