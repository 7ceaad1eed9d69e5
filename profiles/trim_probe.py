"""Plane trimming of the int8-slice FP64 GEMM (umma.cu, planes_publish_* / oz2_issue_trimmed): operands whose low digit planes are
identically zero.  For every (bits of A, bits of B) pair: the result against cuBLAS DGEMM (exact when 2 b + log2 K <= 53: every
partial sum is representable, so both must agree bit for bit), the slice pairs the kernel issued and the time per call."""
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
g = torch.Generator(device="cuda"); g.manual_seed(3)


def operand(rows, cols, bits):
    """column-major rows x cols (torch shape (cols, rows)), entries in [0, 1] with `bits` fractional bits; 0 = dense doubles"""
    x = torch.rand(cols, rows, dtype=torch.float64, device="cuda", generator=g)
    if bits == 0:
        return x
    if bits < 0:
        return torch.ones_like(x)
    return torch.round(x * 2.0 ** bits) / 2.0 ** bits


def run(ta, tb, ba, bb):
    # op(A) m x k, op(B) k x n; stored column-major: A is (k x m if ta else m x k)
    A = operand(k if ta else m, m if ta else k, ba)
    B = operand(n if tb else k, k if tb else n, bb)
    C = torch.zeros(n, m, dtype=torch.float64, device="cuda")
    lda = k if ta else m; ldb = n if tb else k
    f = lambda: ctx.gemm(0, ta, tb, m, n, k, 1.0, A.data_ptr(), lda, B.data_ptr(), ldb, 0.0, C.data_ptr(), m)
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # a column-major r x c matrix X is the torch tensor X^T of shape (c, r)
    opA = A if ta else A.t()           # ta: stored k x m, torch (m, k) = op(A);  else stored m x k, torch (k, m) = op(A)^T
    opB = B if tb else B.t()           # tb: stored n x k, torch (k, n) = op(B);  else stored k x n, torch (n, k) = op(B)^T
    ref = opA @ opB
    got = C.t()                        # stored m x n => torch (n, m) = C^T
    err = float((got - ref).abs().max() / ref.abs().max())
    exact = bool(torch.equal(got, ref))
    return ms, err, exact


cases = [(0, 0), (-1, 0), (0, -1), (-1, -1), (8, 0), (0, 8), (16, 16), (20, 20), (24, 24), (24, 0), (40, 40)]
if os.environ.get("CASES"):
    cases = [tuple(int(v) for v in c.split(":")) for c in os.environ["CASES"].split(",")]
for ta, tb in ((0, 0), (1, 0), (0, 1)):
    for ba, bb in cases:
        ms, err, exact = run(ta, tb, ba, bb)
        print(f"ta={ta} tb={tb} bitsA={ba:3d} bitsB={bb:3d}: {ms:.3f} ms/call  rel.err {err:.2e}  bit-identical to DGEMM: {exact}  "
              f"{2.0 * m * n * k / ms / 1e9:.0f} TFLOP/s FP64-eq", flush=True)
