"""Every BASELINE.json config at its BASELINE size, through the C ABI, in `pytest -m gpu`.

The oracle cannot run these sizes in seconds, so each check is an independent closed form evaluated in FP64 by a
DIFFERENT library than the engine's kernels (torch/cuBLAS on the device, or numpy O(n^2) formulas): standard two-layer backprop
with tanh' = 1 - y^2 for cfg 3, J = diag(t)(A+I) + diag(s(1-t^2))B for cfg 4, the 2x2 block-diagonal Hessian of the extended
Rosenbrock function for cfg 5 (structural zeros exact), rowsum/colsum gradients for cfg 2.  Tolerances: 1e-12 (FP64), 1e-5 (FP32),
relative to max|.| of each result array.  The reference's own stored literals (tests/golden/reference_literals.json) go through the
CUDA path with `==`, as the reference asserts them (LinAlgTests.jl:281-282, compat/CompatTests.jl:5-11).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_literals.json")))


def _rel(got, ref):
    return float((got - ref).abs().max() / ref.abs().max())


# ------------------------------------------------------------------------------------------ reference literals on the GPU
def test_reference_literals_through_cuda(rd):
    g = GOLD["issue235_vector_adjoint_times_matrix"]
    A, x = np.array(g["A"], float), np.array(g["x"], float)
    ct = rd.compile(rd.GradientTape(lambda y: rd.sum(y.T @ A), x))
    assert np.array_equal(rd.gradient(ct, x), np.array(g["gradient"], float))                  # LinAlgTests.jl:281-282 asserts ==
    for key, c in (("compat_abs2_minus_zeros", np.zeros(3)), ("compat_abs2_minus_range", np.arange(1.0, 4.0))):
        g = GOLD[key]
        x = np.array(g["x"], float)
        ct = rd.compile(rd.GradientTape(lambda v: rd.sum(rd.bcast(lambda a, b: rd.abs2(a - b), v, c)), x))
        assert np.array_equal(rd.gradient(ct, x), np.array(g["gradient"], float))              # CompatTests.jl:5-11


def test_reference_index_iterables_through_cuda(rd):
    """getindex with the index forms of TrackedTests.jl:641-716: the device gather/scatter must pick exactly the elements the
    reference's iterables enumerate (value = the picked elements, gradient = 1 at each picked element)."""
    keys = {"[:, :]": (slice(None), slice(None)), "[:, 1:2]": (slice(None), slice(0, 2)), "[2:3, :]": (slice(1, 3), slice(None)),
            "[1:2, 2:3]": (slice(0, 2), slice(1, 3)), "[2:6]": (slice(1, 6),), "[:]": (slice(None),)}
    v = np.asfortranarray(np.arange(1.0, 10.0).reshape(3, 3, order="F"))
    for case in GOLD["index_iterables_3x3"]["cases"]:
        k = keys[case["index"]]
        tape = rd.JacobianTape(lambda a: a[k], v)
        ct = rd.compile(tape)
        m = tape.output.size
        res = rd.DiffResult(None, np.zeros((m, 9), order="F"))
        rd.jacobian_(res, ct, v)
        if "expect_1based" in case:
            lin = [(i - 1) + (j - 1) * 3 for i, j in case["expect_1based"]]
        else:
            lin = [i - 1 for i in case["expect_1based_linear"]]
        want = np.zeros((m, 9)); want[np.arange(m), lin] = 1.0
        assert np.array_equal(res.jacobian, want)
        assert np.array_equal(np.asarray(res.value).reshape(-1, order="F"), v.reshape(-1, order="F")[lin])


# ------------------------------------------------------------------------------------------ cfg 2 at 4096^2, FP64 and FP32
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_cfg2_baseline_size(rd, dtype, tol):
    n = 4096
    rng = np.random.default_rng(1)
    a0 = np.asfortranarray(rng.random((n, n)).astype(dtype)); b0 = np.asfortranarray(rng.random((n, n)).astype(dtype))
    ct = rd.compile(rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0), dtype=dtype))
    a, b = rd.workloads.inputs_gradient_example(n, dtype)
    ga, gb = np.zeros((n, n), dtype, order="F"), np.zeros((n, n), dtype, order="F")
    res = (rd.DiffResult(0.0, ga), rd.DiffResult(0.0, gb))
    rd.gradient_(res, ct, (a, b))
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    one = np.ones(n)
    val = (a64 @ one) @ (b64 @ one) + (a64.T @ one) @ (b64.T @ one)
    assert abs(float(res[0].value) - val) < tol * abs(val)
    for g, other in ((ga, b64), (gb, a64)):
        ref = other.sum(1)[:, None] + other.sum(0)[None, :]
        assert np.max(np.abs(g.astype(np.float64) - ref)) < tol * np.max(np.abs(ref))
    if dtype == np.float64:
        gs = ct.handle.guard_stats()
        assert gs["checked"] >= 6 and gs["flagged"] == 0 and gs["fallbacks"] == 0, gs     # the int8-slice path ran, certified
        assert gs["worst_ratio"] < 9.1e-13


# ------------------------------------------------------------------------------------------ cfg 3 at 1024 x 1024 x 1024 x 65536
def test_cfg3_baseline_size(rd):
    """The only size that reaches the K = 65536 adjoints (dW = dZ * X' / dZ * H', four exact-int32 K-chunks each)."""
    import torch
    d = h = o = 1024; N = 65536
    X, rec = rd.workloads.inputs_mlp(d, h, o, N, seed=1)
    tape = rd.GradientTape(rd.workloads.make_f_mlp(X), rec)
    P = rd.program
    assert [i.opcode for i in tape.tape.program.instrs] == [P.OP_MUL, P.OP_NBCAST, P.OP_MUL, P.OP_NBCAST, P.OP_SUM]
    ct = rd.compile(tape)
    _, ws = rd.workloads.inputs_mlp(d, h, o, N, seed=2)
    res = tuple(rd.DiffResult(0.0, np.zeros(np.shape(w), order="F")) for w in ws)
    rd.gradient_(res, ct, ws)
    Xd = torch.from_numpy(X).cuda()
    W1, b1, W2, b2 = [torch.from_numpy(np.ascontiguousarray(w)).cuda() for w in ws]
    H = torch.tanh(W1 @ Xd + b1[:, None]); Y = torch.tanh(W2 @ H + b2[:, None])
    dZ2 = 1 - Y * Y
    dZ1 = (W2.T @ dZ2) * (1 - H * H)
    refs = [dZ1 @ Xd.T, dZ1.sum(1), dZ2 @ H.T, dZ2.sum(1)]
    val = float(Y.sum())
    assert abs(float(res[0].value) - val) < 1e-12 * abs(val)
    for r, ref, name in zip(res, refs, ("W1", "b1", "W2", "b2")):
        got = torch.from_numpy(np.ascontiguousarray(r.gradient)).cuda()
        assert _rel(got, ref) < 1e-12, name
    gs = ct.handle.guard_stats()
    # 2 forward products + dH (K = 1024, one launch each) + dW2 and dW1 (K = 65536: four exact-int32 K-chunks each)
    assert gs["checked"] >= 11 and gs["fallbacks"] == 0, gs


# ------------------------------------------------------------------------------------------ cfg 4 at n = 8192
def test_cfg4_baseline_size(rd):
    import torch
    n = 8192
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    ct = rd.compile(tape)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    J = torch.empty(n * n, dtype=torch.float64, device="cuda")
    rd.jacobian_(J, ct, rd.engine.device_array(x))
    Ad, Bd, xt = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(x).cuda()
    s = Ad @ xt + xt; t = torch.tanh(Bd @ xt)
    Jcf = t[:, None] * (Ad + torch.eye(n, dtype=torch.float64, device="cuda")) + (s * (1 - t * t))[:, None] * Bd
    assert _rel(J.reshape(n, n).T, Jcf) < 1e-12            # column-major (m x n) result viewed row-major
    gs = ct.handle.guard_stats()
    assert gs["checked"] >= 2 and gs["fallbacks"] == 0, gs     # the two 8192^3 adjoint products ran on int8 slices, certified
    # a shard of seed rows written into its place of the full column-major matrix (ld = m): what a rank of the multi-GPU run does
    J2 = torch.zeros(n * n, dtype=torch.float64, device="cuda")
    rb, re = 3 * n // 8, 4 * n // 8
    ct.handle.jacobian(0, rb, re, J2[rb:], (re - rb, n), ld=n)
    J2m = J2.reshape(n, n).T
    assert _rel(J2m[rb:re], Jcf[rb:re]) < 1e-12
    assert float(J2m[:rb].abs().max()) == 0.0 and float(J2m[re:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ cfg 5 at n = 16384
def test_cfg5_baseline_size(rd):
    import torch
    n = 16384
    tape = rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1))
    ct = rd.compile(tape)
    x = rd.workloads.inputs_rosenbrock(n, seed=2)
    xd = torch.from_numpy(x).cuda()
    H = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
    for _ in range(3):                                            # eager, capture, graph replay
        H.fill_(float("nan"))
        rd.hessian_(H, ct, xd)
        Hm = H.reshape(n, n).T                                    # H is symmetric; .T keeps (row, col) honest anyway
        o, e = xd[0::2], xd[1::2]
        i = torch.arange(n // 2, device="cuda")
        d_oo = 1200 * o * o - 400 * e + 2
        d_oe = -400 * o
        assert _rel(Hm[2 * i, 2 * i], d_oo) < 1e-12
        assert _rel(Hm[2 * i, 2 * i + 1], d_oe) < 1e-12 and _rel(Hm[2 * i + 1, 2 * i], d_oe) < 1e-12
        assert bool((Hm[2 * i + 1, 2 * i + 1] == 200.0).all())
        assert int(torch.count_nonzero(H)) == 2 * n               # every structural zero is an exact zero (and no NaN is left)
    # DiffResult form with host arrays at a size the PCIe copy allows: gradient and value ride along
    g = np.zeros(n); res = rd.DiffResult(0.0, g, np.zeros((n, 64), order="F"))
    rd.hessian_(res, ct, x, cols=(128, 192))
    oo, ee = x[0::2], x[1::2]
    gref = np.zeros(n); gref[0::2] = -400 * oo * (ee - oo * oo) - 2 * (1 - oo); gref[1::2] = 200 * (ee - oo * oo)
    assert np.max(np.abs(g - gref)) < 1e-12 * np.max(np.abs(gref))
    assert abs(res.value - np.sum(100 * (ee - oo ** 2) ** 2 + (1 - oo) ** 2)) < 1e-12 * abs(res.value)
    assert np.array_equal(res.hessian, Hm[:, 128:192].cpu().numpy())
