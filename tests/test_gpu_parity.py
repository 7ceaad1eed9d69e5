"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs.  Tolerances are the north star's: 1e-12 (FP64) / 1e-5 (FP32), relative to max|.| of each result array.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def run_gradient(rd, orc, f, rec_inputs, inputs, dtype=np.float64, check_buffers=True, **opts):
    tape = rd.GradientTape(f, tuple(rec_inputs), dtype=dtype)
    ct = rd.compile(tape, **opts)
    res = tuple(rd.DiffResult(0.0, np.zeros(np.shape(x), dtype, order="F")) for x in inputs)
    rd.gradient_(res, ct, tuple(inputs))
    ot = orc.OracleTape(tape.program_bytes())
    val, grads = ot.gradient(list(inputs))
    tol = TOL[dtype]
    assert abs(float(res[0].value) - float(val)) <= tol * max(abs(float(val)), 1e-300)
    for r, g in zip(res, grads):
        assert relerr(r.gradient, g) < tol
    if check_buffers:
        # every observable buffer: values of all instruction outputs, derivs of intermediates back to zero
        for b in range(len(ot.value)):
            if ot.kind[b] == orc.BUF_CONST:
                continue
            v = ct.handle.get_value(b)
            assert relerr(v, ot.value[b].reshape(-1, order="F")) < tol, f"value of buffer {b}"
            d = ct.handle.get_deriv(b)
            od = ot.deriv[b].reshape(-1, order="F")
            if np.any(od):        # tape inputs (and reshape aliases of them) hold the gradient ...
                assert relerr(d, od) < tol
            else:                 # ... every other deriv is back to zero (each reverse exec ends in unseed!)
                assert b not in ot.inputs or not np.any(od)
                assert not np.any(d), f"deriv of intermediate buffer {b} not unseeded"
    return ct


# ------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("dtype,mode", [(np.float64, 0), (np.float64, 1), (np.float32, 0)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (3, 5, 7), (64, 64, 64), (100, 100, 100), (129, 257, 65), (512, 384, 640),
                                   (1024, 1024, 1024), (8, 1, 300), (1030, 1026, 18), (300, 270, 1000), (257, 2049, 263),
                                   (300, 1, 200), (1, 300, 200), (4097, 1, 513), (1, 2000, 77)])
def test_gemm_kernel(rd, ctx, dtype, mode, ta, tb, m, n, k):
    """mode 0 = auto: FP64 shapes >= 256^3 go through exact int8 slices on tcgen05 (truncation 2^-56 of the row/column
    maxima instead of 2^-53 rounding), FP32 shapes >= 128 through 3xTF32 on tcgen05; mode 1 = DMMA only."""
    import torch
    ctx.set_gemm_f64_mode(mode)
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    A = rng.standard_normal((k, m) if ta else (m, k)).astype(dtype)
    B = rng.standard_normal((n, k) if tb else (k, n)).astype(dtype)
    C0 = rng.standard_normal((m, n)).astype(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    # column-major buffers on the device
    dA = torch.from_numpy(np.ascontiguousarray(A.T)).cuda(); dB = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
    for alpha, beta in [(1.0, 0.0), (1.0, 1.0), (-0.5, 2.0)]:
        dC = torch.from_numpy(np.ascontiguousarray(C0.T)).cuda()
        ctx.gemm(0 if dtype == np.float64 else 1, ta, tb, m, n, k, alpha, dA.data_ptr(), A.shape[0], dB.data_ptr(), B.shape[0],
                 beta, dC.data_ptr(), m)
        ctx.synchronize()
        got = dC.cpu().numpy().T
        opA = A.T if ta else A; opB = B.T if tb else B
        ref = alpha * (opA.astype(np.float64) @ opB.astype(np.float64)) + beta * C0.astype(np.float64)
        scale = np.abs(opA).astype(np.float64) @ np.abs(opB).astype(np.float64) + np.abs(C0) * abs(beta)
        err = np.max(np.abs(got - ref) / np.maximum(scale, 1e-300))
        f64_tol = 1e-14 if (mode == 1 or os.environ.get("RDB_F64_GEMM", "").startswith("d")) else 2e-13
        assert err < (f64_tol if dtype == np.float64 else 2e-6), (alpha, beta, err)
    ctx.set_gemm_f64_mode(0)
    _ = tdt


def test_gemm_int8_path_propagates_nonfinite(rd, ctx):
    """NaN / Inf in an operand row must poison the outputs it touches on the sliced tcgen05 path too (DGEMM semantics)."""
    import torch
    n = 512
    A = torch.rand(n, n, dtype=torch.float64, device="cuda"); B = torch.rand(n, n, dtype=torch.float64, device="cuda")
    A[7, 100] = float("nan")            # column-major view: element (row 100, col 7) of the stored matrix
    B[300, 5] = float("inf")            # element (row 5, col 300)
    C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    ctx.set_gemm_f64_mode(0)
    ctx.gemm(0, 0, 0, n, n, n, 1.0, A.data_ptr(), n, B.data_ptr(), n, 0.0, C.data_ptr(), n)
    ctx.synchronize()
    Cm = C.cpu().numpy().T               # back to (row, col)
    assert np.all(~np.isfinite(Cm[100, :]))      # row 100 of op(A) holds the NaN
    assert np.all(~np.isfinite(Cm[:, 300]))      # column 300 of op(B) holds the Inf
    mask = np.ones((n, n), bool); mask[100, :] = False; mask[:, 300] = False
    ref = (A.cpu().numpy().T @ B.cpu().numpy().T)
    assert np.all(np.isfinite(Cm[mask])) and np.max(np.abs(Cm[mask] - ref[mask])) < 1e-11


def test_gemm_beta0_ignores_nan(rd, ctx):
    import torch
    A = torch.ones(128, 128, dtype=torch.float64, device="cuda")
    C = torch.full((128, 128), float("nan"), dtype=torch.float64, device="cuda")
    ctx.gemm(0, 0, 0, 128, 128, 128, 1.0, A.data_ptr(), 128, A.data_ptr(), 128, 0.0, C.data_ptr(), 128)
    ctx.synchronize()
    assert torch.all(C == 128.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 7, 1000, 1 << 20, (1 << 20) + 3])
def test_add_sum_kernel(rd, ctx, dtype, n):
    import torch
    rng = np.random.default_rng(n)
    a = rng.random(n).astype(dtype); b = rng.random(n).astype(dtype)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty_like(da); s = torch.zeros(1, dtype=da.dtype, device="cuda")
    ctx.add_sum(0 if dtype == np.float64 else 1, n, da.data_ptr(), db.data_ptr(), 1.0, out.data_ptr(), s.data_ptr())
    ctx.synchronize()
    assert np.array_equal(out.cpu().numpy(), a + b)           # elementwise: bit-exact
    ref = np.sum((a + b).astype(np.float64))
    assert abs(float(s.item()) - ref) <= TOL[dtype] * ref


# ------------------------------------------------------------------------------------------ the five configs (small)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 5, 100, 257])
def test_cfg1_gradient(rd, orc, dtype, n):
    rng = np.random.default_rng(1)
    rec = (rng.random((n, n)).astype(dtype), rng.random((n, n)).astype(dtype))
    run_gradient(rd, orc, rd.workloads.f_gradient_example, rec, rd.workloads.inputs_gradient_example(n, dtype), dtype)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n", [256, 512])
def test_cfg2_both_gemm_paths_vs_oracle(rd, orc, mode, n):
    """cfg 2 at sizes where the oracle still runs in a second, through the tcgen05 int8-slice path (mode 0) and DMMA (mode 1)."""
    rng = np.random.default_rng(1)
    rec = (rng.random((n, n)), rng.random((n, n)))
    run_gradient(rd, orc, rd.workloads.f_gradient_example, rec, rd.workloads.inputs_gradient_example(n), gemm_f64_mode=mode)


def test_ozaki_signed_wide_range_inputs(rd, orc):
    """Mixed signs and a 1e6 spread of row scales: the int8-slice path is norm-wise exact like DGEMM."""
    n = 384
    rng = np.random.default_rng(4)
    scale = 10.0 ** rng.uniform(-3, 3, (n, 1))
    a = np.asfortranarray(rng.standard_normal((n, n)) * scale); b = np.asfortranarray(rng.standard_normal((n, n)) * scale.T)
    run_gradient(rd, orc, rd.workloads.f_gradient_example, (a, b), (a, b), check_buffers=True)


@pytest.mark.parametrize("fuse", [0, 1])
def test_cfg1_fusion_toggle_same_result(rd, orc, fuse):
    rng = np.random.default_rng(1)
    n = 65
    rec = (rng.random((n, n)), rng.random((n, n)))
    run_gradient(rd, orc, rd.workloads.f_gradient_example, rec, rd.workloads.inputs_gradient_example(n), fuse_elementwise=fuse)


def test_cfg1_replay_twice_new_inputs(rd, orc):
    """Record at one point, replay at others (GradientTests.jl:40-43); a tape is reusable."""
    n = 33
    rng = np.random.default_rng(1)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((n, n)), rng.random((n, n))))
    ct = rd.compile(tape)
    ot = orc.OracleTape(tape.program_bytes())
    for seed in (2, 3, 4):
        a, b = rd.workloads.inputs_gradient_example(n, seed=seed)
        ga, gb = rd.gradient(ct, (a, b))
        _, (oa, ob) = ot.gradient([a, b])
        assert relerr(ga, oa) < 1e-12 and relerr(gb, ob) < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(7, 5, 4, 9), (64, 48, 32, 200), (128, 128, 128, 1024)])
def test_cfg3_mlp(rd, orc, dtype, shape):
    d, h, o, N = shape
    X, rec = rd.workloads.inputs_mlp(d, h, o, N, dtype, seed=1)
    _, ws = rd.workloads.inputs_mlp(d, h, o, N, dtype, seed=2)
    run_gradient(rd, orc, rd.workloads.make_f_mlp(X), rec, ws, dtype)


@pytest.mark.parametrize("n,block", [(12, 0), (12, 5), (200, 64), (256, 0)])
def test_cfg4_jacobian(rd, orc, n, block):
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    ct = rd.compile(tape, seed_block=block)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    res = rd.DiffResult(None, np.zeros((n, n), order="F"))
    rd.jacobian_(res, ct, x)
    y, J = orc.OracleTape(tape.program_bytes()).jacobian([x])
    assert relerr(res.jacobian, J) < 1e-12
    assert relerr(res.value, y) < 1e-12
    # a shard of rows (the multi-GPU partition) equals the same rows of the full matrix
    part = np.zeros((5, n), order="F")
    rd.jacobian_(part, ct, x, rows=(3, 8))
    # (not bit-equal in general: a 5-seed block runs the DMMA kernel, a >=256-seed block the tcgen05 int8-slice kernel)
    assert relerr(part, res.jacobian[3:8, :]) < 1e-13


@pytest.mark.parametrize("n,block", [(2, 0), (10, 0), (10, 3), (64, 16), (1000, 0)])
def test_cfg5_hessian(rd, orc, n, block):
    x0 = rd.workloads.inputs_rosenbrock(n, seed=1)
    tape = rd.HessianTape(rd.workloads.f_rosenbrock, x0)
    ct = rd.compile(tape, seed_block=block)
    x = rd.workloads.inputs_rosenbrock(n, seed=2)
    H = np.full((n, n), np.nan, order="F")
    g = np.zeros(n)
    res = rd.DiffResult(0.0, g, H)
    rd.hessian_(res, ct, x)
    o, e = x[0::2], x[1::2]
    Hcf = np.zeros((n, n))
    idx = np.arange(n // 2)
    Hcf[2 * idx, 2 * idx] = 1200 * o ** 2 - 400 * e + 2
    Hcf[2 * idx, 2 * idx + 1] = Hcf[2 * idx + 1, 2 * idx] = -400 * o
    Hcf[2 * idx + 1, 2 * idx + 1] = 200
    assert relerr(H, Hcf) < 1e-12
    assert np.all(H[Hcf == 0] == 0.0)                           # structural zeros exact
    val, (og,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert relerr(g, og) < 1e-12 and abs(res.value - val) < 1e-12 * abs(val)
    if n >= 10:
        part = np.zeros((n, 4), order="F")
        rd.hessian_(part, ct, x, cols=(3, 7))
        assert np.array_equal(part, H[:, 3:7])


@pytest.mark.parametrize("n,block", [(64, 0), (64, 16), (1000, 0), (4096, 0)])
def test_cfg5_hessian_graph_replay(rd, n, block):
    """Device results without gradient/value: the second call with the same arguments captures the sweep as a CUDA graph, later
    calls replay it.  New inputs every call, a different column shard in between (the key changes: eager again), then the
    Hessian through a general tape (tangent sweeps instead of the sparse kernel)."""
    import torch
    def closed(x):
        o, e = x[0::2], x[1::2]
        Hcf = np.zeros((n, n)); idx = np.arange(n // 2)
        Hcf[2 * idx, 2 * idx] = 1200 * o ** 2 - 400 * e + 2
        Hcf[2 * idx, 2 * idx + 1] = Hcf[2 * idx + 1, 2 * idx] = -400 * o
        Hcf[2 * idx + 1, 2 * idx + 1] = 200
        return Hcf
    tape = rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1))
    ct = rd.compile(tape, seed_block=block)
    H = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
    Hs = torch.full((n * (n // 4),), float("nan"), dtype=torch.float64, device="cuda")
    for call in range(5):
        x = rd.workloads.inputs_rosenbrock(n, seed=10 + call)
        xd = torch.from_numpy(x).cuda()
        if call == 3:                                            # a shard: other key, runs eagerly, resets the capture
            Hs.fill_(float("nan"))
            rd.hessian_(Hs, ct, xd, cols=(n // 2, n // 2 + n // 4))
            assert np.array_equal(Hs.cpu().numpy().reshape(n, n // 4, order="F"), closed(x)[:, n // 2:n // 2 + n // 4]) or \
                relerr(Hs.cpu().numpy().reshape(n, n // 4, order="F"), closed(x)[:, n // 2:n // 2 + n // 4]) < 1e-12
            continue
        H.fill_(float("nan"))
        rd.hessian_(H, ct, xd)
        Hh = H.cpu().numpy().reshape(n, n, order="F")
        Hcf = closed(x)
        assert relerr(Hh, Hcf) < 1e-12, f"call {call}"
        assert np.all(Hh[Hcf == 0] == 0.0), f"call {call}"
    # the gradient path on the same tape still sees consistent state after graph replays
    g = np.zeros(n); Hn = np.zeros((n, n), order="F")
    x = rd.workloads.inputs_rosenbrock(n, seed=99)
    res = rd.DiffResult(0.0, g, Hn)
    rd.hessian_(res, ct, x)
    o, e = x[0::2], x[1::2]
    gref = np.zeros(n); gref[0::2] = -400 * o * (e - o * o) - 2 * (1 - o); gref[1::2] = 200 * (e - o * o)
    assert relerr(g, gref) < 1e-12 and relerr(Hn, closed(x)) < 1e-12


def test_hessian_vs_reference_algorithm(rd):
    """Forward-over-reverse on the array tape (engine) against the reference's reverse-over-reverse on a scalar tape
    (oracle/scalar_tape.py, literal restatement of api/tape.jl:285-293): same Hessian to 1e-12."""
    from oracle import scalar_tape as st
    from test_oracle import _rosen_scalar
    n = 16
    x = rd.workloads.inputs_rosenbrock(n, seed=2)
    ct = rd.compile(rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1)))
    H = rd.hessian(ct, x)
    _, _, Href = st.hessian_reverse_over_reverse(_rosen_scalar, x)
    assert relerr(H, np.array(Href)) < 1e-12
    # a tape with `*` by a constant matrix, +, and a two-argument closure of tanh/exp
    rng = np.random.default_rng(5)
    m = 7
    A = rng.random((m, m)) / m
    f = lambda v: rd.sum(rd.map_(rd.forward(lambda u, w: rd.tanh(u) * w + u * u * rd.exp(-w)), A @ v + v, v))

    def f_scalar(v):
        s = 0.0
        for i in range(m):
            u = v[i]
            for j in range(m):
                u = u + float(A[i, j]) * v[j]
            s = s + (st.tanh(u) * v[i] + u * u * st.exp(-v[i]))
        return s
    x = rng.random(m)
    H = rd.hessian(rd.compile(rd.HessianTape(f, rng.random(m))), x)
    _, _, Href = st.hessian_reverse_over_reverse(f_scalar, x)
    assert relerr(H, np.array(Href)) < 1e-12


def test_hessian_general_tape_vs_fd(rd, orc):
    """A tape with `*` by constants, +, elementwise closures: forward-over-reverse vs finite differences of the oracle."""
    rng = np.random.default_rng(5)
    n = 9
    A = rng.random((n, n)) / n
    f = lambda x: rd.sum(rd.map_(rd.forward(lambda u, v: rd.tanh(u) * v + u * u * rd.exp(-v)), A @ x + x, x))
    tape = rd.HessianTape(f, rng.random(n))
    ct = rd.compile(tape)
    x = rng.random(n)
    H = rd.hessian(ct, x)
    Hfd = orc.hessian_dense(tape.program_bytes(), x)
    assert np.max(np.abs(H - Hfd)) < 1e-6 * max(1.0, np.max(np.abs(Hfd)))
    assert np.max(np.abs(H - H.T)) < 1e-12 * np.max(np.abs(H))


@pytest.mark.parametrize("block", [0, 4])
def test_hessian_through_dot(rd, orc, block):
    """`dot` (reductions.jl:91-104) in the forward-over-reverse sweep: both operands tracked (bilinear: the cross terms
    delta tv(y), delta tv(x)), one tracked against a constant, the same buffer twice (x . x), and a dot of nonlinear images -
    against closed forms where there is one and finite differences of the oracle's gradient otherwise."""
    rng = np.random.default_rng(11)
    n = 10
    c = rng.random(n)
    A = rng.random((n, n))
    x = rng.random(n) + 0.2
    # 1. (x + c) . x  ->  H = 2 I
    tape = rd.HessianTape(lambda v: rd.dot(v + c, v), rng.random(n))        # (a scalar `+` of two dots would be a scalar instruction)
    H = rd.hessian(rd.compile(tape, seed_block=block), x)
    assert np.max(np.abs(H - 2 * np.eye(n))) < 1e-13
    # 2. (A x) . x  ->  H = A + A'
    tape = rd.HessianTape(lambda v: rd.dot(A @ v, v), rng.random(n))
    H = rd.hessian(rd.compile(tape, seed_block=block), x)
    assert relerr(H, A + A.T) < 1e-13
    # 3. tanh.(x) . exp.(x[perm] / 2): cross terms between nonlinear images
    perm = [int(p) for p in rng.permutation(n)]
    f = lambda v: rd.dot(rd.bcast(lambda u: rd.tanh(u), v), rd.bcast(lambda u: rd.exp(u * 0.5), v[perm]))
    tape = rd.HessianTape(f, rng.random(n))
    ct = rd.compile(tape, seed_block=block)
    H = rd.hessian(ct, x)
    Hfd = orc.hessian_dense(tape.program_bytes(), x)
    assert np.max(np.abs(H - Hfd)) < 1e-6 * max(1.0, np.max(np.abs(Hfd)))
    assert np.max(np.abs(H - H.T)) < 1e-12 * np.max(np.abs(H))
    t, e = np.tanh(x), np.exp(0.5 * x)
    Href = np.zeros((n, n))
    for i in range(n):                      # sum_i tanh(x_i) exp(x_perm[i] / 2)
        j = perm[i]
        Href[i, i] += -2 * t[i] * (1 - t[i] ** 2) * e[j]
        Href[j, j] += t[i] * 0.25 * e[j]
        Href[i, j] += (1 - t[i] ** 2) * 0.5 * e[j]
        Href[j, i] += (1 - t[i] ** 2) * 0.5 * e[j]
    assert relerr(H, Href) < 1e-12
    # the gradient that rides along
    res = rd.DiffResult(0.0, np.zeros(n), np.zeros((n, n), order="F"))
    rd.hessian_(res, ct, x)
    _, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert relerr(res.gradient, g) < 1e-12 and relerr(res.hessian, Href) < 1e-12


@pytest.mark.parametrize("dims", [1, 2, (1, 2)])
@pytest.mark.parametrize("block", [0, 5])
def test_hessian_through_sum_dims(rd, orc, dims, block):
    """`sum(x, dims)` (reductions.jl:39-46) in the forward-over-reverse sweep: column sums, row sums and the full reduction of a
    nonlinear image, a nonlinear closure on top - against the closed form f = sum_g tanh(s_g), s_g = sum of x^2 over group g."""
    rng = np.random.default_rng(13)
    r, c = 4, 3
    n = r * c
    f = lambda v: rd.sum(rd.bcast(lambda u: rd.tanh(u), rd.sum(rd.bcast(lambda u: u * u, v.reshape(r, c)), dims=dims)))
    tape = rd.HessianTape(f, rng.random(n))
    ct = rd.compile(tape, seed_block=block)
    x = rng.random(n) * 0.8 + 0.1
    H = rd.hessian(ct, x)
    X = x.reshape(r, c, order="F")
    grp = np.zeros((r, c), dtype=int)                    # group id of every element
    if dims == 1: grp[:] = np.arange(c)[None, :]
    elif dims == 2: grp[:] = np.arange(r)[:, None]
    g = grp.reshape(-1, order="F")
    s = np.array([np.sum(x[g == k] ** 2) for k in range(g.max() + 1)])
    t = np.tanh(s)
    Href = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if g[i] != g[j]: continue
            k = g[i]
            Href[i, j] = -2 * t[k] * (1 - t[k] ** 2) * 4 * x[i] * x[j] + (2 * (1 - t[k] ** 2) if i == j else 0.0)
    assert relerr(H, Href) < 1e-12
    assert np.max(np.abs(H - orc.hessian_dense(tape.program_bytes(), x))) < 1e-6
    _ = X


@pytest.mark.parametrize("side", ["left", "right"])
@pytest.mark.parametrize("block", [0, 5])
def test_hessian_through_transposed_tracked_operand(rd, orc, side, block):
    """`*` whose tracked operand is an Adjoint (X' * C and C * X'): tangents and adjoint tangents are stored untransposed and the
    products read / write them through the transpose flags, as the first-order sweep does (arithmetic.jl:272-285)."""
    rng = np.random.default_rng(17)
    r, c, m = 4, 3, 5
    n = r * c
    if side == "left":
        C = rng.random((r, m)) - 0.5
        f = lambda v: rd.sum(rd.bcast(lambda u: rd.tanh(u), v.reshape(r, c).T @ C))          # Z = X' C  (c x m)
    else:
        C = rng.random((m, c)) - 0.5
        f = lambda v: rd.sum(rd.bcast(lambda u: rd.tanh(u), C @ v.reshape(r, c).T))          # Z = C X'  (m x r)
    tape = rd.HessianTape(f, rng.random(n))
    ct = rd.compile(tape, seed_block=block)
    x = rng.random(n) - 0.3
    H = rd.hessian(ct, x)
    X = x.reshape(r, c, order="F")
    Z = X.T @ C if side == "left" else C @ X.T
    t = np.tanh(Z); d2 = -2 * t * (1 - t * t)
    Href = np.zeros((n, n))
    lin = lambda i, j: i + j * r
    for i in range(r):
        for j in range(c):
            for i2 in range(r):
                for j2 in range(c):
                    if side == "left":        # Z[j, k] = sum_i X[i, j] C[i, k]
                        if j == j2: Href[lin(i, j), lin(i2, j2)] = np.sum(d2[j, :] * C[i, :] * C[i2, :])
                    else:                     # Z[k, i] = sum_j C[k, j] X[i, j]
                        if i == i2: Href[lin(i, j), lin(i2, j2)] = np.sum(d2[:, i] * C[:, j] * C[:, j2])
    assert relerr(H, Href) < 1e-12
    res = rd.DiffResult(0.0, np.zeros(n), np.zeros((n, n), order="F"))
    rd.hessian_(res, ct, x)
    _, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert relerr(res.gradient, g) < 1e-12


# ------------------------------------------------------------------------------------------ per-instruction
def _cases(rd):
    from test_oracle import CASES
    return CASES


@pytest.mark.parametrize("name", ["a*b", "a'*b", "a*b'", "a'*b'", "a+b", "a-b", "-a", "bcast*", "bcast-", "map", "nbcast_fns",
                                  "getindex", "reshape", "bcast/", "bcast\\", "bcast^", "dot", "sum_dims1", "sum_dims2", "tracker_real",
                                  "tracker_trackedreal", "getindex_generic", "getindex_cartesian"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_instruction(rd, orc, name, dtype):
    rng = np.random.default_rng(11)
    for shp in [(3, 3), (40, 40)]:
        if name in ("reshape",) and shp != (3, 3):
            continue
        rec = (rng.random(shp).astype(dtype), rng.random(shp).astype(dtype))
        ins = (np.asfortranarray(rng.random(shp).astype(dtype)), np.asfortranarray(rng.random(shp).astype(dtype)))
        run_gradient(rd, orc, _cases(rd)[name](rd), rec, ins, dtype)


SHAPE_CASES = {
    # rectangular and vector products, the combinations LinAlgTests.jl:210-278 walks through on 3x3 / length-3 inputs
    "rect a*b": ((3, 5), (5, 2), lambda rd: lambda a, b: rd.sum(a @ b)),
    "rect a'*b": ((5, 3), (5, 2), lambda rd: lambda a, b: rd.sum(a.T @ b)),
    "rect a*b'": ((3, 5), (2, 5), lambda rd: lambda a, b: rd.sum(a @ b.T)),
    "mat*vec": ((4, 6), (6,), lambda rd: lambda a, x: rd.sum(a @ x)),
    "mat'*vec": ((6, 4), (6,), lambda rd: lambda a, x: rd.sum(a.T @ x)),
    "vec'*mat": ((6, 4), (6,), lambda rd: lambda a, x: rd.sum(x.T @ a)),
    "outer x*y'": ((5,), (7,), lambda rd: lambda x, y: rd.sum(rd.bcast(lambda u: u * u, x.reshape(5, 1) @ y.T))),
    "1x1": ((1, 1), (1, 1), lambda rd: lambda a, b: rd.sum(a.T @ b + a @ b.T)),
    "big mat*vec": ((300, 200), (200,), lambda rd: lambda a, x: rd.sum(rd.bcast(lambda u: rd.tanh(u * 0.01), a @ x))),
    "big vec'*mat": ((300, 200), (300,), lambda rd: lambda a, x: rd.mean(x.T @ a)),
}


@pytest.mark.parametrize("name", sorted(SHAPE_CASES))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_shapes(rd, orc, name, dtype):
    sa, sb, mk = SHAPE_CASES[name]
    rng = np.random.default_rng(21)
    rec = (rng.random(sa).astype(dtype), rng.random(sb).astype(dtype))
    ins = tuple(np.asfortranarray(rng.random(s).astype(dtype)) for s in (sa, sb))
    run_gradient(rd, orc, mk(rd), rec, ins, dtype)


def test_jacobian_through_every_reduction_kind(rd, orc):
    """Batched seed rows (R > 1) through dot / sum(x,dims) / scalar-argument broadcasts / getindex adjoints."""
    rng = np.random.default_rng(12)
    n = 6
    A = rng.random((n, n))

    def f(x):
        X = A @ x.reshape(n, 1) @ x.reshape(1, n)                     # n x n, tracked on both sides
        rows = rd.sum(X, dims=2)                                      # n x 1
        s = rd.dot(x, x)                                              # TrackedReal
        y = rd.bcast(lambda u, c: rd.tanh(u) * c, rows.reshape(n), s) # tracker_∇broadcast with a TrackedReal argument
        return rd.broadcast("/", y, rd.bcast(lambda v: v + 2.0, x[[5, 4, 3, 2, 1, 0]]))
    x0 = rng.random(n)
    tape = rd.JacobianTape(f, x0)
    x = rng.random(n)
    for block in (0, 4):
        ct = rd.compile(tape, seed_block=block)
        J = rd.jacobian(ct, x)
        y, Jo = orc.OracleTape(tape.program_bytes()).jacobian([x])
        assert relerr(J, Jo) < 1e-12


def test_tuple_input_jacobian_and_diffresult(rd, orc):
    """jacobian! with a tuple of inputs repeats the seed loop per input (api/utils.jl:66-71)."""
    rng = np.random.default_rng(8)
    n = 9
    A = rng.random((n, n))
    f = lambda x, y: rd.bcast(lambda u, v: u * rd.tanh(v), A @ x + y, y)
    x0, y0 = rng.random(n), rng.random(n)
    tape = rd.JacobianTape(f, (x0, y0))
    ct = rd.compile(tape)
    x, y = rng.random(n), rng.random(n)
    Jx, Jy = rd.jacobian(ct, (x, y))
    ot = orc.OracleTape(tape.program_bytes())
    _, Jx_o = ot.jacobian([x, y], slot=0)
    _, Jy_o = ot.jacobian([x, y], slot=1)
    assert relerr(Jx, Jx_o) < 1e-12 and relerr(Jy, Jy_o) < 1e-12


def test_broadcast_patterns(rd, orc):
    rng = np.random.default_rng(3)
    W, b, r = rng.random((40, 60)), rng.random(40), rng.random((1, 60))
    f = lambda W_, b_, r_: rd.sum(rd.bcast(lambda x, y, z: rd.tanh(x + y) * z, W_, b_, r_))
    run_gradient(rd, orc, f, (W, b, r), (np.asfortranarray(rng.random((40, 60))), rng.random(40), np.asfortranarray(rng.random((1, 60)))))


def test_general_index_map_gather_scatter(rd, orc):
    """An IDX operand that is NOT a pure transpose (row-vector adjoint of a vector times a matrix exercises the
    folded path; a hand-built permuted map exercises gather + scatter-add)."""
    P = rd.program
    rng = np.random.default_rng(9)
    n = 6
    a0, b0 = rng.random((n, n)), rng.random((n, n))
    tape = rd.GradientTape(lambda a, b: rd.sum(a.T @ b), (a0, b0))
    op = tape.tape.program.instrs[0].inputs[0]
    perm = rng.permutation(n * n).astype(np.int64)
    perm[0] = perm[1]                                   # a duplicate: both elements must add
    op.idx_map = perm
    ct = rd.compile(tape)
    a, b = np.asfortranarray(rng.random((n, n))), np.asfortranarray(rng.random((n, n)))
    ga, gb = rd.gradient(ct, (a, b))
    _, (oa, ob) = orc.OracleTape(tape.program_bytes()).gradient([a, b])
    assert relerr(ga, oa) < 1e-12 and relerr(gb, ob) < 1e-12
    _ = P


def test_loader_rejects_bad_programs(rd, ctx):
    tape = rd.GradientTape(lambda v: rd.sum(v + v), np.ones(3))
    raw = bytearray(tape.program_bytes())
    with pytest.raises(rd.engine.EngineError):
        rd.engine.EngineTape(ctx, bytes(raw[:-8]))                        # truncated
    bad = bytearray(raw); bad[0] ^= 0xFF
    with pytest.raises(rd.engine.EngineError):
        rd.engine.EngineTape(ctx, bytes(bad))                             # bad magic
    prog = rd.program.TapeProgram.from_bytes(bytes(raw))
    prog.instrs[0].opcode = 99                                            # e.g. an @grad closure instruction
    with pytest.raises(rd.engine.EngineError, match="cannot be replayed"):
        rd.engine.EngineTape(ctx, prog.to_bytes())


def test_input_shape_errors(rd):
    tape = rd.GradientTape(lambda v: rd.sum(v + v), np.ones(3))
    ct = rd.compile(tape)
    with pytest.raises(ValueError):
        rd.gradient_(np.zeros(3), ct, np.ones(4))


# ------------------------------------------------------------------------------------------ host-array paths at panel sizes
PANEL_CASES = {
    # gradient! with HOST inputs/results takes extra routes once a matrix reaches 32 MB and 2048 columns: inputs upload in column
    # panels (the first `*` that reads the last-uploaded input as its plain right operand starts per panel), and the last one or
    # two `*` accumulations into an input run as (interleaved) column panels with early D2H copies.  Each case pins one route.
    "a'*b + a*b'": lambda rd: lambda a, b: rd.sum(a.T @ b + a @ b.T),          # folded/plain pairs on both inputs (cfg 2)
    "a*b + b*a": lambda rd: lambda a, b: rd.sum(a @ b + b @ a),                  # plain/plain pairs
    "a*a": lambda rd: lambda a, b: rd.sum(a @ a + b),                            # both accumulations in ONE instruction: no partner
    "(a*b)*a": lambda rd: lambda a, b: rd.sum((a @ b) @ a),                      # partner's output adjoint not final yet: no hoist
    "tanh.(a)*b": lambda rd: lambda a, b: rd.sum(rd.bcast(lambda u: rd.tanh(u), a) @ b),   # streamed b; a's first reader is not a `*`
    "b*a (b uploaded last, left operand)": lambda rd: lambda a, b: rd.sum(b @ a),           # last upload is a LEFT operand: settle
}


@pytest.mark.parametrize("name", sorted(PANEL_CASES))
def test_host_arrays_at_panel_sizes(rd, orc, name):
    n = 2048                                                    # 2048 x 2048 FP64 = 32 MiB: the smallest size that takes the routes
    rng = np.random.default_rng(5)
    rec = (rng.random((n, n)), rng.random((n, n)))
    ins = (np.asfortranarray(rng.random((n, n)) / n), np.asfortranarray(rng.random((n, n))))
    f = PANEL_CASES[name](rd)
    ct = run_gradient(rd, orc, f, rec, ins, check_buffers=False)
    # the same compiled tape with DEVICE arrays takes none of the routes: results must agree to the last bit (the int8-slice
    # GEMM is exact integer accumulation per element, so cutting it into panels cannot change a single result)
    import torch
    dev_in = tuple(rd.engine.device_array(x) for x in ins)
    dev_out = tuple(torch.empty(n * n, dtype=torch.float64, device="cuda") for _ in ins)
    rd.gradient_(dev_out, ct, dev_in)
    host_out = tuple(np.zeros((n, n), order="F") for _ in ins)
    rd.gradient_(host_out, ct, ins)
    for h, d in zip(host_out, dev_out):
        assert np.array_equal(h, d.cpu().numpy().reshape(n, n, order="F"))
    # and again with fresh inputs on the same tape (events, panel state and slice caches must not leak between calls)
    ins2 = (np.asfortranarray(rng.random((n, n)) / n), np.asfortranarray(rng.random((n, n))))
    res2 = tuple(np.zeros((n, n), order="F") for _ in ins2)
    rd.gradient_(res2, ct, ins2)
    _, g2 = orc.OracleTape(rd.GradientTape(f, rec).program_bytes()).gradient(list(ins2))
    for r, g in zip(res2, g2):
        assert relerr(r, g) < 1e-12


def test_host_arrays_panels_with_k_chunks(rd, orc):
    """Panelled adjoint GEMMs whose contraction is longer than one exact-int32 K-chunk (K = 20032 > 18688): the operand windows
    are dropped for the chunked product (its tag names the whole operand) and every chunk slices its own panel."""
    import torch
    m, k, n = 20032, 2048, 2048
    rng = np.random.default_rng(7)
    rec = (rng.random((m, k)), rng.random((k, n)))
    ins = (np.asfortranarray(rng.random((m, k)) / k), np.asfortranarray(rng.random((k, n))))
    f = lambda a, b: rd.sum(a @ b)
    ct = run_gradient(rd, orc, f, rec, ins, check_buffers=False)
    dev_in = tuple(rd.engine.device_array(x) for x in ins)
    dev_out = tuple(torch.empty(x.size, dtype=torch.float64, device="cuda") for x in ins)
    rd.gradient_(dev_out, ct, dev_in)
    host_out = tuple(np.zeros(x.shape, order="F") for x in ins)
    rd.gradient_(host_out, ct, ins)
    for h, d in zip(host_out, dev_out):
        assert relerr(h, d.cpu().numpy().reshape(h.shape, order="F")) < 1e-13


def test_host_arrays_at_panel_sizes_f32(rd, orc):
    """FP32 takes the same host-array routes (3xTF32 tcgen05 GEMMs cut into column panels, panelled uploads) from 32 MiB on."""
    import torch
    n = 3072                                                    # 3072 x 3072 FP32 = 36 MiB
    rng = np.random.default_rng(6)
    rec = (rng.random((n, n)).astype(np.float32), rng.random((n, n)).astype(np.float32))
    ins = (np.asfortranarray((rng.random((n, n)) / n).astype(np.float32)), np.asfortranarray(rng.random((n, n)).astype(np.float32)))
    ct = run_gradient(rd, orc, rd.workloads.f_gradient_example, rec, ins, dtype=np.float32, check_buffers=False)
    dev_in = tuple(rd.engine.device_array(x) for x in ins)
    dev_out = tuple(torch.empty(n * n, dtype=torch.float32, device="cuda") for _ in ins)
    rd.gradient_(dev_out, ct, dev_in)
    host_out = tuple(np.zeros((n, n), np.float32, order="F") for _ in ins)
    rd.gradient_(host_out, ct, ins)
    for h, d in zip(host_out, dev_out):
        assert relerr(h, d.cpu().numpy().reshape(n, n, order="F")) < 1e-6


# ------------------------------------------------------------------------------------------ full-size properties
def test_cfg2_full_size_closed_form(rd):
    """4096x4096 FP64: the oracle needs ~3 s of DGEMM per eval, so the full-size check uses the size-independent
    closed form — O(n^2), shares no GEMM with the engine."""
    n = 4096
    rng = np.random.default_rng(1)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((8, 8)), rng.random((8, 8))))
    _ = tape
    a0 = np.asfortranarray(rng.random((n, n))); b0 = np.asfortranarray(rng.random((n, n)))
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0))
    ct = rd.compile(tape)
    a, b = rd.workloads.inputs_gradient_example(n)
    ga, gb = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
    res = (rd.DiffResult(0.0, ga), rd.DiffResult(0.0, gb))
    rd.gradient_(res, ct, (a, b))
    one = np.ones(n)
    val = (a @ one) @ (b @ one) + (a.T @ one) @ (b.T @ one)
    assert abs(res[0].value - val) < 1e-12 * abs(val)
    assert relerr(ga, b.sum(1)[:, None] + b.sum(0)[None, :]) < 1e-12
    assert relerr(gb, a.sum(1)[:, None] + a.sum(0)[None, :]) < 1e-12


@pytest.mark.parametrize("env", [{"RDB_OZAKI_BITS": "7"},                                   # first-generation digits: 8 slices of 7 bits
                                 {"RDB_PRESLICE": "0", "RDB_STREAM_INPUTS": "0"},              # no side stream, no panelled uploads
                                 {"RDB_OZAKI_CLUSTER": "1"}])                                  # single-CTA tcgen05 kernel
def test_gemm_path_variants_in_child_process(env):
    """The GEMM-path knobs are read once per process: rerun the cfg 2 and panel-size tests with each alternative forced on."""
    import subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", "(cfg2 or panel_sizes) and not child_process", "-p", "no:cacheprovider"],
                       cwd=os.path.dirname(here), env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("env,select", [({"RDB_HESS_FUSED": "0"}, "cfg5_hessian"),          # hessian!: the general graph-replayed sweep
                                        ({"RDB_OZAKI_PAIR": "1"}, "cfg2 or panel_sizes"),     # cta_group::2 CTA-pair int8-slice kernel
                                        ({"RDB_OZAKI_PERSISTENT": "2"}, "cfg2 or panel_sizes or cfg3")])   # persistent cluster kernel for every K
def test_alternative_paths_in_child_process(env, select):
    import subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", f"({select}) and not child_process", "-p", "no:cacheprovider"],
                       cwd=os.path.dirname(here), env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_hessian_separable_fast_path_shapes(rd, orc):
    """The one-kernel hessian! (device result, nothing else requested) against the general sweep (host result) on the shapes the
    planner accepts: views with a repeated index (two contributions into one entry), `mean`, and a closure of x itself."""
    import torch
    E, T = rd.expr, rd.tracked
    n = 64
    x0 = np.random.default_rng(3).random(n) + 0.5
    x = np.random.default_rng(4).random(n) + 0.5
    ia = [0, 5, 5, 9, 63, 2]; ib = [1, 5, 7, 9, 0, 2]

    def f_views(v):
        g = T.forward(lambda a, b: a * a * b + E.exp(a * b))
        return T.mean(T.map_(g, v[ia], v[ib]))

    def f_self(v):
        return T.sum(T.bcast(lambda a: E.sin(a) * a ** 3, v))

    for f in (f_views, f_self):
        ct = rd.compile(rd.HessianTape(f, x0))
        Hh = np.zeros((n, n), order="F")
        rd.hessian_(Hh, ct, x)                                    # host result: general sweep
        Hd = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
        rd.hessian_(Hd, ct, torch.from_numpy(x).cuda())          # device result: fused kernel
        Hf = Hd.cpu().numpy().reshape(n, n, order="F")
        assert relerr(Hf, Hh) < 1e-13
        assert np.all(Hf[Hh == 0] == 0.0) and np.allclose(Hf, Hf.T, rtol=0, atol=1e-13 * np.max(np.abs(Hf)))
        part = torch.full((n * 16,), float("nan"), dtype=torch.float64, device="cuda")
        rd.hessian_(part, ct, torch.from_numpy(x).cuda(), cols=(40, 56))
        assert np.array_equal(part.cpu().numpy().reshape(n, 16, order="F"), Hf[:, 40:56])


def test_hessian_async_reads_device_input_in_place(rd):
    """rdb_tape_set_async + the one-kernel hessian!: a device input is read in place (no copy into the tape), calls are only
    enqueued, and an entry point that needs the tape's own copy of the input refuses to run on a stale one."""
    import torch
    n = 512
    ctx = rd.engine.Context(0)
    ct = rd.compile(rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1)), ctx=ctx)

    def closed(x):
        o, e = x[0::2], x[1::2]
        Hcf = np.zeros((n, n)); idx = np.arange(n // 2)
        Hcf[2 * idx, 2 * idx] = 1200 * o ** 2 - 400 * e + 2
        Hcf[2 * idx, 2 * idx + 1] = Hcf[2 * idx + 1, 2 * idx] = -400 * o
        Hcf[2 * idx + 1, 2 * idx + 1] = 200
        return Hcf
    xs = [rd.workloads.inputs_rosenbrock(n, seed=20 + k) for k in range(3)]
    xd = [torch.from_numpy(x).cuda() for x in xs]
    Hd = [torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda") for _ in xs]
    rd.hessian_(Hd[0], ct, xd[0])                                # synchronous first call: plans the fast path
    ct.handle.set_async(True)
    for k in (1, 2):
        rd.hessian_(Hd[k], ct, xd[k])                            # enqueued only
    ctx.synchronize()
    for k in range(3):
        Hh = Hd[k].cpu().numpy().reshape(n, n, order="F")
        assert relerr(Hh, closed(xs[k])) < 1e-12 and np.all(Hh[closed(xs[k]) == 0] == 0.0)
    ct.handle.set_async(False)
    with pytest.raises(Exception):                               # the tape never received a copy of xd[2]
        ct.handle.forward()
    g = np.zeros(n); res = rd.DiffResult(0.0, g, np.zeros((n, n), order="F"))
    rd.hessian_(res, ct, xs[1])                                  # a fresh set_input: everything works again
    assert relerr(res.hessian, closed(xs[1])) < 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_elementwise_run_fusion_is_bit_identical(rd, orc, dtype):
    """North star (a): a run of adjacent same-shape elementwise instructions replays as ONE forward kernel, and the `sum` behind
    the run and the reverse {sum, +} pair are fused too.  Every instruction's value and every deriv must be bit-identical to the
    one-kernel-per-instruction replay (`fuse_elementwise=0`), and agree with the oracle."""
    E, T = rd.expr, rd.tracked
    rng = np.random.default_rng(5)
    m, n = 96, 160
    A0, B0 = rng.random((m, n)).astype(dtype), rng.random((m, n)).astype(dtype)
    c0 = rng.random(m).astype(dtype)
    A, B = (rng.random((m, n)) - 0.5).astype(dtype), (rng.random((m, n)) - 0.5).astype(dtype)
    c = (rng.random(m) - 0.5).astype(dtype)

    def f(a, b, bias):
        t = T.broadcast(T.forward(E.tanh), b)                      # unfused spelling of cfg 4: broadcast(tanh) ...
        y = T.broadcast("*", a, t)                                 # ... then (broadcast, *) on its output
        z = T.bcast(lambda u, w: E.exp(u * 0.5) + w, y, bias)      # joins the run; `bias` broadcasts along columns (loaded, clamped)
        q = T.broadcast("-", a, b)                                 # an independent instruction of the same shape joins as well
        return T.sum(z + q)                                        # array + then sum: forward {+, sum} and reverse {sum, +}

    def f2(a, b, bias):
        # `sum` right behind an elementwise instruction with a same-shape and a column-vector argument: forward {elementwise, sum},
        # reverse {sum, elementwise adjoint, bias column-reduce} in one pass each
        h = T.bcast(lambda u, w: E.tanh(u + w), a, bias)
        return T.sum(T.bcast(lambda u, v, w: u * v + E.exp(w * 0.25), h, b, bias))

    for fn in (f, f2):
        _check_fused_against_unfused(rd, orc, fn, (A0, B0, c0), (A, B, c), dtype)


def _check_fused_against_unfused(rd, orc, f, rec, ins, dtype):
    A0, B0, c0 = rec
    A, B, c = ins
    tape = rd.GradientTape(f, (A0, B0, c0), dtype=dtype)
    ot = orc.OracleTape(tape.program_bytes())
    val, grads = ot.gradient([A, B, c])
    got = {}
    for fuse in (1, 0):
        ct = rd.compile(tape, fuse_elementwise=fuse)
        res = tuple(rd.DiffResult(0.0, np.zeros(np.shape(x), dtype, order="F")) for x in (A, B, c))
        l0 = ct.handle.launch_count()
        rd.gradient_(res, ct, (A, B, c))
        launches = ct.handle.launch_count() - l0
        nb = len(ot.value)
        got[fuse] = ([np.asarray(r.gradient).copy() for r in res], float(res[0].value),
                     [ct.handle.get_value(b).copy() for b in range(nb) if ot.kind[b] != orc.BUF_CONST], launches)
        for r, g in zip(res, grads):
            assert relerr(r.gradient, g) < TOL[dtype]
    for x, y in zip(got[1][0], got[0][0]):
        assert np.array_equal(x, y)
    assert got[1][1] == got[0][1]
    for x, y in zip(got[1][2], got[0][2]):
        assert np.array_equal(x, y)
    assert got[1][3] < got[0][3]                                  # and it really took fewer launches


def test_gradient_pipeline_matches_synchronous_calls(rd):
    """rdb_gradient_async / rdb_tape_wait through rd.GradientPipeline: six consecutive evaluations with different inputs over two
    engine copies of the tape, page-locked host arrays; every result must be bit-identical to the synchronous gradient! of the
    same input (same kernels, same order), values included."""
    import torch
    n = 512                                                   # the tcgen05 int8-slice GEMMs and the panel logic are in play
    rng = np.random.default_rng(7)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((n, n)), rng.random((n, n))))
    ct = rd.compile(tape)

    def pinned(shape):
        t = torch.empty(int(np.prod(shape)), dtype=torch.float64).pin_memory()
        return t, t.numpy().reshape(shape, order="F")
    keep, ins, outs = [], [], []
    for k in range(6):
        (ta, a), (tb, b) = pinned((n, n)), pinned((n, n))
        a[...] = rng.random((n, n)); b[...] = rng.random((n, n))
        (tga, ga), (tgb, gb) = pinned((n, n)), pinned((n, n))
        keep += [ta, tb, tga, tgb]
        ins.append((a, b)); outs.append((rd.DiffResult(0.0, ga), rd.DiffResult(0.0, gb)))
    pipe = rd.GradientPipeline(tape, depth=2)
    for k in range(6):
        pipe.submit(outs[k], ins[k])
    pipe.drain()
    for k in range(6):
        ref = (rd.DiffResult(0.0, np.zeros((n, n), order="F")), rd.DiffResult(0.0, np.zeros((n, n), order="F")))
        rd.gradient_(ref, ct, ins[k])
        assert np.array_equal(outs[k][0].gradient, ref[0].gradient) and np.array_equal(outs[k][1].gradient, ref[1].gradient), k
        assert outs[k][0].value == ref[0].value and outs[k][1].value == ref[1].value
    # a slot refuses a second asynchronous call before its first one was completed
    h = pipe.slots[0].handle
    rd.seeded_forward_pass(pipe.slots[0], ins[0])
    h.gradient_async([outs[0][0].gradient, outs[0][1].gradient], [(n, n), (n, n)])
    with pytest.raises(Exception):
        h.gradient_async([outs[0][0].gradient, outs[0][1].gradient], [(n, n), (n, n)])
    h.wait()


def test_record_and_replay_with_device_arrays(rd):
    """a tape recorded from device arrays (shapes-only: no element of them crosses to the host, see
    tests/test_record_shapes_only.py) replays with device inputs and device results - no host array is ever handed to the engine."""
    import torch
    n = 300
    rng = np.random.default_rng(9)
    a0, b0 = rng.random((n, n)), rng.random((n, n))
    a, b = rng.random((n, n)), rng.random((n, n))
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rd.engine.device_array(a0), rd.engine.device_array(b0)))
    ref = rd.GradientTape(rd.workloads.f_gradient_example, (a0, b0))
    assert tape.tape.shapes_only and tape.program_bytes() == ref.program_bytes()
    ct = rd.compile(tape)
    ga = torch.empty(n * n, dtype=torch.float64, device="cuda"); gb = torch.empty_like(ga)
    rd.gradient_((ga, gb), ct, (rd.engine.device_array(a), rd.engine.device_array(b)))
    assert relerr(ga.cpu().numpy().reshape(n, n, order="F"), b.sum(1)[:, None] + b.sum(0)[None, :]) < 1e-12
    assert relerr(gb.cpu().numpy().reshape(n, n, order="F"), a.sum(1)[:, None] + a.sum(0)[None, :]) < 1e-12


def _f_many_args(rd, k):
    E, T = rd.expr, rd.tracked

    def f(*xs):
        # one fused dot-broadcast of k array arguments (broadcast.jl:142-146 is N-ary): matrices, a column vector and a row vector
        def g(*v):
            acc = v[0] * v[1]
            for i in range(2, len(v)):
                acc = acc + (E.tanh(v[i]) * v[i - 1] if i % 2 else v[i] * v[i] * 0.5)
            return acc
        return T.sum(T.bcast(g, *xs))
    return f


def _many_args_inputs(k, seed, m=24, n=40):
    rng = np.random.default_rng(seed)
    shapes = [(m, n)] * k
    shapes[2] = (m,)                 # broadcasts along columns
    shapes[-1] = (1, n)              # broadcasts along rows
    return [np.asfortranarray(rng.random(s) - 0.3) for s in shapes]


@pytest.mark.parametrize("k", [5, 6, 8])
def test_nbcast_more_than_four_arguments(rd, orc, k):
    """`∇broadcast` with more than four array arguments: operands 5.. travel in OP_MORE_OPERANDS records; NVRTC kernel on the
    device (the interpreter in a child process with RDB_JIT=0), oracle on the host."""
    run_gradient(rd, orc, _f_many_args(rd, k), _many_args_inputs(k, 1), _many_args_inputs(k, 2), np.float64)


def test_nbcast_many_arguments_interpreter_in_child_process():
    import subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ); e["RDB_JIT"] = "0"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-k", "more_than_four or run_fusion", "-p", "no:cacheprovider"],
                       cwd=os.path.dirname(here), env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("block", [0, 5])
def test_hessian_bilinear_mul_of_two_tracked_operands(rd, orc, block):
    """hessian! through `*` with BOTH operands tracked (d2(a b) carries the cross terms delta tv(b)' and tv(a)' delta): X*X with the
    same buffer on both sides, X * g.(X), and a chain into a third product; against the closed form where there is one and
    against finite differences of the oracle's gradient."""
    rng = np.random.default_rng(6)
    m = 4
    C = rng.random((m, m))

    def f_square(x):                                         # f = sum(C .* (X X)):  H[(i,k),(k',j)] = C[i,j] d(k,k') + C[k',k... ] (closed form below)
        X = x.reshape(m, m)
        return rd.sum(rd.bcast(lambda u, c: u * c, X @ X, C))

    def f_mixed(x):
        X = x.reshape(m, m)
        Y = rd.bcast(lambda u: rd.exp(u * 0.5), X)
        Z = X @ Y                                             # two different tracked buffers
        return rd.sum(rd.bcast(rd.tanh, Z @ X + Y))            # and a product whose left operand carries tangents already

    x0 = rng.random(m * m); x = rng.random(m * m) - 0.3
    for f in (f_square, f_mixed):
        tape = rd.HessianTape(f, x0)
        ct = rd.compile(tape, seed_block=block)
        g = np.zeros(m * m); Hn = np.zeros((m * m, m * m), order="F")
        res = rd.DiffResult(0.0, g, Hn)
        rd.hessian_(res, ct, x)
        val, (go,) = orc.OracleTape(tape.program_bytes()).gradient([x])
        assert abs(res.value - val) < 1e-12 * abs(val) and relerr(g, go) < 1e-12
        Hfd = orc.hessian_dense(tape.program_bytes(), x)
        assert np.max(np.abs(Hn - Hfd)) < 1e-6 * max(1.0, np.max(np.abs(Hfd)))
        assert np.max(np.abs(Hn - Hn.T)) < 1e-12 * np.max(np.abs(Hn))
        if f is f_square:
            # f = sum_ijk C[i,j] X[i,k] X[k,j]  =>  d2f / dX[a,b] dX[c,d] = C[a,d] [b == c] + C[c,b] [d == a]   (column-major vec)
            Hcf = np.zeros((m * m, m * m))
            for a_ in range(m):
                for b_ in range(m):
                    for c_ in range(m):
                        for d_ in range(m):
                            Hcf[a_ + b_ * m, c_ + d_ * m] = C[a_, d_] * (b_ == c_) + C[c_, b_] * (d_ == a_)
            assert relerr(Hn, Hcf) < 1e-13


@pytest.mark.parametrize("block", [0, 3])
def test_hessian_broadcasting_closures(rd, orc, block):
    """hessian! through elementwise closures whose arguments broadcast against the output (column vector, row vector, an untracked
    constant matrix): tangents gathered through the clamps, adjoint tangents reduced into the clamped arguments."""
    rng = np.random.default_rng(8)
    m, n = 4, 3
    C = rng.random((m, n))

    def f(x):
        X = x[0:m * n].reshape(m, n)
        b = x[m * n:m * n + m]                                   # length-m vector: broadcasts along columns
        r = x[m * n + m:m * n + m + n].reshape(1, n)             # 1 x n row: broadcasts along rows
        Y = rd.bcast(lambda u, v, w, c: rd.tanh(u * v + w) * c + v * v * w, X, b, r, C)
        return rd.sum(rd.bcast(lambda y, v: y * y + rd.exp(v * 0.3), Y, b))

    N = m * n + m + n
    x0 = rng.random(N); x = rng.random(N) - 0.4
    tape = rd.HessianTape(f, x0)
    ct = rd.compile(tape, seed_block=block)
    g = np.zeros(N); Hn = np.zeros((N, N), order="F")
    res = rd.DiffResult(0.0, g, Hn)
    rd.hessian_(res, ct, x)
    val, (go,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert abs(res.value - val) < 1e-12 * abs(val) and relerr(g, go) < 1e-12
    Hfd = orc.hessian_dense(tape.program_bytes(), x)
    assert np.max(np.abs(Hn - Hfd)) < 1e-6 * max(1.0, np.max(np.abs(Hfd)))
    assert np.max(np.abs(Hn - Hn.T)) < 1e-12 * np.max(np.abs(Hn))
