"""CPU tests (-m "not gpu"): the oracle restatement against the reference's known-answer literals,
independent closed forms and finite differences; the host recorder; the wire format."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_literals.json")))


def relerr(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


# ------------------------------------------------------------------ reference known-answer literals
def test_literal_issue235(rd, orc):
    g = GOLD["issue235_vector_adjoint_times_matrix"]
    A = np.array(g["A"], float)
    x = np.array(g["x"], float)
    tape = rd.GradientTape(lambda y: rd.sum(y.T @ A), x)
    assert len(tape) == 2
    _, (grad,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert np.array_equal(grad, np.array(g["gradient"], float))      # reference asserts ==


@pytest.mark.parametrize("key,c", [("compat_abs2_minus_zeros", np.zeros(3)), ("compat_abs2_minus_range", np.arange(1.0, 4.0))])
def test_literal_compat(rd, orc, key, c):
    g = GOLD[key]
    x = np.array(g["x"], float)
    tape = rd.GradientTape(lambda v: rd.sum(rd.bcast(lambda a, b: rd.abs2(a - b), v, c)), x)
    _, (grad,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert np.array_equal(grad, np.array(g["gradient"], float))


def test_index_iterables(rd):
    """getindex index iterables (TrackedTests.jl:641-716): our 0-based linear maps must enumerate the same
    elements in the same (column-major) order as the reference's Cartesian iterables."""
    keys = {"[:, :]": (slice(None), slice(None)), "[:, 1:2]": (slice(None), slice(0, 2)),
            "[2:3, :]": (slice(1, 3), slice(None)), "[1:2, 2:3]": (slice(0, 2), slice(1, 3)),
            "[2:6]": (slice(1, 6),), "[:]": (slice(None),)}
    for case in GOLD["index_iterables_3x3"]["cases"]:
        tape = rd.InstructionTape()
        v = np.arange(9.0).reshape(3, 3, order="F")
        ta = rd.TrackedArray(v, tape, tape.program.add_buffer(rd.program.BUF_TA, v.shape))
        ta[keys[case["index"]]]
        assert len(tape) == 1
        ins = tape.program.instrs[0]
        assert ins.opcode == rd.program.OP_GETINDEX
        if "expect_1based" in case:
            lin = [(i - 1) + (j - 1) * 3 for i, j in case["expect_1based"]]
        else:
            lin = [i - 1 for i in case["expect_1based_linear"]]
        assert list(ins.inputs[0].idx_map) == lin


def test_transpose_map_matches_scalar_getindex(rd):
    c = GOLD["scalar_getindex_linear_index"]["cases"][0]
    i, j = c["ij"]
    rows, cols = c["shape"]
    # element (j, i) of the adjoint is parent[i, j] whose 1-based linear index is the reference's TrackedReal.index
    m = rd.program.transpose_map(rows, cols)
    k = (j - 1) + (i - 1) * cols
    assert m[k] + 1 == c["index_1based"]


# ------------------------------------------------------------------ closed forms
@pytest.mark.parametrize("n", [1, 3, 100])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cfg1_closed_form(rd, orc, n, dtype):
    rng = np.random.default_rng(1)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((n, n)).astype(dtype), rng.random((n, n)).astype(dtype)))
    assert [i.opcode for i in tape.tape.program.instrs] == [rd.program.OP_MUL, rd.program.OP_MUL, rd.program.OP_ADD, rd.program.OP_SUM]
    a, b = rd.workloads.inputs_gradient_example(n, dtype)
    val, (ga, gb) = orc.OracleTape(tape.program_bytes()).gradient([a, b])
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    one = np.ones(n)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert abs(val - ((a64 @ one) @ (b64 @ one) + (a64.T @ one) @ (b64.T @ one))) <= tol * abs(val)
    assert relerr(ga, b64.sum(1)[:, None] + b64.sum(0)[None, :]) < tol
    assert relerr(gb, a64.sum(1)[:, None] + a64.sum(0)[None, :]) < tol
    assert ga.dtype == dtype


def test_cfg3_mlp_backprop(rd, orc):
    d, h, o, N = 7, 5, 4, 9
    X, (W1, b1, W2, b2) = rd.workloads.inputs_mlp(d, h, o, N, seed=1)
    tape = rd.GradientTape(rd.workloads.make_f_mlp(X), (W1, b1, W2, b2))
    P = rd.program
    assert [i.opcode for i in tape.tape.program.instrs] == [P.OP_MUL, P.OP_NBCAST, P.OP_MUL, P.OP_NBCAST, P.OP_SUM]
    _, (W1, b1, W2, b2) = rd.workloads.inputs_mlp(d, h, o, N, seed=2)
    val, (gW1, gb1, gW2, gb2) = orc.OracleTape(tape.program_bytes()).gradient([W1, b1, W2, b2])
    H = np.tanh(W1 @ X + b1[:, None]); Y = np.tanh(W2 @ H + b2[:, None])
    dZ2 = 1 - Y ** 2
    dH = W2.T @ dZ2
    dZ1 = dH * (1 - H ** 2)
    assert abs(val - Y.sum()) < 1e-12 * abs(Y.sum())
    assert relerr(gW2, dZ2 @ H.T) < 1e-12 and relerr(gb2, dZ2.sum(1)) < 1e-12
    assert relerr(gW1, dZ1 @ X.T) < 1e-12 and relerr(gb1, dZ1.sum(1)) < 1e-12


def test_cfg4_jacobian_closed_form(rd, orc):
    n = 12
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    P = rd.program
    assert [i.opcode for i in tape.tape.program.instrs] == [P.OP_MUL, P.OP_ADD, P.OP_MUL, P.OP_NBCAST]
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    y, J = orc.OracleTape(tape.program_bytes()).jacobian([x])
    s = A @ x + x; v = B @ x; t = np.tanh(v)
    assert relerr(y, s * t) < 1e-13
    Jcf = np.diag(t) @ (A + np.eye(n)) + np.diag(s * (1 - t ** 2)) @ B
    assert relerr(J, Jcf) < 1e-12


def test_cfg5_rosenbrock_gradient_and_fd_hessian(rd, orc):
    n = 10
    x0 = rd.workloads.inputs_rosenbrock(n, seed=1)
    tape = rd.HessianTape(rd.workloads.f_rosenbrock, x0)
    P = rd.program
    assert [i.opcode for i in tape.tape.program.instrs] == [P.OP_GETINDEX, P.OP_GETINDEX, P.OP_MAP, P.OP_SUM]
    x = rd.workloads.inputs_rosenbrock(n, seed=2)
    val, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    o, e = x[0::2], x[1::2]
    gcf = np.zeros(n)
    gcf[0::2] = -400 * o * (e - o ** 2) - 2 * (1 - o)
    gcf[1::2] = 200 * (e - o ** 2)
    assert abs(val - np.sum(100 * (e - o ** 2) ** 2 + (1 - o) ** 2)) < 1e-12 * abs(val)
    assert relerr(g, gcf) < 1e-12
    H = orc.hessian_dense(tape.program_bytes(), x)
    Hcf = np.zeros((n, n))
    for i in range(n // 2):
        Hcf[2 * i, 2 * i] = 1200 * o[i] ** 2 - 400 * e[i] + 2
        Hcf[2 * i, 2 * i + 1] = Hcf[2 * i + 1, 2 * i] = -400 * o[i]
        Hcf[2 * i + 1, 2 * i + 1] = 200
    assert relerr(H, Hcf) < 1e-6      # finite differences


# ------------------------------------------------------------------ finite differences per instruction
def _fd_grad(make_tape_bytes, xs, orc, eps=1e-6):
    t = orc.OracleTape(make_tape_bytes)
    grads = []
    for k, x in enumerate(xs):
        g = np.zeros(x.size)
        flat = x.reshape(-1, order="F")
        for i in range(x.size):
            xp = [y.copy(order="F") for y in xs]; xm = [y.copy(order="F") for y in xs]
            xp[k].reshape(-1, order="F")[i] += eps; xm[k].reshape(-1, order="F")[i] -= eps
            for s, y in enumerate(xp): t.set_input(s, y)
            t.forward(); vp = float(np.asarray(t.value[t.output]).reshape(()))
            for s, y in enumerate(xm): t.set_input(s, y)
            t.forward(); vm = float(np.asarray(t.value[t.output]).reshape(()))
            g[i] = (vp - vm) / (2 * eps)
        grads.append(g.reshape(x.shape, order="F"))
        _ = flat
    return grads


CASES = {
    "a*b": lambda rd: (lambda a, b: rd.sum(a @ b)),
    "a'*b": lambda rd: (lambda a, b: rd.sum(a.T @ b)),
    "a*b'": lambda rd: (lambda a, b: rd.sum(a @ b.T)),
    "a'*b'": lambda rd: (lambda a, b: rd.sum(a.T @ b.T)),
    "a+b": lambda rd: (lambda a, b: rd.sum(a + b)),
    "a-b": lambda rd: (lambda a, b: rd.mean(a - b)),
    "-a": lambda rd: (lambda a, b: rd.sum((-a) @ b)),
    "bcast*": lambda rd: (lambda a, b: rd.sum(rd.broadcast("*", a, b))),
    "bcast-": lambda rd: (lambda a, b: rd.sum(rd.broadcast("-", a, b) @ b)),
    "map": lambda rd: (lambda a, b: rd.sum(rd.map_(rd.forward(lambda x, y: rd.sin(x) * rd.exp(y) / (1.0 + x * x)), a, b))),
    "nbcast_fns": lambda rd: (lambda a, b: rd.sum(rd.bcast(lambda x, y: rd.log(x + 1.5) * rd.sqrt(y + 0.5) + rd.logistic(x) - rd.cos(y) ** 3, a, b))),
    "getindex": lambda rd: (lambda a, b: rd.sum(a[0:2, :] @ b[:, 1:3])),
    "reshape": lambda rd: (lambda a, b: rd.sum(a.reshape(9) + b.reshape(9))),
    "bcast/": lambda rd: (lambda a, b: rd.sum(rd.broadcast("/", a, rd.bcast(lambda y: y + 1.0, b)))),
    "bcast\\": lambda rd: (lambda a, b: rd.sum(rd.broadcast("\\", rd.bcast(lambda y: y + 1.0, a), b))),
    "bcast^": lambda rd: (lambda a, b: rd.sum(rd.broadcast("^", rd.bcast(lambda y: y + 0.5, a), b))),
    "dot": lambda rd: (lambda a, b: rd.dot(a, b)),
    "sum_dims1": lambda rd: (lambda a, b: rd.dot(rd.sum(a, dims=1), rd.sum(b, dims=1))),
    "sum_dims2": lambda rd: (lambda a, b: rd.sum(rd.broadcast("*", rd.sum(a, dims=2), b))),
    "tracker_real": lambda rd: (lambda a, b: rd.sum(rd.bcast(lambda u, c, v: u * c + rd.tanh(v), a, 2.5, b))),
    "tracker_trackedreal": lambda rd: (lambda a, b: rd.sum(rd.bcast(lambda u, s: u * s, a, rd.sum(b)))),
    "getindex_cartesian": lambda rd: (lambda a, b: rd.dot(a[[(0, 0), (2, 1), (1, 2), (2, 1)]], b[[(1, 1), (0, 2), (2, 2), (0, 0)]])),
    "getindex_generic": lambda rd: (lambda a, b: rd.sum(a[[0, 2, 2], :] @ b[:, [1, 2]])),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_instruction_fd(rd, orc, name):
    rng = np.random.default_rng(7)
    a0, b0 = rng.random((3, 3)), rng.random((3, 3))
    tape = rd.GradientTape(CASES[name](rd), (a0, b0))
    a, b = np.asfortranarray(rng.random((3, 3))), np.asfortranarray(rng.random((3, 3)))
    raw = tape.program_bytes()
    _, (ga, gb) = orc.OracleTape(raw).gradient([a, b])
    fa, fb = _fd_grad(raw, [a, b], orc)
    assert np.max(np.abs(ga - fa)) < 1e-6 and np.max(np.abs(gb - fb)) < 1e-6


def test_broadcast_bias_and_row(rd, orc):
    rng = np.random.default_rng(3)
    W, b, r = rng.random((4, 6)), rng.random(4), rng.random((1, 6))
    tape = rd.GradientTape(lambda W_, b_, r_: rd.sum(rd.bcast(lambda x, y, z: rd.tanh(x + y) * z, W_, b_, r_)), (W, b, r))
    raw = tape.program_bytes()
    _, gs = orc.OracleTape(raw).gradient([W, b, r])
    fs = _fd_grad(raw, [np.asfortranarray(W), b, np.asfortranarray(r)], orc)
    for g, f in zip(gs, fs):
        assert np.max(np.abs(g - f)) < 1e-6


def test_intermediate_derivs_are_unseeded(rd, orc):
    """After a full reverse sweep only tape inputs hold non-zero derivs (every reverse exec ends in unseed!)."""
    rng = np.random.default_rng(1)
    tape = rd.GradientTape(rd.workloads.f_gradient_example, (rng.random((4, 4)), rng.random((4, 4))))
    t = orc.OracleTape(tape.program_bytes())
    t.gradient(list(rd.workloads.inputs_gradient_example(4)))
    for b in range(len(t.deriv)):
        if t.deriv[b] is not None and b not in t.inputs:
            assert not np.any(t.deriv[b])


# ------------------------------------------------------------------ wire format / recorder behaviour
def test_program_roundtrip(rd):
    X, ws = rd.workloads.inputs_mlp(5, 4, 3, 6)
    tape = rd.GradientTape(rd.workloads.make_f_mlp(X), ws)
    raw = tape.program_bytes()
    assert rd.program.TapeProgram.from_bytes(raw).to_bytes() == raw


def test_capture_freezes_constants(rd, orc):
    """`capture` copies untracked arrays into the tape (tracked.jl:225; TapeTests.jl:8-38)."""
    A = np.ones((2, 2)); x = np.ones(2)
    tape = rd.GradientTape(lambda v: rd.sum(A @ v), x)
    A[:] = 5.0
    _, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert np.array_equal(g, [2.0, 2.0])


def test_recorder_rejections(rd):
    x = np.ones(3)
    with pytest.raises(TypeError):
        rd.GradientTape(lambda v: rd.sum(rd.map_(rd.forward(lambda a, b: a * b), v, 2.0)), x)   # map needs arrays
    assert len(rd.GradientTape(lambda v: v[0], x)) == 1      # scalar getindex records nothing (tracked.jl:348-351); the tape holds
    #                                                          only the identity slot that gives the output a buffer of its own
    with pytest.raises(TypeError):
        rd.GradientTape(lambda v: rd.sum(rd.bcast(lambda a: a if a > 0 else -a, v)), x)  # data-dependent branch
    with pytest.raises(TypeError):
        rd.HessianTape(lambda a, b: rd.sum(a + b), (x, x))                            # hessians.jl:103-108
    with pytest.raises(TypeError):
        rd.GradientTape(lambda v: v + v, x)                                           # non-scalar output


def test_skip_records_nothing(rd):
    tape = rd.GradientTape(lambda v: rd.sum(v + rd.skip(lambda w: w * 2.0)(v)), np.ones(3))
    assert len(tape) == 2                                                             # `+` with a frozen constant, sum


def test_uncompiled_replay_is_refused(rd):
    tape = rd.GradientTape(lambda v: rd.sum(v + v), np.ones(3))
    with pytest.raises(TypeError):
        rd.gradient_(np.zeros(3), tape, np.ones(3))


# ------------------------------------------------------------------ the reference's own Hessian algorithm (reverse-over-reverse)
def _rosen_scalar(v):
    s = 0.0
    for i in range(0, len(v), 2):
        o, e = v[i], v[i + 1]
        s = s + (100.0 * (e - o ** 2) ** 2 + (1.0 - o) ** 2)
    return s


def test_scalar_tape_reverse_over_reverse_rosenbrock(rd, orc):
    """oracle/scalar_tape.py restates HessianTape literally (api/tape.jl:285-293); pin it on the closed form and on the
    array-tape oracle's gradient."""
    from oracle import scalar_tape as st
    n = 12
    x = rd.workloads.inputs_rosenbrock(n, seed=2)
    val, g, H = st.hessian_reverse_over_reverse(_rosen_scalar, x)
    H = np.array(H)
    o, e = x[0::2], x[1::2]
    Hcf = np.zeros((n, n)); idx = np.arange(n // 2)
    Hcf[2 * idx, 2 * idx] = 1200 * o ** 2 - 400 * e + 2
    Hcf[2 * idx, 2 * idx + 1] = Hcf[2 * idx + 1, 2 * idx] = -400 * o
    Hcf[2 * idx + 1, 2 * idx + 1] = 200
    assert relerr(H, Hcf) < 1e-14
    assert np.all(H[Hcf == 0] == 0.0)
    tape = rd.HessianTape(rd.workloads.f_rosenbrock, x)
    oval, (og,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert relerr(g, og) < 1e-14 and abs(val - oval) < 1e-13 * abs(oval)


def test_golden_program_hashes_are_current():
    """tests/golden/cfg*.program.sha256 is the byte-level contract the Julia lowering (julia/ReverseDiffB200.jl, checked by
    julia/golden_roundtrip.jl) is held to: the Python recorder must still emit exactly those bytes, or the hashes must be
    regenerated on purpose (tests/golden/make_program_hashes.py)."""
    import hashlib, importlib.util, os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_program_hashes", os.path.join(here, "make_program_hashes.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    for name, raw in mod.programs().items():
        want, size = open(os.path.join(here, f"{name}.program.sha256")).read().split()
        assert hashlib.sha256(raw).hexdigest() == want and len(raw) == int(size), name


def test_oracle_nbcast_with_eight_arguments(rd, orc):
    """N-ary `∇broadcast` (broadcast.jl:142-146) beyond the four operand slots of one record: the continuation records round-trip
    through the wire format and the oracle's gradient matches central differences."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("tgp", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_gpu_parity.py"))
    tgp = importlib.util.module_from_spec(spec); spec.loader.exec_module(tgp)
    k = 8
    f = tgp._f_many_args(rd, k)
    x0, x = tgp._many_args_inputs(k, 1, 5, 7), tgp._many_args_inputs(k, 2, 5, 7)
    tape = rd.GradientTape(f, tuple(x0))
    raw = tape.program_bytes()
    prog = rd.program.TapeProgram.from_bytes(raw)
    assert len(prog.instrs) == 2 and len(prog.instrs[0].inputs) == k and prog.to_bytes() == raw
    ot = orc.OracleTape(raw)
    val, grads = ot.gradient(x)
    h = 1e-6
    for a in (0, 2, 4, 7):
        fd = np.zeros_like(x[a])
        for i in np.ndindex(x[a].shape):
            xp = [v.copy() for v in x]; xm = [v.copy() for v in x]
            xp[a][i] += h; xm[a][i] -= h
            fd[i] = (ot.gradient(xp)[0] - ot.gradient(xm)[0]) / (2 * h)
        assert np.max(np.abs(grads[a] - fd)) < 1e-6 * max(1.0, np.max(np.abs(fd)))
