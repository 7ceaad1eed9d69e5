"""Scalar-instruction tapes (``ScalarInstruction``, reference ``src/derivatives/scalars.jl:52-123``,
``src/macros.jl:84-131``; scalar ``getindex`` ``src/tracked.jl:348-351``).

CPU part: the recorder, the wire format and the oracle's sequential loops (pure Python and the C helper) against closed forms
and finite differences.  GPU part (``-m gpu``): the level-parallel device replay through the C ABI against the oracle.
"""
import numpy as np
import pytest

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


# ------------------------------------------------------------------------------------------ workloads
def rosenbrock_loop(x):
    """``benchmark/benchmark.jl:60-76`` spelled with scalar indexing: every operation is one ScalarInstruction."""
    acc = 0.0
    for i in range(x.size // 2):
        o, e = x[2 * i], x[2 * i + 1]
        acc = acc + (100.0 * (e - o ** 2) ** 2 + (1.0 - o) ** 2)
    return acc


def make_ackley(rd):
    def ackley(x):
        """``benchmark/benchmark.jl:84-95``."""
        a, b, c = 20.0, -0.2, 2.0 * np.pi
        n = x.size
        s1 = 0.0
        s2 = 0.0
        for i in range(n):
            s1 = s1 + x[i] ** 2
            s2 = s2 + rd.cos(c * x[i])
        return -a * rd.exp(b * rd.sqrt(s1 / n)) - rd.exp(s2 / n) + a + np.e
    return ackley


def ackley_np(x):
    a, b, c = 20.0, -0.2, 2.0 * np.pi
    n = x.size
    return -a * np.exp(b * np.sqrt(np.sum(x * x) / n)) - np.exp(np.sum(np.cos(c * x)) / n) + a + np.e


def make_mixed(rd):
    def mixed(x, w):
        """Array and scalar instructions interleaved: two scalar blocks around array instructions, a TrackedReal produced by an
        array instruction consumed by scalar ones, block slots consumed by a tracker broadcast and by a later block."""
        s = rd.sum(x)                                  # array instruction -> TrackedReal
        y = s * x[0] + rd.tanh(x[1]) / w[2]            # block 1
        z = rd.bcast(lambda v, c: v * c, w, y)         # tracker_∇broadcast with a block slot as the scalar argument
        t = rd.sum(z)
        u = (t - y) * x[0] * x[0]                      # block 2: reads a slot of block 1, the same leaf twice
        return u * u + rd.forward(lambda p, q: p * rd.exp(q) - q)(y, x[3])
    return mixed


def mixed_np(x, w):
    s = np.sum(x)
    y = s * x[0] + np.tanh(x[1]) / w[2]
    t = np.sum(w * y)
    u = (t - y) * x[0] * x[0]
    return u * u + (y * np.exp(x[3]) - x[3])


def fd_grad(f, xs, k, h=1e-6):
    g = np.zeros(xs[k].size)
    for i in range(xs[k].size):
        xp = [v.copy() for v in xs]; xm = [v.copy() for v in xs]
        xp[k][i] += h; xm[k][i] -= h
        g[i] = (f(*xp) - f(*xm)) / (2 * h)
    return g


def make_vecfun(rd):
    def vecfun(x):
        """R^n -> R^n, written element by element; returned as an Array{TrackedReal}."""
        n = x.size
        out = []
        for i in range(n):
            a, b = x[i], x[(i + 1) % n]
            out.append(a * rd.sin(b) + b / (1.0 + a ** 2))
        out[1] = x[1]            # a tape input element passed straight through
        out[2] = 3.0             # a plain number
        return out
    return vecfun


def vecfun_jac(x):
    n = x.size
    J = np.zeros((n, n))
    for i in range(n):
        j = (i + 1) % n
        a, b = x[i], x[j]
        J[i, i] += np.sin(b) - b * 2 * a / (1.0 + a * a) ** 2
        J[i, j] += a * np.cos(b) + 1.0 / (1.0 + a * a)
    J[1, :] = 0; J[1, 1] = 1
    J[2, :] = 0
    return J


# ------------------------------------------------------------------------------------------ CPU: recorder + oracle
def test_recorder_one_block_no_getindex_instructions(rd):
    x0 = np.random.default_rng(1).random(8)
    tape = rd.GradientTape(rosenbrock_loop, x0)
    assert len(tape) == 1                                            # one maximal run of scalar instructions
    blk = tape.tape.program.instrs[0].scalar
    assert len(blk) == 4 * 8                                          # o^2, e-., (.)^2, 100*., 1-o, (.)^2, +, acc+ per pair
    assert blk.refbufs[0] == tape.tape.program.inputs[0]
    raw = tape.program_bytes()
    assert rd.program.TapeProgram.from_bytes(raw).to_bytes() == raw   # wire round trip


@pytest.mark.parametrize("use_c", [False, True])
def test_oracle_rosenbrock_closed_form(rd, orc, use_c):
    n = 64
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.GradientTape(rosenbrock_loop, x0)
    orc._ScalarBlock.use_c = use_c
    try:
        if use_c:
            assert orc._scalar_lib() is not None, "oracle/libscalar_block.so not built (run __graft_entry__.build())"
        val, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    finally:
        orc._ScalarBlock.use_c = True
    o, e = x[0::2], x[1::2]
    ref = np.zeros(n); ref[0::2] = -400 * o * (e - o * o) - 2 * (1 - o); ref[1::2] = 200 * (e - o * o)
    assert abs(val - np.sum(100 * (e - o * o) ** 2 + (1 - o) ** 2)) < 1e-12 * abs(val)
    assert relerr(g, ref) < 1e-13


def test_oracle_c_and_python_loops_agree(rd, orc):
    """Same statements in both: bit-identical for pure arithmetic; within an ulp or two where libm and numpy differ (exp, tanh)."""
    x0 = np.random.default_rng(1).random(16); w0 = np.random.default_rng(3).random(5)
    x = np.random.default_rng(2).random(16); w = np.random.default_rng(4).random(5)
    for f, rec, inp, exact in ((make_mixed(rd), (x0, w0), [x, w], False), (rosenbrock_loop, x0, [x], True)):
        tape = rd.GradientTape(f, rec)
        res = []
        for use_c in (False, True):
            orc._ScalarBlock.use_c = use_c
            try:
                res.append(orc.OracleTape(tape.program_bytes()).gradient(inp))
            finally:
                orc._ScalarBlock.use_c = True
        if exact:
            assert res[0][0] == res[1][0]
        for a, b in zip(res[0][1], res[1][1]):
            assert np.array_equal(a, b) if exact else relerr(a, b) < 1e-14


def test_oracle_mixed_and_ackley_finite_differences(rd, orc):
    x0 = np.random.default_rng(1).random(16); w0 = np.random.default_rng(3).random(5)
    x = np.random.default_rng(2).random(16); w = np.random.default_rng(4).random(5)
    tape = rd.GradientTape(make_mixed(rd), (x0, w0))
    assert [i.opcode for i in tape.tape.program.instrs].count(rd.program.OP_SCALAR_BLOCK) == 2
    val, (gx, gw) = orc.OracleTape(tape.program_bytes()).gradient([x, w])
    assert abs(val - mixed_np(x, w)) < 1e-12 * abs(val)
    assert relerr(gx, fd_grad(mixed_np, [x, w], 0)) < 1e-6
    assert relerr(gw, fd_grad(mixed_np, [x, w], 1)) < 1e-6
    tape = rd.GradientTape(make_ackley(rd), x0)
    val, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert abs(val - ackley_np(x)) < 1e-12 * abs(val)
    assert relerr(g, fd_grad(ackley_np, [x], 0)) < 1e-6


def make_converted(rd):
    def f(x, w):
        """``convert`` (``src/tracked.jl:237-257``) between scalar and array instructions: of a scalar-getindex leaf, of a block
        slot, of the TrackedReal an array instruction produced; the same source converted twice."""
        a = rd.convert(x[2])                           # leaf -> slot
        s = rd.convert(rd.sum(w))                      # output of an array instruction -> slot
        y = rd.convert(a * s + x[0])                   # slot -> slot
        z = rd.sum(rd.bcast(lambda v, c: v * c, w, rd.convert(y)))
        return z * rd.convert(x[2]) + rd.exp(rd.convert(y))
    return f


def converted_np(x, w):
    y = x[2] * np.sum(w) + x[0]
    return np.sum(w * y) * x[2] + np.exp(y)


def test_oracle_convert_instruction(rd, orc):
    x0 = np.random.default_rng(1).random(6); w0 = np.random.default_rng(3).random(5)
    x = np.random.default_rng(2).random(6); w = np.random.default_rng(4).random(5)
    tape = rd.GradientTape(make_converted(rd), (x0, w0))
    val, (gx, gw) = orc.OracleTape(tape.program_bytes()).gradient([x, w])
    assert abs(val - converted_np(x, w)) < 1e-13 * abs(val)
    assert relerr(gx, fd_grad(converted_np, [x, w], 0)) < 1e-6 and relerr(gw, fd_grad(converted_np, [x, w], 1)) < 1e-6
    y = x[2] * np.sum(w) + x[0]
    gx_cf = np.zeros(6); gx_cf[0] = np.sum(w) * x[2] + np.exp(y); gx_cf[2] = (np.sum(w) * x[2] + np.exp(y)) * np.sum(w) + np.sum(w * y)
    assert relerr(gx, gx_cf) < 1e-13
    assert rd.convert(3.5) == 3.5


@pytest.mark.gpu
def test_gpu_convert_instruction(rd, orc):
    x0 = np.random.default_rng(1).random(6); w0 = np.random.default_rng(3).random(5)
    x = np.random.default_rng(2).random(6); w = np.random.default_rng(4).random(5)
    tape = rd.GradientTape(make_converted(rd), (x0, w0))
    ct = rd.compile(tape)
    res = (rd.DiffResult(0.0, np.zeros(6)), rd.DiffResult(0.0, np.zeros(5)))
    rd.gradient_(res, ct, (x, w))
    val, (gx, gw) = orc.OracleTape(tape.program_bytes()).gradient([x, w])
    assert abs(res[0].value - val) <= 1e-12 * abs(val)
    assert relerr(res[0].gradient, gx) < 1e-12 and relerr(res[1].gradient, gw) < 1e-12


def test_oracle_jacobian_of_tracked_real_array(rd, orc):
    n = 12
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.JacobianTape(make_vecfun(rd), x0)
    val, J = orc.OracleTape(tape.program_bytes()).jacobian([x])
    assert relerr(J, vecfun_jac(x)) < 1e-13
    assert val[2] == 3.0 and val[1] == x[1]


def test_compiled_tape_freezes_branches(rd, orc):
    """Predicates on TrackedReals are evaluated at record time only (SkipOptimize, scalars.jl:24-46; docs/src/api.md:39-44)."""
    f = lambda x: x[0] * 2.0 if x[0] > 0.5 else x[0] * 3.0
    tape = rd.GradientTape(f, np.array([0.9, 0.0]))
    _, (g,) = orc.OracleTape(tape.program_bytes()).gradient([np.array([0.1, 0.0])])
    assert g[0] == 2.0


# ------------------------------------------------------------------------------------------ GPU parity
def _check_all_buffers(ct, ot, orc, tol):
    for b in range(len(ot.value)):
        if ot.kind[b] == orc.BUF_CONST:
            continue
        assert relerr(ct.handle.get_value(b), ot.value[b].reshape(-1, order="F")) < tol, f"value of buffer {b}"
        d = ct.handle.get_deriv(b); od = ot.deriv[b].reshape(-1, order="F")
        if np.any(od):
            assert relerr(d, od) < tol
        else:
            assert not np.any(d), f"deriv of buffer {b} not unseeded"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [2, 8, 1024, 40000])
def test_gpu_gradient_rosenbrock_loop(rd, orc, dtype, n):
    x0 = np.random.default_rng(1).random(n).astype(dtype); x = np.random.default_rng(2).random(n).astype(dtype)
    tape = rd.GradientTape(rosenbrock_loop, x0, dtype=dtype)
    ct = rd.compile(tape)
    for rep in range(3):       # replays: eager, graph capture, graph launch for small tapes
        res = rd.DiffResult(0.0, np.zeros(n, dtype))
        rd.gradient_(res, ct, x)
        ot = orc.OracleTape(tape.program_bytes())
        val, (g,) = ot.gradient([x])
        assert abs(float(res.value) - float(val)) <= TOL[dtype] * abs(float(val))
        assert relerr(res.gradient, g) < TOL[dtype]
    if dtype == np.float64:
        o, e = x[0::2], x[1::2]
        ref = np.zeros(n); ref[0::2] = -400 * o * (e - o * o) - 2 * (1 - o); ref[1::2] = 200 * (e - o * o)
        assert relerr(res.gradient, ref) < 1e-12
        if n <= 1024:
            assert np.array_equal(res.gradient, g)      # same operations in the same order, no contraction: bit-identical


@pytest.mark.gpu
def test_gpu_gradient_mixed_blocks(rd, orc):
    x0 = np.random.default_rng(1).random(16); w0 = np.random.default_rng(3).random(5)
    x = np.random.default_rng(2).random(16); w = np.random.default_rng(4).random(5)
    tape = rd.GradientTape(make_mixed(rd), (x0, w0))
    for use_graph in (0, 1):
        ct = rd.compile(tape, use_graph=use_graph)
        for rep in range(3):
            res = (rd.DiffResult(0.0, np.zeros(16)), rd.DiffResult(0.0, np.zeros(5)))
            rd.gradient_(res, ct, (x, w))
            ot = orc.OracleTape(tape.program_bytes())
            val, (gx, gw) = ot.gradient([x, w])
            assert abs(res[0].value - val) <= 1e-12 * abs(val)
            assert relerr(res[0].gradient, gx) < 1e-12 and relerr(res[1].gradient, gw) < 1e-12
        if not use_graph:
            _check_all_buffers(ct, ot, orc, 1e-12)


@pytest.mark.gpu
def test_gpu_gradient_ackley(rd, orc):
    n = 3000
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.GradientTape(make_ackley(rd), x0)
    ct = rd.compile(tape)
    res = rd.DiffResult(0.0, np.zeros(n))
    rd.gradient_(res, ct, x)
    val, (g,) = orc.OracleTape(tape.program_bytes()).gradient([x])
    assert abs(res.value - val) <= 1e-12 * abs(val)
    assert relerr(res.gradient, g) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed_block", [(5, 0), (12, 0), (40, 0), (100, 0), (100, 16), (100, 33), (300, 7)])
def test_gpu_jacobian_of_tracked_real_array(rd, orc, n, seed_block):
    """R < 16 runs the slot-parallel kernel, R >= 16 the seed-lane kernel (lanes = seeds); seed_block forces ragged blocks."""
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.JacobianTape(make_vecfun(rd), x0)
    ct = rd.compile(tape, seed_block=seed_block)
    J = np.zeros((n, n), order="F")
    res = rd.DiffResult(np.zeros(n), J)
    rd.jacobian_(res, ct, x)
    ot = orc.OracleTape(tape.program_bytes())
    val, Jo = ot.jacobian([x])
    assert relerr(J, Jo) < 1e-12 and relerr(J, vecfun_jac(x)) < 1e-12
    assert relerr(res.value, val) < 1e-12
    # row shard (multi-GPU unit of work)
    Jb = np.zeros((n - n // 2, n), order="F")
    rd.jacobian_(Jb, ct, x, rows=(n // 2, n))
    assert np.array_equal(Jb, J[n // 2:, :])


def make_hubfun(rd):
    def hubfun(x):
        """R^n -> R^n with everything the streaming reverse kernel plans around: two hub slots consumed by every output (fan-out far
        beyond the four in-record edges, read hundreds of schedule positions after they are produced), a running chain, values
        produced early and consumed late, one slot exported to two outputs."""
        n = x.size
        hub = x[0] * x[1] + rd.exp(x[2] * 0.5)         # hub slot
        early = [rd.sin(x[i]) * hub for i in range(n)]  # produced first ...
        acc = x[3] * 1.0
        out = []
        for i in range(n):
            acc = acc * 0.9 + x[i] * hub               # chain through the whole tape
            out.append(early[n - 1 - i] * acc + x[(i + 5) % n] / (1.5 + rd.cos(hub)))   # ... consumed last-to-first
        out[4] = out[7]                                # one TrackedReal in two output elements
        return out
    return hubfun


def hubfun_np(x):
    n = x.size
    hub = x[0] * x[1] + np.exp(x[2] * 0.5)
    early = np.sin(x) * hub
    acc = x[3] * 1.0
    out = np.zeros(n)
    for i in range(n):
        acc = acc * 0.9 + x[i] * hub
        out[i] = early[n - 1 - i] * acc + x[(i + 5) % n] / (1.5 + np.cos(hub))
    out[4] = out[7]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,seed_block", [(24, 0), (100, 0), (100, 37), (333, 0), (333, 64)])
def test_gpu_jacobian_hub_chain_long_range(rd, orc, dtype, n, seed_block):
    """Seed-lane streaming kernel (R >= 16): ring / staged / direct operand sources, staged and looped exports, ragged warps."""
    x0 = np.random.default_rng(1).random(n).astype(dtype); x = np.random.default_rng(2).random(n).astype(dtype)
    tape = rd.JacobianTape(make_hubfun(rd), x0)
    ct = rd.compile(tape, seed_block=seed_block)
    J = np.zeros((n, n), dtype=dtype, order="F")
    res = rd.DiffResult(np.zeros(n, dtype=dtype), J)
    rd.jacobian_(res, ct, x)
    val, Jo = orc.OracleTape(tape.program_bytes()).jacobian([x])
    assert relerr(res.value, val) < TOL[dtype]
    assert relerr(J, Jo) < TOL[dtype]
    if dtype is np.float64:
        assert relerr(res.value, hubfun_np(x)) < 1e-12
        h = 1e-6
        for j in (0, 2, 3, n - 1):                       # central differences on a few columns
            xp = x.copy(); xm = x.copy(); xp[j] += h; xm[j] -= h
            assert relerr(J[:, j], (hubfun_np(xp) - hubfun_np(xm)) / (2 * h)) < 1e-6
    # a second sweep on the same compiled tape (ring / staging state must not leak) and a row shard
    J2 = np.zeros((n, n), dtype=dtype, order="F")
    rd.jacobian_(J2, ct, x)
    assert np.array_equal(J2, J)
    Jb = np.zeros((n - n // 3, n), dtype=dtype, order="F")
    rd.jacobian_(Jb, ct, x, rows=(n // 3, n))
    assert np.array_equal(Jb, J[n // 3:, :])


@pytest.mark.gpu
def test_gpu_rejects_bad_scalar_block(rd, ctx):
    P = rd.program
    tape = rd.GradientTape(rosenbrock_loop, np.random.default_rng(1).random(4))
    blk = tape.tape.program.instrs[0].scalar
    blk.a_idx[3] = 10 ** 6                        # operand refers to a slot that does not exist yet
    blk.a_space[3] = P.SP_POOL
    with pytest.raises(Exception):
        rd.compile(tape)


# ------------------------------------------------------------------------------------------ the reference's HessianTape
def rosenbrock_hessian(x):
    n = x.size
    H = np.zeros((n, n))
    o, e = x[0::2], x[1::2]
    i = np.arange(0, n, 2)
    H[i, i] = 1200 * o * o - 400 * e + 2
    H[i, i + 1] = H[i + 1, i] = -400 * o
    H[i + 1, i + 1] = 200
    return H


def test_oracle_reverse_over_reverse_hessian_tape(rd, orc):
    """mode="reverse" records the reference's own Hessian program (api/tape.jl:285-293): the Jacobian of the recorded gradient
    tape must equal the closed form and the independent literal restatement in oracle/scalar_tape.py."""
    import oracle.scalar_tape as st
    n = 8
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.HessianTape(rosenbrock_loop, x0, mode="reverse")
    assert all(i.opcode == rd.program.OP_SCALAR_BLOCK for i in tape.tape.program.instrs)
    g, H = orc.OracleTape(tape.program_bytes()).jacobian([x])
    assert relerr(H, rosenbrock_hessian(x)) < 1e-13
    assert np.all(H[rosenbrock_hessian(x) == 0] == 0)                 # structural zeros exact
    _, g2, H2 = st.hessian_reverse_over_reverse(lambda v: rosenbrock_loop(np.asarray(v, dtype=object)), list(x))
    assert relerr(H, np.array(H2)) < 1e-13 and relerr(g, np.array(g2)) < 1e-13


def test_oracle_reverse_hessian_ackley_matches_forward_mode_oracle(rd, orc):
    n = 6
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.HessianTape(make_ackley(rd), x0, mode="reverse")
    g, H = orc.OracleTape(tape.program_bytes()).jacobian([x])
    gt = rd.GradientTape(make_ackley(rd), x0)
    Hfd = orc.hessian_dense(gt.program_bytes(), x)
    assert relerr(H, Hfd) < 1e-6 and relerr(H, H.T) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 8, 64, 512])
def test_gpu_reverse_over_reverse_hessian(rd, orc, n):
    x0 = np.random.default_rng(1).random(n); x = np.random.default_rng(2).random(n)
    tape = rd.HessianTape(rosenbrock_loop, x0, mode="reverse")
    ct = rd.compile(tape)
    H = np.zeros((n, n), order="F")
    res = rd.DiffResult(0.0, np.zeros(n), H)
    rd.hessian_(res, ct, x)
    Href = rosenbrock_hessian(x)
    assert relerr(H, Href) < 1e-12
    assert np.all(H[Href == 0] == 0)
    o, e = x[0::2], x[1::2]
    gref = np.zeros(n); gref[0::2] = -400 * o * (e - o * o) - 2 * (1 - o); gref[1::2] = 200 * (e - o * o)
    assert relerr(res.gradient, gref) < 1e-12
    assert abs(res.value - np.sum(100 * (e - o * o) ** 2 + (1 - o) ** 2)) < 1e-12 * abs(res.value)
    if n <= 64:
        _, Ho = orc.OracleTape(tape.program_bytes()).jacobian([x])
        assert relerr(H, Ho) < 1e-12
    # the north star's mechanism on the array tape gives the same matrix
    ft = rd.HessianTape(lambda v: rd.sum(rd.map_(rd.forward(lambda o_, e_: 100.0 * (e_ - o_ ** 2) ** 2 + (1.0 - o_) ** 2), v[0::2], v[1::2])), x0)
    Hf = np.zeros((n, n), order="F")
    rd.hessian_(Hf, rd.compile(ft), x)
    assert relerr(H, Hf) < 1e-12
    # row shard
    Hb = np.zeros((n - n // 2, n), order="F")
    rd.hessian_(Hb, ct, x, rows=(n // 2, n))
    assert np.array_equal(Hb, H[n // 2:, :])


@pytest.mark.gpu
def test_gpu_forward_mode_hessian_rejects_scalar_tapes(rd):
    tape = rd.HessianTape(rosenbrock_loop, np.random.default_rng(1).random(4))      # mode="forward" on a scalar-instruction tape
    ct = rd.compile(tape)
    with pytest.raises(Exception, match="scalar instructions"):
        rd.hessian_(np.zeros((4, 4), order="F"), ct, np.ones(4))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"RDB_SCALAR_BUNDLES": "1", "RDB_SCALAR_MANYSEED": "b"},      # bundle kernels everywhere
                                 {"RDB_SCALAR_BUNDLES": "0", "RDB_SCALAR_MANYSEED": "s"}])     # level kernels + streaming kernel
def test_gpu_every_scalar_kernel_variant(env):
    """The planner picks a kernel per graph shape and seed count (depth > 48 levels -> bundles; 16+ seeds -> streaming kernel or
    one bundle warp per seed by a cost model), so at default settings a small test graph only ever meets some of them.  The
    choices are process-wide (read once from the environment): rerun the GPU tests of this file in a child process with each
    kernel family forced on."""
    import os, subprocess, sys
    here = os.path.dirname(os.path.abspath(__file__))
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_scalar_tape.py"), "-m", "gpu", "-q", "-x",
                        "-k", "not every_scalar_kernel_variant", "-p", "no:cacheprovider"],
                       cwd=os.path.dirname(here), env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
