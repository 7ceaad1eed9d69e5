"""Second and third tranche of the DiffRules scalar-function whitelist (``log1p expm1 asin acos atan sinh cosh tan atan(y,x) hypot``,
``exp2 exp10 log2 log10 cbrt asinh acosh atanh sec csc cot sech csch coth``;
reference: ``src/derivatives/scalars.jl:5-22``, ``src/derivatives/elementwise.jl:70-76,186-230``), after the pattern of
``test/derivatives/ElementwiseTests.jl:22-66``: every function as ``map``/``broadcast`` of a ``@forward`` closure, as a fused
dot-broadcast, and as a scalar instruction, value and derivative against closed forms (CPU: oracle) and oracle against device
(``-m gpu``), second derivatives through ``hessian!``."""
import numpy as np
import pytest

UNARY = {
    # name: (numpy value, closed-form derivative, second derivative)
    "log1p": (np.log1p, lambda x: 1 / (1 + x), lambda x: -1 / (1 + x) ** 2),
    "expm1": (np.expm1, np.exp, np.exp),
    "asin": (np.arcsin, lambda x: 1 / np.sqrt(1 - x * x), lambda x: x / (1 - x * x) ** 1.5),
    "acos": (np.arccos, lambda x: -1 / np.sqrt(1 - x * x), lambda x: -x / (1 - x * x) ** 1.5),
    "atan": (np.arctan, lambda x: 1 / (1 + x * x), lambda x: -2 * x / (1 + x * x) ** 2),
    "sinh": (np.sinh, np.cosh, np.sinh),
    "cosh": (np.cosh, np.sinh, np.cosh),
    "tan": (np.tan, lambda x: 1 + np.tan(x) ** 2, lambda x: 2 * np.tan(x) * (1 + np.tan(x) ** 2)),
    # third tranche: exp2 exp10 log2 log10 cbrt asinh acosh atanh sec csc cot sech csch coth
    "exp2": (np.exp2, lambda x: np.exp2(x) * np.log(2), lambda x: np.exp2(x) * np.log(2) ** 2),
    "exp10": (lambda x: 10.0 ** x, lambda x: 10.0 ** x * np.log(10), lambda x: 10.0 ** x * np.log(10) ** 2),
    "log2": (np.log2, lambda x: 1 / (x * np.log(2)), lambda x: -1 / (x * x * np.log(2))),
    "log10": (np.log10, lambda x: 1 / (x * np.log(10)), lambda x: -1 / (x * x * np.log(10))),
    "cbrt": (np.cbrt, lambda x: x ** (-2 / 3) / 3, lambda x: -2 * x ** (-5 / 3) / 9),
    "asinh": (np.arcsinh, lambda x: 1 / np.sqrt(x * x + 1), lambda x: -x / (x * x + 1) ** 1.5),
    # (acosh lives on x > 1 and the test data in [0.1, 0.9]: tested as x -> acosh(x + 1.5))
    "acosh": (lambda x: np.arccosh(x + 1.5), lambda x: 1 / np.sqrt((x + 1.5) ** 2 - 1), lambda x: -(x + 1.5) / ((x + 1.5) ** 2 - 1) ** 1.5),
    "atanh": (np.arctanh, lambda x: 1 / (1 - x * x), lambda x: 2 * x / (1 - x * x) ** 2),
    "sec": (lambda x: 1 / np.cos(x), lambda x: np.tan(x) / np.cos(x), lambda x: (np.tan(x) ** 2 + 1 / np.cos(x) ** 2) / np.cos(x)),
    "csc": (lambda x: 1 / np.sin(x), lambda x: -1 / (np.sin(x) * np.tan(x)), lambda x: (1 / np.tan(x) ** 2 + 1 / np.sin(x) ** 2) / np.sin(x)),
    "cot": (lambda x: 1 / np.tan(x), lambda x: -1 / np.sin(x) ** 2, lambda x: 2 * np.cos(x) / np.sin(x) ** 3),
    "sech": (lambda x: 1 / np.cosh(x), lambda x: -np.tanh(x) / np.cosh(x), lambda x: (np.tanh(x) ** 2 - 1 / np.cosh(x) ** 2) / np.cosh(x)),
    "csch": (lambda x: 1 / np.sinh(x), lambda x: -1 / (np.sinh(x) * np.tanh(x)), lambda x: (1 / np.tanh(x) ** 2 + 1 / np.sinh(x) ** 2) / np.sinh(x)),
    "coth": (lambda x: 1 / np.tanh(x), lambda x: -1 / np.sinh(x) ** 2, lambda x: 2 * np.cosh(x) / np.sinh(x) ** 3),
}
BINARY = {
    # name: (numpy value, d/da, d/db)
    "atan2": (np.arctan2, lambda a, b: b / (a * a + b * b), lambda a, b: -a / (a * a + b * b)),
    "hypot": (np.hypot, lambda a, b: a / np.hypot(a, b), lambda a, b: b / np.hypot(a, b)),
}


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def fn(rd, name):
    return {"atan2": lambda a, b: rd.atan(a, b), "acosh": lambda x: rd.acosh(x + 1.5)}.get(name, getattr(rd, name, None))


def data(seed, shape=(7, 5)):
    return np.asfortranarray(np.random.default_rng(seed).random(shape) * 0.8 + 0.1)


# ------------------------------------------------------------------------------------------ CPU: recorder + oracle
@pytest.mark.parametrize("name", sorted(UNARY))
def test_oracle_unary(rd, orc, name):
    f, d1, _ = UNARY[name]
    g = fn(rd, name)
    x0, x = data(1), data(2)
    for spell in ("map", "bcast", "scalar"):
        if spell == "map":
            tape = rd.GradientTape(lambda v: rd.sum(rd.map_(rd.forward(g), v)), x0)
        elif spell == "bcast":
            tape = rd.GradientTape(lambda v: rd.sum(rd.bcast(lambda u: g(u) * 2.0, v)), x0)
        else:
            tape = rd.GradientTape(lambda v: g(v[3]) * v[0] + g(v[0] * v[1]), x0.reshape(-1))
        xin = x.reshape(-1) if spell == "scalar" else x
        val, (grad,) = orc.OracleTape(tape.program_bytes()).gradient([xin])
        if spell == "map":
            assert abs(val - f(x).sum()) < 1e-13 * abs(val) and relerr(grad, d1(x)) < 1e-13
        elif spell == "bcast":
            assert relerr(grad, 2 * d1(x)) < 1e-13
        else:
            v = xin
            ref = np.zeros_like(v); ref[3] = d1(v[3]) * v[0]; ref[0] = f(v[3]) + d1(v[0] * v[1]) * v[1]; ref[1] = d1(v[0] * v[1]) * v[0]
            assert relerr(grad, ref) < 1e-13


@pytest.mark.parametrize("name", sorted(UNARY))
def test_oracle_unary_second_order(rd, orc, name):
    """The oracle's hyper-dual evaluation (what the device's second-order kernels are checked against): value, first and second
    derivative of every unary function against the closed forms."""
    f, d1, d2 = UNARY[name]
    g = fn(rd, name)
    x = np.random.default_rng(2).random(12) * 0.8 + 0.1
    ops, fc = rd.expr.trace(lambda u: g(u), 1)
    v, D, H = orc.dual_eval(np.asarray(ops, np.int32).reshape(-1, 4), fc, [x], np.float64, order=2)
    assert relerr(v, f(x)) < 1e-14 and relerr(D[0], d1(x)) < 1e-14 and relerr(H[0][0], d2(x)) < 1e-13


@pytest.mark.parametrize("name", sorted(BINARY))
def test_oracle_binary(rd, orc, name):
    f, da, db = BINARY[name]
    g = fn(rd, name)
    a0, b0, a, b = data(1), data(2), data(3), data(4)
    tape = rd.GradientTape(lambda u, v: rd.sum(rd.map_(rd.forward(g), u, v)), (a0, b0))
    val, (ga, gb) = orc.OracleTape(tape.program_bytes()).gradient([a, b])
    assert abs(val - f(a, b).sum()) < 1e-13 * abs(val)
    assert relerr(ga, da(a, b)) < 1e-13 and relerr(gb, db(a, b)) < 1e-13
    # one tracked, one plain argument (the @forward closure sees a Dual and a Real), and as scalar instructions
    tape = rd.GradientTape(lambda u: rd.sum(rd.map_(rd.forward(g), u, b)), a0)
    _, (ga2,) = orc.OracleTape(tape.program_bytes()).gradient([a])
    assert relerr(ga2, da(a, b)) < 1e-13
    tape = rd.GradientTape(lambda v: g(v[0], v[1]) * g(v[2], 0.5), a0.reshape(-1))
    v = a.reshape(-1)
    _, (gs,) = orc.OracleTape(tape.program_bytes()).gradient([v])
    ref = np.zeros_like(v)
    ref[0] = da(v[0], v[1]) * f(v[2], 0.5); ref[1] = db(v[0], v[1]) * f(v[2], 0.5); ref[2] = f(v[0], v[1]) * da(v[2], 0.5)
    assert relerr(gs, ref) < 1e-13


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", sorted(UNARY) + sorted(BINARY))
def test_gpu_functions_every_spelling(rd, orc, name, dtype):
    g = fn(rd, name)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    a0, b0 = data(1).astype(dtype), data(2).astype(dtype)
    a, b = data(3).astype(dtype), data(4).astype(dtype)
    binary = name in BINARY

    def f_arrays(u, v):
        if binary:
            m = rd.map_(rd.forward(g), u, v)
            w = rd.bcast(lambda p, q: g(p, q) * q, u, m)                 # fused dot-broadcast (NVRTC kernel; a run with `m`)
        else:
            m = rd.map_(rd.forward(g), u)
            w = rd.bcast(lambda p, q: g(p * 0.9) + q, v, m)
        return rd.sum(w)

    def f_scalars(u, v):
        x, y = u.reshape(-1), v.reshape(-1)
        acc = 0.0
        for i in range(6):
            acc = acc + (g(x[i], y[i]) * y[i + 1] if binary else g(x[i]) * y[i])
        return acc

    for f in (f_arrays, f_scalars):
        tape = rd.GradientTape(f, (a0, b0), dtype=dtype)
        ct = rd.compile(tape)
        res = (rd.DiffResult(0.0, np.zeros(a.shape, dtype, order="F")), rd.DiffResult(0.0, np.zeros(b.shape, dtype, order="F")))
        rd.gradient_(res, ct, (a, b))
        val, (ga, gb) = orc.OracleTape(tape.program_bytes()).gradient([a, b])
        assert abs(float(res[0].value) - float(val)) <= tol * abs(float(val))
        assert relerr(res[0].gradient, ga) < tol and relerr(res[1].gradient, gb) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(UNARY) + sorted(BINARY))
def test_gpu_functions_second_order(rd, orc, name):
    """hessian! of sum(g.(...)): the second partials of every new function (device: general sweep with host result, fused kernel
    with device result) against closed forms / central differences of the closed-form first derivative."""
    import torch
    g = fn(rd, name)
    n = 12
    x0 = np.random.default_rng(1).random(n) * 0.8 + 0.1
    x = np.random.default_rng(2).random(n) * 0.8 + 0.1
    if name in UNARY:
        f = lambda v: rd.sum(rd.bcast(lambda u: g(u), v))
        Href = np.diag(UNARY[name][2](x))
    else:
        f = lambda v: rd.sum(rd.map_(rd.forward(g), v[0:n // 2], v[n // 2:n]))
        _, da, db = BINARY[name]
        Href = np.zeros((n, n)); h = 1e-6
        for i in range(n // 2):
            j = i + n // 2
            Href[i, i] = (da(x[i] + h, x[j]) - da(x[i] - h, x[j])) / (2 * h)
            Href[j, j] = (db(x[i], x[j] + h) - db(x[i], x[j] - h)) / (2 * h)
            Href[i, j] = Href[j, i] = (da(x[i], x[j] + h) - da(x[i], x[j] - h)) / (2 * h)
    ct = rd.compile(rd.HessianTape(f, x0))
    Hh = np.zeros((n, n), order="F")
    rd.hessian_(Hh, ct, x)
    tol = 1e-12 if name in UNARY else 1e-7
    assert relerr(Hh, Href) < tol
    Hd = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
    rd.hessian_(Hd, ct, torch.from_numpy(x).cuda())
    assert relerr(Hd.cpu().numpy().reshape(n, n, order="F"), Hh) < 1e-14
