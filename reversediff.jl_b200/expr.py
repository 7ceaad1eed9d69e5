"""Scalar-closure tracing: turns the ``f`` of ``f.(args...)`` / ``map(@forward(f), ...)`` into
SSA bytecode.

In the reference the closure lives inside the instruction cache as ``df`` and is re-run with
ForwardDiff ``Dual`` numbers on every forward replay (``src/derivatives/broadcast.jl:147-153,200``,
``src/derivatives/elementwise.jl:230-310,323-343``).  A Julia closure cannot cross a C ABI, so
``compile`` runs it once on symbolic ``Sym`` reals and ships the resulting straight-line program;
the device interpreter then evaluates value + partials per element exactly as one ``Dual{N}``
evaluation would.  Closures with data-dependent branches cannot be traced — the same limit the
reference documents for compiled tapes (``docs/src/api.md:39-44``).
"""
from __future__ import annotations

import math
import numpy as np

from . import program as P


class Tracer:
    """``dual[i]`` says whether argument ``i`` reaches the closure as a ForwardDiff ``Dual`` (tracked arguments of
    ``map``/``broadcast``/scalar instructions, EVERY argument of ``∇broadcast``, ``broadcast.jl:142-153``) or as the plain
    ``Real`` it is (untracked arguments of ``@forward`` closures, ``elementwise.jl:107-167,234-301``; literals).  Only ``^``
    cares: ``Dual^Real``, ``Real^Dual`` and ``Dual^Dual`` are three different ForwardDiff methods (see ``pow_imm``)."""

    def __init__(self, n_args: int, dual=None, pow_flavor: int = 0):
        if n_args > P.MAX_ARGS:
            raise ValueError(f"elementwise closures with more than {P.MAX_ARGS} array arguments are not supported")
        self.n_args = n_args
        self.dual = [True] * n_args if dual is None else [bool(d) for d in dual]
        self.pow_flavor = int(pow_flavor)
        self.ops: list[tuple[int, int, int, int]] = []
        self.fconsts: list[float] = []

    def emit(self, code, a=0, b=0, imm=0, dual=True) -> "Sym":
        self.ops.append((code, a, b, imm))
        reg = self.n_args + len(self.ops) - 1
        if reg >= P.MAX_EXPR_REGS:
            raise ValueError(f"scalar closure too large (> {P.MAX_EXPR_REGS} SSA values)")
        return Sym(self, reg, dual)

    def const(self, c) -> "Sym":
        self.fconsts.append(float(c))
        return self.emit(P.E_CONST, 0, 0, len(self.fconsts) - 1, dual=False)

    def lift(self, x) -> "Sym":
        if isinstance(x, Sym):
            if x.tr is not self:
                raise ValueError("symbolic value from a different closure trace")
            return x
        if isinstance(x, (int, float, np.floating, np.integer)):
            return self.const(x)
        raise TypeError(f"cannot trace through a value of type {type(x).__name__}")


class Sym:
    """Symbolic ``Real`` used to run a closure once at compile time."""

    __slots__ = ("tr", "reg", "dual")
    __array_priority__ = 1000

    def __init__(self, tr: Tracer, reg: int, dual: bool = True):
        self.tr, self.reg, self.dual = tr, reg, dual

    def _bin(self, code, other, swap=False):
        o = self.tr.lift(other)
        a, b = (o, self) if swap else (self, o)
        imm = pow_imm(self.tr.pow_flavor, a.dual, b.dual) if code == P.E_POW else 0
        return self.tr.emit(code, a.reg, b.reg, imm, dual=a.dual or b.dual)

    def __add__(self, o): return self._bin(P.E_ADD, o)
    def __radd__(self, o): return self._bin(P.E_ADD, o, True)
    def __sub__(self, o): return self._bin(P.E_SUB, o)
    def __rsub__(self, o): return self._bin(P.E_SUB, o, True)
    def __mul__(self, o): return self._bin(P.E_MUL, o)
    def __rmul__(self, o): return self._bin(P.E_MUL, o, True)
    def __truediv__(self, o): return self._bin(P.E_DIV, o)
    def __rtruediv__(self, o): return self._bin(P.E_DIV, o, True)
    def __neg__(self): return self.tr.emit(P.E_NEG, self.reg, dual=self.dual)
    def __pos__(self): return self

    def __pow__(self, o):
        if isinstance(o, (int, np.integer)) and not isinstance(o, bool):
            return self.tr.emit(P.E_POWI, self.reg, 0, int(o), dual=self.dual)   # Base.literal_pow
        return self._bin(P.E_POW, o)

    def __rpow__(self, o): return self._bin(P.E_POW, o, True)

    def _nobranch(self, *_):
        raise TypeError("data-dependent control flow cannot be recorded on a compiled tape "
                        "(reference docs/src/api.md:39-44)")

    __bool__ = __lt__ = __le__ = __gt__ = __ge__ = _nobranch


def pow_imm(flavor: int, a_dual: bool, b_dual: bool) -> int:
    """``imm`` of an ``E_POW`` op: bits 0-1 = flavor, bit 2 = base is a plain Real, bit 3 = exponent is a plain Real.

    flavor 0: ForwardDiff ``^`` (``Dual^Dual`` checks ``isconstant(y)`` on the whole partials vector; one ``Dual{N}`` evaluation
    for all partials — ``∇broadcast``, ``@forward`` closures, scalar instructions); flavor 1: the same, but every partial comes
    from its own ``Dual{1}`` evaluation (``tracker_∇broadcast``'s ``_deriv``, ``broadcast.jl:221-224``), i.e. ``isconstant`` per
    component; flavor 2: the uncached formulas of ``(broadcast,^)`` (``elementwise.jl:633-643``: ``e*b^(e-1)``, ``log(b)*b^e``)."""
    return (int(flavor) & 3) | (0 if a_dual else 4) | (0 if b_dual else 8)


def _unary(code, npf):
    def f(x):
        if isinstance(x, Sym):
            return x.tr.emit(code, x.reg, dual=x.dual)
        if hasattr(x, "_scalar_unary"):          # TrackedReal: DiffRules unary function -> one ScalarInstruction (scalars.jl:11-12)
            return x._scalar_unary(code)
        return npf(x)
    return f


tanh = _unary(P.E_TANH, np.tanh)
exp = _unary(P.E_EXP, np.exp)
log = _unary(P.E_LOG, np.log)
sqrt = _unary(P.E_SQRT, np.sqrt)
sin = _unary(P.E_SIN, np.sin)
cos = _unary(P.E_COS, np.cos)
abs2 = _unary(P.E_ABS2, lambda v: v * v)
abs_ = _unary(P.E_ABS, np.abs)
inv = _unary(P.E_INV, lambda v: 1.0 / v)
logistic = _unary(P.E_SIGMOID, lambda v: 1.0 / (1.0 + np.exp(-v)))
log1p = _unary(P.E_LOG1P, np.log1p)
expm1 = _unary(P.E_EXPM1, np.expm1)
asin = _unary(P.E_ASIN, np.arcsin)
acos = _unary(P.E_ACOS, np.arccos)
sinh = _unary(P.E_SINH, np.sinh)
cosh = _unary(P.E_COSH, np.cosh)
tan = _unary(P.E_TAN, np.tan)
_atan1 = _unary(P.E_ATAN, np.arctan)
# third tranche; sec/csc/cot/sech/csch/coth are the reciprocals Julia defines them as (Base: sec(z) = inv(cos(z)), ...)
exp2 = _unary(P.E_EXP2, np.exp2)
exp10 = _unary(P.E_EXP10, lambda v: np.power(10.0, v))
log2 = _unary(P.E_LOG2, np.log2)
log10 = _unary(P.E_LOG10, np.log10)
cbrt = _unary(P.E_CBRT, np.cbrt)
asinh = _unary(P.E_ASINH, np.arcsinh)
acosh = _unary(P.E_ACOSH, np.arccosh)
atanh = _unary(P.E_ATANH, np.arctanh)
sec = _unary(P.E_SEC, lambda v: 1.0 / np.cos(v))
csc = _unary(P.E_CSC, lambda v: 1.0 / np.sin(v))
cot = _unary(P.E_COT, lambda v: 1.0 / np.tan(v))
sech = _unary(P.E_SECH, lambda v: 1.0 / np.cosh(v))
csch = _unary(P.E_CSCH, lambda v: 1.0 / np.sinh(v))
coth = _unary(P.E_COTH, lambda v: 1.0 / np.tanh(v))


def _binary(code, npf):
    def f(x, y):
        if isinstance(x, Sym):
            return x._bin(code, y)
        if isinstance(y, Sym):
            return y._bin(code, x, True)
        for z in (x, y):                         # TrackedReal (or its nested wrapper): one binary ScalarInstruction
            if hasattr(z, "_scalar_binary_code"):
                return z._scalar_binary_code(code, x, y)
        return npf(x, y)
    return f


max_ = _binary(P.E_MAX, np.maximum)
min_ = _binary(P.E_MIN, np.minimum)
hypot = _binary(P.E_HYPOT, np.hypot)
_atan2 = _binary(P.E_ATAN2, np.arctan2)


def atan(y, x=None):
    """``atan(x)`` and the two-argument ``atan(y, x)`` (both in ``DiffRules.diffrules()``)."""
    return _atan1(y) if x is None else _atan2(y, x)


def trace(f, n_args: int, dual=None, pow_flavor: int = 0):
    """Run ``f`` on ``n_args`` symbolic reals; return ``(ops int32[n,4], fconsts float64[])``.  ``dual``/``pow_flavor``: see
    ``Tracer`` / ``pow_imm``."""
    tr = Tracer(n_args, dual, pow_flavor)
    out = f(*[Sym(tr, i, tr.dual[i]) for i in range(n_args)])
    out = tr.lift(out)
    if out.reg < n_args:                       # closure returned one of its arguments
        out = tr.emit(P.E_ARG, 0, 0, out.reg, dual=out.dual)
    elif out.reg != n_args + len(tr.ops) - 1:  # result must be the last SSA value
        out = tr.emit(P.E_ADD, out.reg, tr.const(0.0).reg, dual=out.dual)
    return np.asarray(tr.ops, np.int32).reshape(-1, 4), np.asarray(tr.fconsts, np.float64)


def eval_value(expr, args, dtype):
    """Record-time primal evaluation (host side, values only).

    Mirrors ``out_value = DiffResults.value.(results)`` at record time
    (``broadcast.jl:155-158``); replays never come through here.
    """
    ops, fc = expr
    regs = [np.asarray(a, dtype) for a in args]
    for code, a, b, imm in ops:
        if code == P.E_CONST: r = np.asarray(fc[imm], dtype)
        elif code == P.E_ARG: r = regs[imm]
        elif code == P.E_ADD: r = regs[a] + regs[b]
        elif code == P.E_SUB: r = regs[a] - regs[b]
        elif code == P.E_MUL: r = regs[a] * regs[b]
        elif code == P.E_DIV: r = regs[a] / regs[b]
        elif code == P.E_NEG: r = -regs[a]
        elif code == P.E_POWI: r = _powi(regs[a], int(imm))
        elif code == P.E_POW: r = regs[a] ** regs[b]
        elif code == P.E_TANH: r = np.tanh(regs[a])
        elif code == P.E_EXP: r = np.exp(regs[a])
        elif code == P.E_LOG: r = np.log(regs[a])
        elif code == P.E_SQRT: r = np.sqrt(regs[a])
        elif code == P.E_SIN: r = np.sin(regs[a])
        elif code == P.E_COS: r = np.cos(regs[a])
        elif code == P.E_ABS2: r = regs[a] * regs[a]
        elif code == P.E_ABS: r = np.abs(regs[a])
        elif code == P.E_MAX: r = np.maximum(regs[a], regs[b])
        elif code == P.E_MIN: r = np.minimum(regs[a], regs[b])
        elif code == P.E_INV: r = 1.0 / regs[a]
        elif code == P.E_SIGMOID: r = 1.0 / (1.0 + np.exp(-regs[a]))
        elif code == P.E_LOG1P: r = np.log1p(regs[a])
        elif code == P.E_EXPM1: r = np.expm1(regs[a])
        elif code == P.E_ASIN: r = np.arcsin(regs[a])
        elif code == P.E_ACOS: r = np.arccos(regs[a])
        elif code == P.E_ATAN: r = np.arctan(regs[a])
        elif code == P.E_SINH: r = np.sinh(regs[a])
        elif code == P.E_COSH: r = np.cosh(regs[a])
        elif code == P.E_TAN: r = np.tan(regs[a])
        elif code == P.E_ATAN2: r = np.arctan2(regs[a], regs[b])
        elif code == P.E_HYPOT: r = np.hypot(regs[a], regs[b])
        elif code == P.E_EXP2: r = np.exp2(regs[a])
        elif code == P.E_EXP10: r = np.power(10.0, regs[a])
        elif code == P.E_LOG2: r = np.log2(regs[a])
        elif code == P.E_LOG10: r = np.log10(regs[a])
        elif code == P.E_CBRT: r = np.cbrt(regs[a])
        elif code == P.E_ASINH: r = np.arcsinh(regs[a])
        elif code == P.E_ACOSH: r = np.arccosh(regs[a])
        elif code == P.E_ATANH: r = np.arctanh(regs[a])
        elif code == P.E_SEC: r = 1.0 / np.cos(regs[a])
        elif code == P.E_CSC: r = 1.0 / np.sin(regs[a])
        elif code == P.E_COT: r = 1.0 / np.tan(regs[a])
        elif code == P.E_SECH: r = 1.0 / np.cosh(regs[a])
        elif code == P.E_CSCH: r = 1.0 / np.sinh(regs[a])
        elif code == P.E_COTH: r = 1.0 / np.tanh(regs[a])
        else:
            raise ValueError(f"unknown scalar opcode {code}")
        regs.append(r)
    return np.asarray(regs[-1], dtype)


def _powi(x, n: int):
    if n == 0:
        return np.ones_like(x)
    if n < 0:
        return 1.0 / _powi(x, -n)
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


_ = math
