"""Record-time mirror of ``TrackedReal{TrackedReal}``: how the reference builds a ``HessianTape``.

``HessianTape(f, x)`` (``src/api/tape.jl:285-293``) tracks ``x`` on an OUTER tape ``jtp``, wraps those tracked scalars in a second
layer on an INNER tape ``gtp`` (``HessianConfig``, ``src/api/Config.jl:152-156``), runs ``f`` — every scalar operation records one
``ScalarInstruction`` with its partials on the inner tape (``src/macros.jl:84-131``) while the arithmetic on the wrapped values,
including the DiffRules partial formulas ForwardDiff evaluates on them, records on the outer tape — and then runs the inner
reverse sweep ONCE (``seeded_reverse_pass!``, ``src/api/utils.jl:27-34``; ``scalars.jl:52-74``), whose adjoint arithmetic is
recorded on the outer tape as well.  What is kept is the outer tape only: a scalar tape computing ``x -> grad f(x)``, and
``hessian!`` is its Jacobian (one one-hot seeded reverse sweep per gradient component, ``api/utils.jl:44-58``).

Everything in this file runs once, at record time, on the host — exactly like the reference; the inner tape is never shipped.
The outer tape is the product recorder's scalar block (``tracked.py``), replayed on the device.

ForwardDiff / DiffRules are not part of the reference tree (``Project.toml:23,25``); the partial formulas below are the DiffRules
"1.4" table, spelled on the wrapped values so that each formula term records the outer instruction it would record in Julia.
"""
from __future__ import annotations

import numpy as np

from . import program as P
from . import expr as E
from .tracked import TrackedArray, TrackedReal


def _one_minus_sq(t):
    return 1.0 - t * t


# DiffRules: code -> (value, d/da[, d/db]) on wrapped values
_LN2, _LN10 = float(np.log(2.0)), float(np.log(10.0))
_UNARY = {
    P.E_NEG: (lambda v: -v, lambda v, y: -1.0),
    P.E_TANH: (E.tanh, lambda v, y: _one_minus_sq(y)),                    # 1 - tanh(x)^2
    P.E_EXP: (E.exp, lambda v, y: y),                                      # exp(x)
    P.E_LOG: (E.log, lambda v, y: 1.0 / v),                                # inv(x)
    P.E_SQRT: (E.sqrt, lambda v, y: 1.0 / (2.0 * y)),                      # inv(2 * sqrt(x))
    P.E_SIN: (E.sin, lambda v, y: E.cos(v)),
    P.E_COS: (E.cos, lambda v, y: -E.sin(v)),
    P.E_ABS2: (E.abs2, lambda v, y: 2.0 * v),
    P.E_INV: (E.inv, lambda v, y: -(y * y)),                               # -abs2(inv(x))
    P.E_SIGMOID: (E.logistic, lambda v, y: y * (1.0 - y)),
    P.E_LOG1P: (E.log1p, lambda v, y: 1.0 / (v + 1.0)),                    # inv(x + 1)
    P.E_EXPM1: (E.expm1, lambda v, y: E.exp(v)),
    P.E_ASIN: (E.asin, lambda v, y: 1.0 / E.sqrt(1.0 - v * v)),
    P.E_ACOS: (E.acos, lambda v, y: -(1.0 / E.sqrt(1.0 - v * v))),
    P.E_ATAN: (E.atan, lambda v, y: 1.0 / (1.0 + v * v)),
    P.E_SINH: (E.sinh, lambda v, y: E.cosh(v)),
    P.E_COSH: (E.cosh, lambda v, y: E.sinh(v)),
    P.E_TAN: (E.tan, lambda v, y: 1.0 + y * y),
    P.E_EXP2: (E.exp2, lambda v, y: y * _LN2),                             # exp2(x) * logtwo
    P.E_EXP10: (E.exp10, lambda v, y: y * _LN10),
    P.E_LOG2: (E.log2, lambda v, y: 1.0 / v / _LN2),                       # inv(x) / logtwo
    P.E_LOG10: (E.log10, lambda v, y: 1.0 / v / _LN10),
    P.E_CBRT: (E.cbrt, lambda v, y: 1.0 / (3.0 * (y * y))),                # inv(3 * cbrt(x)^2)
    P.E_ASINH: (E.asinh, lambda v, y: 1.0 / E.sqrt(v * v + 1.0)),
    P.E_ACOSH: (E.acosh, lambda v, y: 1.0 / E.sqrt(v * v - 1.0)),
    P.E_ATANH: (E.atanh, lambda v, y: 1.0 / (1.0 - v * v)),
    P.E_SEC: (E.sec, lambda v, y: y * E.tan(v)),
    P.E_CSC: (E.csc, lambda v, y: -(y * E.cot(v))),
    P.E_COT: (E.cot, lambda v, y: -(1.0 + y * y)),
    P.E_SECH: (E.sech, lambda v, y: -(E.tanh(v) * y)),
    P.E_CSCH: (E.csch, lambda v, y: -(E.coth(v) * y)),
    P.E_COTH: (E.coth, lambda v, y: -(E.csch(v) * E.csch(v))),
}
_BINARY = {
    P.E_ADD: (lambda a, b: a + b, lambda a, b, y: 1.0, lambda a, b, y: 1.0),
    P.E_SUB: (lambda a, b: a - b, lambda a, b, y: 1.0, lambda a, b, y: -1.0),
    P.E_MUL: (lambda a, b: a * b, lambda a, b, y: b, lambda a, b, y: a),
    P.E_DIV: (lambda a, b: a / b, lambda a, b, y: 1.0 / b, lambda a, b, y: -(a / b / b)),   # (one(x)/y, -(x/y/y))
    P.E_POW: (lambda a, b: a ** b, lambda a, b, y: b * a ** (b - 1.0), lambda a, b, y: y * E.log(a)),
    P.E_ATAN2: (lambda a, b: E.atan(a, b), lambda a, b, y: b / (a * a + b * b), lambda a, b, y: -a / (a * a + b * b)),
    P.E_HYPOT: (E.hypot, lambda a, b, y: a / y, lambda a, b, y: b / y),
}


class InnerTape:
    """The inner ``InstructionTape`` (gtp): ``(inputs, output, partials)`` per ScalarInstruction."""

    def __init__(self):
        self.instrs = []


class InnerReal:
    """``TrackedReal{V,D}`` whose ``value``/``deriv`` are scalars of the OUTER tape (or plain numbers)."""

    __slots__ = ("value", "deriv", "tape")
    __array_ufunc__ = None
    __array_priority__ = 2000

    def __init__(self, value, tape):
        self.value, self.deriv, self.tape = value, 0.0, tape

    def _record(self, inputs, value, partials):
        out = InnerReal(value, self.tape)
        self.tape.instrs.append((inputs, out, partials))
        return out

    def _scalar_binary_code(self, code, a, b):
        return self._binary(code, a, b)

    def _binary(self, code, a, b):
        if code not in _BINARY:
            raise TypeError(f"no derivative rule for scalar function {P.E_NAMES.get(code, code)} in the nested recorder")
        fv, fa, fb = _BINARY[code]
        ta, tb = isinstance(a, InnerReal), isinstance(b, InnerReal)
        va, vb = (a.value if ta else a), (b.value if tb else b)
        y = fv(va, vb)
        ins, parts = [], []
        if ta:
            ins.append(a); parts.append(fa(va, vb, y))
        if tb:
            ins.append(b); parts.append(fb(va, vb, y))
        return self._record(ins, y, parts)

    def __add__(self, o): return self._binary(P.E_ADD, self, o)
    def __radd__(self, o): return self._binary(P.E_ADD, o, self)
    def __sub__(self, o): return self._binary(P.E_SUB, self, o)
    def __rsub__(self, o): return self._binary(P.E_SUB, o, self)
    def __mul__(self, o): return self._binary(P.E_MUL, self, o)
    def __rmul__(self, o): return self._binary(P.E_MUL, o, self)
    def __truediv__(self, o): return self._binary(P.E_DIV, self, o)
    def __rtruediv__(self, o): return self._binary(P.E_DIV, o, self)
    def __neg__(self): return self._scalar_unary(P.E_NEG)
    def __pos__(self): return self

    def __pow__(self, o):
        if isinstance(o, (int, np.integer)) and not isinstance(o, bool):
            v, n = self.value, int(o)
            part = float(n) * v ** (n - 1) if n not in (0, 1) else float(n)       # y * x^(y-1)
            return self._record([self], v ** n, [part])
        return self._binary(P.E_POW, self, o)

    def __rpow__(self, o): return self._binary(P.E_POW, o, self)

    def _scalar_unary(self, code):          # the hook expr.tanh/exp/... dispatch on
        if code not in _UNARY:
            raise TypeError(f"no derivative rule for scalar function {P.E_NAMES.get(code, code)} in the nested recorder")
        fv, fd = _UNARY[code]
        y = fv(self.value)
        return self._record([self], y, [fd(self.value, y)])

    def _cmp(self, o):
        return float(_plain(self)), float(_plain(o))

    def __lt__(self, o): a, b = self._cmp(o); return a < b
    def __le__(self, o): a, b = self._cmp(o); return a <= b
    def __gt__(self, o): a, b = self._cmp(o); return a > b
    def __ge__(self, o): a, b = self._cmp(o); return a >= b
    __hash__ = object.__hash__


def _plain(x):
    while isinstance(x, (InnerReal, TrackedReal)):
        x = x.value
    return x


class InnerArray:
    """``TrackedArray{TrackedReal}`` (gcfg.input, ``Config.jl:154``): indexing yields the wrapped scalars; one wrapper per element,
    so every use of ``x[i]`` accumulates into the same ``deriv`` (in the reference through ``origin.deriv[index]``)."""

    def __init__(self, outer: TrackedArray, tape: InnerTape):
        self.outer = outer
        self.shape = outer.shape
        self.size = outer.size
        self.ndim = outer.ndim
        self.elems = [InnerReal(outer[int(i)], tape) for i in range(outer.size)]

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        return iter(self.elems) if self.ndim == 1 else iter(self[i] for i in range(self.shape[0]))

    def __getitem__(self, key):
        lin = np.arange(self.size).reshape(self.shape, order="F")[key]
        if np.ndim(lin) == 0:
            return self.elems[int(lin)]
        return [self.elems[int(i)] for i in np.asarray(lin).reshape(-1, order="F")]

    def _inner_sum(self):
        acc = self.elems[0]
        for e in self.elems[1:]:
            acc = acc + e
        return acc


def inner_reverse_pass(tape: InnerTape):
    """``reverse_pass!`` over the inner scalar instructions (``src/tape.jl:85-90``; ``scalars.jl:52-74``), executed on wrapped
    values: every ``deriv(output) * partial`` and every ``+=`` records an outer instruction."""
    for ins, out, parts in reversed(tape.instrs):
        for x, p in zip(ins, parts):
            x.deriv = x.deriv + out.deriv * p          # increment_deriv!(input, output_deriv * partial)
        out.deriv = 0.0                                 # unseed!(output)


def record_gradient_program(f, outer: TrackedArray):
    """Run ``f`` on the doubly-tracked input and the inner reverse sweep; returns ``(y, [d f / d x_j])`` as outer-tape scalars."""
    gtp = InnerTape()
    x = InnerArray(outer, gtp)
    y = f(x)
    if not isinstance(y, InnerReal):
        raise TypeError("HessianTape: f must return a tracked scalar")
    for e in x.elems:
        e.deriv = 0.0                                   # unseed!(input)
    y.deriv = 1.0                                       # seed!(output)
    inner_reverse_pass(gtp)
    return y, [e.deriv for e in x.elems]
